#!/usr/bin/env python
"""Whole-capture goldens from the UNMODIFIED reference (oracle/_ref/libref_harness.so), made in the
development container (needs /root/reference and `make -C oracle`):

    python tests/golden/make_golden_full.py

Writes
  ref_peaks_nottingham_full.npy       (snr, lo_shift, ca_shift, sv) of ALL 10,880 chunks (340 runs) of
                                      gps.samples.1bit.I.fs5456.if4092.bin as the reference's own
                                      Sample()+Correlate() return them (what SearchTask() prints, unrounded)
  nottingham_strided_runs.bin         whole runs 100-103, 200-203, 336-339 of the capture (12 x 163,840 B):
                                      the part of the whole-file test that ALWAYS travels with the repository
  nottingham_strided_runs.json        which runs those are
  nottingham_marginal_chunks.bin      every chunk of the capture whose reference SNR lies in [24, 26] (the hits and
                                      misses hugging the SNR-25 threshold) + nottingham_marginal_chunks.json
                                      (chunk index, sv) per entry

The reference keeps its state in file statics, so each worker is its own process; workers take
contiguous run ranges.  FFT backend: MKL behind the fftw3.h stand-in (text identical to the built-in FFT).
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
SRC = Path("/root/reference/gps.samples.1bit.I.fs5456.if4092.bin")
sys.path.insert(0, str(ROOT / "oracle"))
import oracle  # noqa: E402

RUN_BYTES = 32 * 5120
STRIDED = list(range(100, 104)) + list(range(200, 204)) + list(range(336, 340))


def worker(lo: int, hi: int, out: str):
    data = SRC.read_bytes()[lo * RUN_BYTES: hi * RUN_BYTES]
    r = oracle.RefHarness(4.092e6, 5.456e6, 5000.0)
    pk = r.search_blocks(data)
    np.save(out, pk[["snr", "lo_shift", "ca_shift", "sv"]])


def main():
    if len(sys.argv) == 5 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
        return
    oracle.build()
    raw = SRC.read_bytes()
    n_runs = len(raw) // RUN_BYTES
    nproc = os.cpu_count() or 1
    bounds = [n_runs * i // nproc for i in range(nproc + 1)]
    tmp = [f"/tmp/ref_full_{i}.npy" for i in range(nproc)]
    procs = [subprocess.Popen([sys.executable, __file__, "--worker", str(bounds[i]), str(bounds[i + 1]), tmp[i]],
                              env=oracle.mkl_env()) for i in range(nproc)]
    for p in procs:
        assert p.wait() == 0
    pk = np.concatenate([np.load(t) for t in tmp])
    assert len(pk) == n_runs * 32
    packed = np.zeros(len(pk), np.dtype([("snr", "<f4"), ("lo_shift", "<i4"), ("ca_shift", "<i4"), ("sv", "<i4")]))
    for k in packed.dtype.names:
        packed[k] = pk[k]
    np.save(HERE / "ref_peaks_nottingham_full.npy", packed)
    (HERE / "nottingham_strided_runs.bin").write_bytes(b"".join(raw[r * RUN_BYTES:(r + 1) * RUN_BYTES] for r in STRIDED))
    json.dump({"runs": STRIDED}, open(HERE / "nottingham_strided_runs.json", "w"))
    marg = np.nonzero((pk["snr"] >= 24.0) & (pk["snr"] <= 26.0))[0]
    (HERE / "nottingham_marginal_chunks.bin").write_bytes(b"".join(raw[c * 5120:(c + 1) * 5120] for c in marg))
    json.dump({"chunk": marg.tolist(), "sv": (marg % 32).tolist()}, open(HERE / "nottingham_marginal_chunks.json", "w"))
    print(n_runs, "runs;", int((pk["snr"] >= 25).sum()), "hits;", len(marg), "chunks with SNR in [24, 26]")


if __name__ == "__main__":
    main()

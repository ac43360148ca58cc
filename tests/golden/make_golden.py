#!/usr/bin/env python
"""Regenerate the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the development container (needs /root/reference and `make -C oracle`):

    python tests/golden/make_golden.py            # everything (full capture: ~4 min with MKL)
    python tests/golden/make_golden.py --no-full  # skip the two whole-file stdout goldens

What is written (all small; the script that made them is this file):
  nottingham_fs5456_if4092_runs0-3.bin   first 4 runs (128 chunks) of gps.samples.1bit.I.fs5456.if4092.bin
  gps_sig_fs8184_if2046_runs0-1.bin      first 2 runs (64 chunks) of gps_sig_tmp.bin
  nottingham_full.stdout.txt             stdout of `gps_test_ref <capture> 4.092e6 5.456e6 5000` (340 runs)
  gps_sig_full.stdout.txt                stdout of `gps_test_ref gps_sig_tmp.bin 2.046e6 8.184e6 5000` (12 runs)
  ref_peaks_*.npy                        (snr, lo_shift, ca_shift) per chunk of the two .bin fixtures, from the
                                         reference's own Sample()+Correlate() via oracle/_ref/libref_harness.so
  ref_probe_*.npz                        256 pseudo-random bins of every replica spectrum code[sv] and of the
                                         spectrum of chunk 0, as the reference holds them
  cacode_first10_octal.json              IS-GPS-200 Table 3-Ia "first 10 chips (octal)" for PRN 1-32, recomputed
                                         from the reference's CACODE via SearchCode-independent chip dump

The reference binary is built with the fftw3.h stand-in (oracle/shim); its printed text is identical
with the built-in float FFT and with MKL (checked over all 340 + 12 runs when this was generated).
"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle"))
import oracle  # noqa: E402

CAPS = {
    "nottingham": dict(src=REF / "gps.samples.1bit.I.fs5456.if4092.bin", fc=4.092e6, fs=5.456e6, runs=4,
                       bin="nottingham_fs5456_if4092_runs0-3.bin", full="nottingham_full.stdout.txt"),
    "gps_sig": dict(src=REF / "gps_sig_tmp.bin", fc=2.046e6, fs=8.184e6, runs=2,
                    bin="gps_sig_fs8184_if2046_runs0-1.bin", full="gps_sig_full.stdout.txt"),
}


def harness_job(name: str):
    """Runs in a subprocess: the reference keeps state in file statics (one config per process)."""
    c = CAPS[name]
    data = (HERE / c["bin"]).read_bytes()
    r = oracle.RefHarness(c["fc"], c["fs"], 5000.0)
    pk = r.search_blocks(data)
    np.save(HERE / f"ref_peaks_{name}.npy", pk[["snr", "lo_shift", "ca_shift", "sv"]])
    rng = np.random.default_rng(20140501)
    idx = np.sort(rng.choice(40000, 256, replace=False)).astype(np.int32)
    code = np.stack([r.code_spectrum(sv)[idx] for sv in range(32)])
    x0 = r.sample(data[:5120])[idx]
    np.savez(HERE / f"ref_probe_{name}.npz", idx=idx, code=code, block0=x0,
             code_abs_sum=np.array([np.abs(r.code_spectrum(sv)).astype(np.float64).sum() for sv in range(32)]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-full", action="store_true")
    ap.add_argument("--job")
    a = ap.parse_args()
    if a.job:
        harness_job(a.job)
        return
    oracle.build(quiet=False)
    gps_test_ref = ROOT / "oracle" / "_ref" / "gps_test_ref"
    for name, c in CAPS.items():
        (HERE / c["bin"]).write_bytes(c["src"].read_bytes()[: c["runs"] * 32 * 5120])
        subprocess.run([sys.executable, __file__, "--job", name], check=True)
        if not a.no_full:
            out = subprocess.run([str(gps_test_ref), str(c["src"]), repr(c["fc"]), repr(c["fs"]), "5000"],
                                 check=True, capture_output=True, env=oracle.mkl_env()).stdout
            (HERE / c["full"]).write_bytes(out)
    # C/A code known-answer: first 10 chips of each PRN as octal (IS-GPS-200 Table 3-Ia)
    firsts = []
    for sv in range(32):
        chips = oracle.cacode_chips(sv)[:10]
        firsts.append(int("".join(str(int(b)) for b in chips), 2))
    json.dump({"first10_octal": [oct(v)[2:] for v in firsts]}, open(HERE / "cacode_first10_octal.json", "w"))


if __name__ == "__main__":
    main()

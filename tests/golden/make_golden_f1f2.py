#!/usr/bin/env python
"""Goldens for the two "next" rows that have reference-made artifacts (development container only):

  f1 (reverse converter)  the UNMODIFIED reference program c/conv_1bit_bin_to_hackrf_bin.cpp (oracle/_ref/conv_1bit_ref)
                          run on the bundled capture -> SHA-256 of its 892,665,856-byte output, and of the prefix that
                          belongs to the committed 4-run fixture (the phase NCO starts at 0, so a prefix of the input
                          gives a prefix of the output).  FC / FS are the macros of c/gps.h (2.6 MHz / 10 MHz) --
                          that is what the program is compiled with, whatever the file name says.
  f2 (signal generator)   the NAV bits gps_sig_gen.m drew with rand, recovered from the bundled gps_sig_tmp.bin, and the
                          file's SHA-256: gps_sig_gen.m:8-41 restated must reproduce the file bit for bit.

Writes tests/golden/f1f2_golden.json.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle"))
import oracle  # noqa: E402


def main():
    oracle.build()
    cap = REF / "gps.samples.1bit.I.fs5456.if4092.bin"
    fixture = (HERE / "nottingham_fs5456_if4092_runs0-3.bin").read_bytes()
    with tempfile.TemporaryDirectory() as d:
        os.symlink(cap, Path(d) / cap.name)
        r = subprocess.run([str(ROOT / "oracle" / "_ref" / "conv_1bit_ref")], cwd=d, capture_output=True, text=True, check=True)
        assert "seems run out!" in r.stdout
        out = np.fromfile(Path(d) / "gps.samples.8bit.IQinterleave.fs5456.if0.bin", np.int8)
    raw = np.fromfile(cap, np.uint8)
    assert out.size == 16 * raw.size
    # the C restatement against the reference program, whole file
    mine = oracle.conv_1bit_iq8(raw, 2.6e6, 10e6, 30)
    assert np.array_equal(mine, out), "oracle restatement differs from the reference converter"
    conv = {"fc": 2.6e6, "fs": 10e6, "amplitude": 30, "n_in_bytes": int(raw.size), "n_out_bytes": int(out.size),
            "sha256_full": hashlib.sha256(out.tobytes()).hexdigest(),
            "fixture": "nottingham_fs5456_if4092_runs0-3.bin", "fixture_in_bytes": len(fixture),
            "sha256_fixture_prefix": hashlib.sha256(out[: 16 * len(fixture)].tobytes()).hexdigest(),
            "first_32_out_bytes": out[:32].tolist()}
    sig = np.fromfile(REF / "gps_sig_tmp.bin", np.uint8)
    nav = oracle.recover_nav_bits(sig, 7)
    assert np.array_equal(oracle.sig_gen_literal(7, nav), sig), "gps_sig_gen.m restatement does not reproduce gps_sig_tmp.bin"
    fx = (HERE / "gps_sig_fs8184_if2046_runs0-1.bin").read_bytes()
    assert fx == sig[: len(fx)].tobytes()
    gen = {"prn": 8, "nav_bits01": nav.tolist(), "n_bytes": int(sig.size), "sha256_file": hashlib.sha256(sig.tobytes()).hexdigest(),
           "fixture": "gps_sig_fs8184_if2046_runs0-1.bin (the first 327,680 bytes of the file)"}
    json.dump({"conv_1bit_bin_to_hackrf_bin": conv, "gps_sig_gen": gen}, open(HERE / "f1f2_golden.json", "w"), indent=1)
    print(json.dumps(conv)[:300]); print(gen["sha256_file"])


if __name__ == "__main__":
    main()

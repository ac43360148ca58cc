#!/usr/bin/env python
"""Golden vectors at sampling rates for which the reference bundles no capture: the UNMODIFIED reference's
(snr, lo_shift, ca_shift) for the synthetic captures that tests/test_gpu_parity.py searches at those rates
(test_synthetic_vs_oracle, test_high_sampling_rates_vs_oracle, test_max_fo_changes_the_doppler_grid).

Run in the development container (needs /root/reference and `make -C oracle`):   python tests/golden/make_golden_rates.py
Writes tests/golden/ref_peaks_rates.npz: one record array per case (key = case name) and the SHA-256 of the case's input
bits, so that a drift of the numpy generator is noticed instead of silently comparing different captures.
The inputs are NOT stored: they are regenerated from (fs, fc, seed, cn0) by gnss_gps_sdr_b200/siggen.py."""
import hashlib
import importlib
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

# name: fs, fc, max_fo, generator arguments -- exactly what the GPU tests build
CASES = {
    "syn_2800": dict(fs=2.8e6, fc=0.62e6, max_fo=5000.0, seed=1575420001, cn0=45.0, chunks=32),
    "syn_10000": dict(fs=10e6, fc=2.6e6, max_fo=5000.0, seed=3, cn0=45.0, chunks=32),
    "syn_5456": dict(fs=5.456e6, fc=4.092e6, max_fo=5000.0, seed=1575420000, cn0=45.0, chunks=32),
    "syn_8000": dict(fs=8.0e6, fc=2.0e6, max_fo=5000.0, seed=11, cn0=45.0, chunks=32),
    "syn_4000": dict(fs=4.0e6, fc=1.0e6, max_fo=5000.0, seed=12, cn0=45.0, chunks=32),
    "hi_16368": dict(fs=16.368e6, fc=4.092e6, max_fo=5000.0, seed=13, cn0=57.0, chunks=32),
    "hi_25000": dict(fs=25e6, fc=6.25e6, max_fo=5000.0, seed=14, cn0=57.0, chunks=32),
    "hi_40000": dict(fs=40e6, fc=10e6, max_fo=5000.0, seed=15, cn0=57.0, chunks=32),
    # test_max_fo_changes_the_doppler_grid: one satellite outside the default +-5 kHz span, searched with max_fo = 10 kHz
    "maxfo_10000": dict(fs=5.456e6, fc=4.092e6, max_fo=10000.0, seed=5, cn0=None, chunks=9,
                        sats=[dict(prn=9, doppler_hz=7300.0, code_phase_chips=100.25, amp=0.3)]),
}


def case_bits(name: str) -> np.ndarray:
    import gpsacq_loader
    gpsacq_loader.load()
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    c = CASES[name]
    sats = c.get("sats") or sg.default_constellation(c["fs"], cn0_dbhz=c["cn0"], seed=c["seed"])
    return sg.synth_capture(40960 * c["chunks"], c["fs"], c["fc"], sats, seed=c["seed"])


def harness_job(name: str):
    """Runs in a subprocess: the reference keeps its state in file statics (one configuration per process)."""
    import oracle
    c = CASES[name]
    bits = case_bits(name)
    r = oracle.RefHarness(c["fc"], c["fs"], c["max_fo"])
    pk = r.search_blocks(bits.tobytes())
    np.save(HERE / f"_tmp_{name}.npy", pk)
    print(json.dumps({"sha256": hashlib.sha256(bits.tobytes()).hexdigest(), "backend": r.fft_backend,
                      "detected": int((pk["snr"] >= 25).sum())}))


def main():
    import oracle
    if len(sys.argv) > 2 and sys.argv[1] == "--job":
        return harness_job(sys.argv[2])
    out, meta = {}, {}
    for name in CASES:
        r = subprocess.run([sys.executable, __file__, "--job", name], check=True, capture_output=True, text=True, env=oracle.mkl_env())
        meta[name] = json.loads(r.stdout.strip().split("\n")[-1])
        tmp = HERE / f"_tmp_{name}.npy"
        pk = np.load(tmp)
        tmp.unlink()
        out[name] = np.rec.fromarrays([pk["snr"].astype(np.float32), pk["lo_shift"].astype(np.int32), pk["ca_shift"].astype(np.int32)],
                                      names="snr,lo_shift,ca_shift")
        print(name, meta[name])
    np.savez_compressed(HERE / "ref_peaks_rates.npz", **out)
    (HERE / "ref_peaks_rates.json").write_text(json.dumps({"cases": CASES, "made_by": "tests/golden/make_golden_rates.py", "inputs": meta}, indent=1))


if __name__ == "__main__":
    main()

"""Shared pytest plumbing.  `-m "not gpu"` runs in the GPU-less development container;
`-m gpu` runs on a B200 and calls the CUDA engine through the C ABI."""
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

CAPTURES = {
    # name: (fixture file, FC, FS, golden stdout, reference peaks, reference spectral probes)
    "nottingham": dict(bin=GOLD / "nottingham_fs5456_if4092_runs0-3.bin", fc=4.092e6, fs=5.456e6, runs=4,
                       stdout=GOLD / "nottingham_full.stdout.txt", peaks=GOLD / "ref_peaks_nottingham.npy",
                       probe=GOLD / "ref_probe_nottingham.npz"),
    "gps_sig": dict(bin=GOLD / "gps_sig_fs8184_if2046_runs0-1.bin", fc=2.046e6, fs=8.184e6, runs=2,
                    stdout=GOLD / "gps_sig_full.stdout.txt", peaks=GOLD / "ref_peaks_gps_sig.npy",
                    probe=GOLD / "ref_probe_gps_sig.npz"),
}

# SNR threshold band inside which a hit may legitimately flip between implementations
# (SURVEY.md App. A6): compared by SNR tolerance only.
MARGINAL = 1e-3
SNR_RTOL = 1e-4          # north_star: 1e-4 relative on the correlation peak


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def ga():
    import gpsacq_loader
    return gpsacq_loader.load()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


BANNER_LINES = 6


def parse_stdout(text: str):
    """Parse gps_test output (after the banner) into a list of runs:
    dict(sv=[...], snr=[...], lo=[...], ca=[...], all_snr=[32 rounded ints]); plus the tail text."""
    lines = text.split("\n")
    runs, i, tail = [], 0, []
    while i < len(lines):
        m = re.match(r"^\s*(\d+) satellite: (.*)$", lines[i])
        if not m:
            if lines[i].strip():
                tail.append(lines[i])
            i += 1
            continue
        r = int(m.group(1))
        sv = [int(v) for v in m.group(2).split()]
        snr = [float(v) for v in lines[i + 1].split(":")[1].split()]
        lo = [int(v) for v in lines[i + 2].split(":")[1].split()]
        ca = [int(v) for v in lines[i + 3].split(":")[1].split()]
        alls = [float(v) for v in lines[i + 4].split()]
        runs.append(dict(run=r, sv=sv, snr=snr, lo=lo, ca=ca, all_snr=alls))
        i += 6
    return runs, tail


def strip_banner(text: str) -> str:
    return "\n".join(text.split("\n")[BANNER_LINES:])


def compare_peaks(got, ref, snr_rtol=SNR_RTOL):
    """got/ref: structured arrays with snr, lo_shift, ca_shift.  Integer fields must be equal for
    every chunk that is a clear detection in the reference; SNR within snr_rtol everywhere a
    signal is present; noise-only chunks are compared through SNR only (their argmax is a
    coin toss between near-equal noise peaks in any two float implementations)."""
    got_snr, ref_snr = got["snr"].astype(np.float64), ref["snr"].astype(np.float64)
    clear = ref_snr >= 25.0 * (1 + MARGINAL)
    assert clear.any()
    assert np.array_equal(got["lo_shift"][clear], ref["lo_shift"][clear]), "Doppler bin differs on a detected PRN"
    assert np.array_equal(got["ca_shift"][clear], ref["ca_shift"][clear]), "code phase differs on a detected PRN"
    rel = np.abs(got_snr - ref_snr) / ref_snr
    assert rel[clear].max() <= snr_rtol, f"SNR of detected PRNs differs by {rel[clear].max():.2e}"
    # undetected chunks: same cell wins unless two noise cells are within rounding of each other
    same = (got["lo_shift"] == ref["lo_shift"]) & (got["ca_shift"] == ref["ca_shift"])
    assert rel[same].max() <= snr_rtol
    assert rel.max() <= 1e-3, "a noise-only chunk's best SNR moved by more than 1e-3"
    # detection decision identical outside the marginal band
    decided = np.abs(ref_snr / 25.0 - 1) > MARGINAL
    assert np.array_equal((got_snr >= 25.0)[decided], (ref_snr >= 25.0)[decided])
    return float(rel.max())


def compare_runs(got_runs, ref_runs):
    """Compare parsed stdout run lists (see parse_stdout) with the print-rounding tolerance policy."""
    assert len(got_runs) == len(ref_runs)
    for g, r in zip(got_runs, ref_runs):
        assert g["run"] == r["run"]
        ref_hits = {sv: (s, lo, ca) for sv, s, lo, ca in zip(r["sv"], r["snr"], r["lo"], r["ca"])}
        got_hits = {sv: (s, lo, ca) for sv, s, lo, ca in zip(g["sv"], g["snr"], g["lo"], g["ca"])}
        for sv in set(ref_hits) | set(got_hits):
            if sv in ref_hits and sv in got_hits:
                rs, gs = ref_hits[sv], got_hits[sv]
                assert gs[1:] == rs[1:], f"run {r['run']} sv {sv}: {gs} vs {rs}"
                assert abs(gs[0] - rs[0]) <= 0.1 + 1e-4 * rs[0]
            else:  # only allowed when hugging the threshold (printed 25.0 either way)
                s = (ref_hits.get(sv) or got_hits.get(sv))[0]
                assert s <= 25.0 + 0.05, f"run {r['run']} sv {sv} detected on one side only with SNR {s}"
        assert len(g["all_snr"]) == 32
        assert max(abs(a - b) for a, b in zip(g["all_snr"], r["all_snr"])) <= 1.0


# ---- the receiver's own search loop + hand-off (tests/golden/ref_target_events.json, made from the UNMODIFIED
# c/search.cpp + c/channel.cpp by tests/golden/make_golden_target.py) ---------------------------------------------------
def target_golden():
    import hashlib
    import importlib
    import json
    g = json.loads((GOLD / "ref_target_events.json").read_text())
    import gpsacq_loader
    gpsacq_loader.load()
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    i = g["input"]
    sats = sg.default_constellation(i["fs"], cn0_dbhz=i["cn0_dbhz"], seed=i["seed_constellation"])
    bits = sg.synth_capture(40960 * i["n_chunks"], i["fs"], i["fc"], sats, seed=i["seed_noise"])
    assert hashlib.sha256(bits.tobytes()).hexdigest() == i["sha256"], "the numpy generator no longer reproduces the golden's input"
    return g, bits


def replay_target(g, bits, feed_chunk, signal_lost):
    """Feed the stream chunk by chunk; after chunk c apply the losses the reference logged while chunk c was its most
    recently sampled one (CHANNEL::SignalLost(): channel freed, SearchEnable(sv)).  feed_chunk(bytes) -> list of
    (sv, ch, lo_shift, ca_shift, record) for the detections of that chunk; signal_lost(ch, sv).
    Returns [(chunk, sv, ch, lo_shift, ca_shift, record)], to be compared with the golden's start events."""
    lost_after = {}
    for e in g["events"]:
        if e["type"] == "lost":
            lost_after.setdefault(e["chunk"], []).append(e)
    got = []
    for c in range(g["input"]["n_chunks"]):
        for (sv, ch, lo, ca, rec) in feed_chunk(bits[c * 5120:(c + 1) * 5120].tobytes()):
            got.append((c, sv, ch, lo, ca, rec))
        for e in lost_after.get(c, []):
            signal_lost(e["ch"], e["sv"])
    return got


# ---- the unmodified reference's peaks at sampling rates without a bundled capture (tests/golden/ref_peaks_rates.npz, made
# by tests/golden/make_golden_rates.py from synthetic captures that are regenerated, not stored) -------------------------
def rates_golden(name: str, bits: np.ndarray) -> np.ndarray:
    """Reference records of case `name`; `bits` = the regenerated input, checked against the SHA-256 recorded with the golden."""
    import hashlib
    import json
    meta = json.loads((GOLD / "ref_peaks_rates.json").read_text())
    assert hashlib.sha256(np.ascontiguousarray(bits).tobytes()).hexdigest() == meta["inputs"][name]["sha256"], \
        f"the numpy generator no longer reproduces the input of golden case {name}"
    return np.load(GOLD / "ref_peaks_rates.npz")[name]


def rates_case(fs: float, fc: float, seed: int) -> str:
    import json
    for name, c in json.loads((GOLD / "ref_peaks_rates.json").read_text())["cases"].items():
        if (c["fs"], c["fc"], c["seed"]) == (fs, fc, seed):
            return name
    raise KeyError((fs, fc, seed))

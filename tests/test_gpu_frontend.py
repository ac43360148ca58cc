"""GPU tests of the 8-bit IQ front-end: the MATLAB pre-processing of the reference's
SDR workflows, done on the device, then searched with the REF engine."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 1], ids=["v1", "v2"])
def frontend_version(request, monkeypatch):
    """Every test of this file runs through both kernel generations of the converters: the double-precision table kernel /
    compare-select expander, and the threshold-table / byte-permute kernels (csrc/ga_frontend.cuh; the library reads
    GPSACQ_FRONTEND_V2 on every call)."""
    monkeypatch.setenv("GPSACQ_FRONTEND_V2", str(request.param))
    return request.param


def matlab_restatement(iq_u8: np.ndarray, fc: float, fs: float, signed: bool) -> np.ndarray:
    """proc_rtl_bin_for_gps.m:31-47 (uint8) / proc_hackrf_bin_for_gps.m:7-19 (int8), line by line, in double.
    PARITY UNPINNED: neither MATLAB nor Octave is available and the reference bundles no output of these scripts, so this
    restatement is checked against nothing but its source text."""
    y = iq_u8.view(np.int8).astype(np.float64) if signed else iq_u8.astype(np.float64) - 128      # :34  y = y - 128
    y = y[0::2] + 1j * y[1::2]                                                                     # :35
    y = y - y.mean()                                                                               # :36
    n = np.arange(y.size, dtype=np.float64)
    r = np.real(y * np.exp(1j * (((2.0 * np.pi * fc) * n) * (1.0 / fs))))                          # :41-42
    bits = (1 - np.sign(r)) / 2                                                                    # :44
    return np.packbits((bits > 0.5).astype(np.uint8)[: y.size // 8 * 8].reshape(-1, 8), axis=1, bitorder="little").reshape(-1)


@pytest.mark.parametrize("signed", [False, True])
def test_iq8_frontend_then_search(ga, oracle_mod, signed):
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    fs, fc = 2.8e6, 0.62e6                                    # README.md:69-85 rtl-sdr workflow
    sats = sg.default_constellation(fs, cn0_dbhz=47.0, seed=5)
    n = 40960 * 32
    iq = sg.synth_iq8(n, fs, sats, seed=9, signed=signed)
    acq = ga.Acquisition(fc, fs)
    try:
        bits = acq.iq8_to_bits(iq, fc, fs, signed=signed)
        ref_bits = matlab_restatement(iq, fc, fs, signed)
        assert bits.size == ref_bits.size == n // 8
        flips = int(np.unpackbits(bits ^ ref_bits).sum())
        assert flips <= 2, f"{flips} of {n} samples differ from the MATLAB restatement"   # only |r| ~ 1e-13 cases may flip
        got = acq.search_blocks(bits)
        for s in sats:
            p = got[s["prn"] - 1]
            assert p["snr"] >= 25 and abs(p["lo_shift"] - s["doppler_hz"] * 40000 / fs) <= 1.0
        # and the whole chain equals the CPU oracle run on the restated bits
        ref = oracle_mod.Oracle(fc, fs).search_blocks(ref_bits)
        det = ref["snr"] >= 26
        assert np.array_equal(got["lo_shift"][det], ref["lo_shift"][det]) and np.array_equal(got["ca_shift"][det], ref["ca_shift"][det])
    finally:
        acq.close()


def test_iq8_frontend_edges(ga):
    acq = ga.Acquisition(0.62e6, 2.8e6)
    try:
        assert acq.iq8_to_bits(np.zeros(0, np.uint8), 0.62e6).size == 0
        # constant input: after mean removal everything is exactly 0 -> sign(0) = 0 -> bit 0 (never negative)
        assert not acq.iq8_to_bits(np.full(2 * 64, 131, np.uint8), 0.62e6).any()
        # 13 samples -> 2 bytes, upper bits of the last byte stay 0
        out = acq.iq8_to_bits(np.arange(26, dtype=np.uint8) * 9 % 251, 0.0)
        assert out.size == 2 and out[1] < 32
    finally:
        acq.close()


@pytest.mark.parametrize("fc,fs,signed,n", [(0.62e6, 2.8e6, False, 40960 * 24 + 5), (2.6e6, 10e6, True, 1_000_003),
                                            (4.092e6, 5.456e6, False, 123_457), (0.0, 2.8e6, True, 65_536)])
def test_iq8_threshold_kernel_equals_double_kernel(ga, monkeypatch, fc, fs, signed, n):
    """Threshold-table pass 2 == the double-precision pass 2, bit for bit (the table is made by evaluating the same
    expression; tests/test_frontend_emu.py shows it on the CPU), including samples hugging the mean and the tail."""
    rng = np.random.default_rng(n)
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    iq[: 2 * 50_000] = np.clip(rng.normal(128, 2.5, 2 * 50_000), 0, 255).astype(np.uint8)
    acq = ga.Acquisition(0.62e6, 2.8e6)
    try:
        monkeypatch.setenv("GPSACQ_FRONTEND_V2", "0")
        a = acq.iq8_to_bits(iq, fc, fs, signed=signed)
        monkeypatch.setenv("GPSACQ_FRONTEND_V2", "1")
        b = acq.iq8_to_bits(iq, fc, fs, signed=signed)
        c = acq.iq8_to_bits(iq, -fc, fs, signed=signed)          # the other shift direction rebuilds both tables
    finally:
        acq.close()
    assert a.size == (n + 7) // 8 and np.array_equal(a, b)
    assert a.any() and (fc == 0.0 or not np.array_equal(b, c))


# ---- the reverse converter: 1-bit IF -> int8 IQ (c/conv_1bit_bin_to_hackrf_bin.cpp:29-86), pinned by the reference program ----
def test_bits_to_iq8_vs_the_reference_program(ga):
    """Bit-exact against the UNMODIFIED reference converter: SHA-256 of its output for the prefix of the capture that is
    the committed fixture, and -- when the capture travelled -- for the whole 55.8 MB file (892,665,856 output bytes)."""
    import hashlib, json
    from conftest import GOLD, ROOT
    g = json.loads((GOLD / "f1f2_golden.json").read_text())["conv_1bit_bin_to_hackrf_bin"]
    out = ga.bits_to_iq8((GOLD / g["fixture"]).read_bytes(), g["fc"], g["fs"], g["amplitude"])
    assert out[:32].tolist() == g["first_32_out_bytes"]
    assert hashlib.sha256(out.tobytes()).hexdigest() == g["sha256_fixture_prefix"]
    full = ROOT / "oracle" / "_ref" / "data" / "gps.samples.1bit.I.fs5456.if4092.bin"
    if full.exists():
        raw = np.fromfile(full, np.uint8)
        out = ga.bits_to_iq8(raw, g["fc"], g["fs"], g["amplitude"])
        assert out.size == g["n_out_bytes"] and hashlib.sha256(out.tobytes()).hexdigest() == g["sha256_full"]


@pytest.mark.parametrize("fc,fs", [(4.092e6, 5.456e6), (0.62e6, 2.8e6), (4.1304e6, 16.368e6), (2.6e6, 10e6)])
def test_bits_to_iq8_vs_oracle_other_rates(ga, oracle_mod, fc, fs):
    """Other LO rates (exact: period 4; inexact float rates: periods 262,144 ... 4,194,305 samples -- longer than the
    input or wrapped several times), odd lengths, a piece that starts in the middle of the stream, other amplitudes."""
    rng = np.random.default_rng(int(fs) % 97)
    bits = rng.integers(0, 256, 1_500_001, dtype=np.uint8)            # 12 M samples
    want = oracle_mod.conv_1bit_iq8(bits, fc, fs, 30)
    got = ga.bits_to_iq8(bits, fc, fs, 30)
    assert np.array_equal(got, want)
    part = ga.bits_to_iq8(bits[700_001:], fc, fs, 127, first_sample=8 * 700_001)
    assert np.array_equal(part, (want[16 * 700_001:].astype(np.int16) * 127 // 30).astype(np.int8))
    assert ga.bits_to_iq8(b"", fc, fs).size == 0
    with pytest.raises(ga.GpsAcqError):
        ga.bits_to_iq8(bits[:8], 6e6, 5e6)                               # 4*fc/fs >= 4: int(phase) would leave the LO tables


def test_bits_to_iq8_round_trip_through_the_front_end(ga):
    """1-bit IF -> int8 IQ at baseband (reverse converter) -> the 8-bit front-end shifting back up by the same IF
    reproduces the sign of the original samples wherever the mixed value is not zero: an end-to-end consistency check
    of the two converters' sign and phase conventions."""
    fc, fs = 2.046e6, 8.184e6                                            # LO rate exactly 1.0: phase index = n mod 4
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 256, 40960, dtype=np.uint8)
    iq = ga.bits_to_iq8(bits, fc, fs, 30)
    acq = ga.Acquisition(fc, fs)
    try:
        back = acq.iq8_to_bits(iq.view(np.uint8), fc, fs, signed=True)
    finally:
        acq.close()
    a = np.unpackbits(bits, bitorder="little").astype(int)
    b = np.unpackbits(back, bitorder="little").astype(int)
    # I = s*lo_sin', Q = s*lo_cos' (bipolar s = +-1): re((I + jQ - mean) e^{j pi n/2}) has the sign of +-s in a fixed
    # 4-sample pattern; recover the pattern from the first 4 samples and require it to hold everywhere
    flip = a[:4] ^ b[:4]
    assert np.array_equal(a ^ b, np.tile(flip, a.size // 4))


def test_conv_tool_binary(tmp_path):
    """The C++ drop-in of the reference's converter program: same default file names, packet rule and messages
    (c/conv_1bit_bin_to_hackrf_bin.cpp:25,41-59,91); on the whole capture its output file has the reference's SHA-256."""
    import hashlib, json, os, subprocess
    from conftest import GOLD, ROOT
    subprocess.run(["make", "-s", "gps_test"], cwd=ROOT, check=True)
    exe = ROOT / "gnss-gps-sdr_b200" / "c" / "conv_1bit_bin_to_hackrf_bin"
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "can not open file for read!\n"
    (tmp_path / "gps.samples.1bit.I.fs5456.if4092.bin").write_bytes(bytes(1000))     # shorter than one packet
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "seems run out!\n"
    assert (tmp_path / "gps.samples.8bit.IQinterleave.fs5456.if0.bin").stat().st_size == 0
    full = ROOT / "oracle" / "_ref" / "data" / "gps.samples.1bit.I.fs5456.if4092.bin"
    if full.exists():
        g = json.loads((GOLD / "f1f2_golden.json").read_text())["conv_1bit_bin_to_hackrf_bin"]
        os.remove(tmp_path / "gps.samples.1bit.I.fs5456.if4092.bin")
        os.symlink(full, tmp_path / "gps.samples.1bit.I.fs5456.if4092.bin")
        r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and r.stdout == "1\nseems run out!\n", r.stderr
        h = hashlib.sha256()
        with open(tmp_path / "gps.samples.8bit.IQinterleave.fs5456.if0.bin", "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        assert h.hexdigest() == g["sha256_full"]

"""GPU tests of the 8-bit IQ front-end: the MATLAB pre-processing of the reference's
SDR workflows, done on the device, then searched with the REF engine."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def matlab_restatement(iq_u8: np.ndarray, fc: float, fs: float, signed: bool) -> np.ndarray:
    """proc_rtl_bin_for_gps.m:31-47 (uint8) / proc_hackrf_bin_for_gps.m:7-19 (int8), line by line, in double."""
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

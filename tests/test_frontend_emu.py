"""CPU-only: the thread functions of the second-version stream converters (csrc/ga_frontend_math.h, run by
iq8_to_bits_thr_kernel and bits_to_iq8_v2_kernel) replayed thread by thread in tests/emu, against
 * the double expression itself (threshold table == sign of r, sample for sample: the table is exact, not approximate),
 * a numpy restatement of proc_rtl_bin_for_gps.m:31-47 / proc_hackrf_bin_for_gps.m:7-19,
 * a plain restatement of c/conv_1bit_bin_to_hackrf_bin.cpp:62-80 over an arbitrary LO cycle (pre-period, wrap-around),
   and the oracle's restatement of the whole program for real LO rates."""
import ctypes
import subprocess

import numpy as np
import pytest

from conftest import ROOT

C = ctypes
U8P = C.POINTER(C.c_ubyte)


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "emu"], cwd=ROOT, check=True)
    L = C.CDLL(str(ROOT / "tests/emu/libemu.so"))
    L.emu_iq8_thr.argtypes = [U8P, C.c_size_t, C.c_size_t, C.c_int, C.c_longlong, C.c_longlong, C.c_size_t,
                              C.POINTER(C.c_double), C.c_uint, C.c_uint, C.c_uint, U8P]
    L.emu_iq8_direct.argtypes = [U8P, C.c_size_t, C.c_size_t, C.c_int, C.c_longlong, C.c_longlong, C.c_size_t,
                                 C.POINTER(C.c_double), C.c_uint, C.c_uint, U8P]
    L.emu_conv_v2.argtypes = [U8P, C.c_size_t, C.c_size_t, U8P, C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_size_t, U8P]
    return L


def p8(a):
    return a.ctypes.data_as(U8P)


def phasor_table(p, q, neg=False):
    """What iq8_convert_piece uploads: (cos, +-sin)(2 pi k / q) in long double, rounded once."""
    k = np.arange(q, dtype=np.longdouble)
    a = np.longdouble(2) * np.longdouble("3.14159265358979323846264338327950288") * k / np.longdouble(q)
    t = np.empty(2 * q, np.float64)
    t[0::2] = np.cos(a).astype(np.float64)
    t[1::2] = (-np.sin(a) if neg else np.sin(a)).astype(np.float64)
    return t


def sums_of(iq, signed):
    v = iq.view(np.int8).astype(np.int64) if signed else iq.astype(np.int64) - 128
    return int(v[0::2].sum()), int(v[1::2].sum())


@pytest.mark.parametrize("p,q,signed,n,n0,threads", [
    (31, 140, False, 40960, 0, 148 * 1024),       # rtl-sdr: 0.62 / 2.8 MHz
    (31, 140, True, 40960 + 13, 0, 1024),         # tail of 5 samples, many trips per thread
    (13, 50, True, 30000, 1 << 26, 4096),         # HackRF: 2.6 / 10 MHz; a later piece of a long capture
    (3, 4, False, 8192, 8 * 77, 512),             # fs/4-type ratios
    (0, 1, False, 4099, 0, 256),                  # no shift at all
    (113, 227, False, 20000, 0, 2048),            # the largest table that fits
    (1, 2, True, 999, 8, 64),
])
def test_threshold_table_equals_the_double_expression(emu, p, q, signed, n, n0, threads):
    rng = np.random.default_rng(p * 1000 + q)
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    if q == 140 and not signed:
        iq[: 2 * 4000] = np.clip(rng.normal(128, 3, 2 * 4000), 0, 255).astype(np.uint8)     # samples hugging the mean
    si, sq = sums_of(iq, signed)
    tab = phasor_table(p, q, neg=(q == 50))
    a = np.zeros((n + 7) // 8, np.uint8)
    b = np.zeros_like(a)
    assert emu.emu_iq8_thr(p8(iq), n, n0, int(signed), si, sq, n, tab.ctypes.data_as(C.POINTER(C.c_double)), p, q, threads, p8(a)) == 0
    assert emu.emu_iq8_direct(p8(iq), n, n0, int(signed), si, sq, n, tab.ctypes.data_as(C.POINTER(C.c_double)), p, q, p8(b)) == 0
    assert np.array_equal(a, b)
    assert a.any() and not a.all()


def test_threshold_table_exact_zero_and_constant_input(emu):
    """Constant input: every (I - mean, Q - mean) is exactly 0, r = 0, sign(0) = 0 -> bit 0 (never 'negative');
    and a capture whose mean is an integer, so that r = 0 happens for real samples."""
    n, p, q = 4096, 31, 140
    tab = phasor_table(p, q)
    iq = np.full(2 * n, 131, np.uint8)
    out = np.ones(n // 8, np.uint8)
    assert emu.emu_iq8_thr(p8(iq), n, 0, 0, 3 * n, 3 * n, n, tab.ctypes.data_as(C.POINTER(C.c_double)), p, q, 1024, p8(out)) == 0
    assert not out.any()
    rng = np.random.default_rng(2)
    iq = rng.integers(120, 137, 2 * n, dtype=np.uint8)
    a, b = np.zeros(n // 8, np.uint8), np.zeros(n // 8, np.uint8)
    # pretend sums that make both means exactly 0: I = 128 / Q = 128 samples then give r = -yq*sin or yi*cos = +-0
    for fn, o in ((emu.emu_iq8_thr, a), (emu.emu_iq8_direct, b)):
        args = [p8(iq), n, 0, 0, 0, 0, n, tab.ctypes.data_as(C.POINTER(C.c_double)), p, q]
        assert fn(*(args + ([1024] if fn is emu.emu_iq8_thr else []) + [p8(o)])) == 0
    assert np.array_equal(a, b)


@pytest.mark.parametrize("signed", [False, True])
def test_threshold_path_vs_matlab_restatement(emu, signed):
    """The whole conversion against proc_rtl_bin_for_gps.m / proc_hackrf_bin_for_gps.m restated in numpy (phase from
    2*pi*fc*n/fs in double, as MATLAB evaluates it): only samples with |r| ~ 1e-13 may differ."""
    fc, fs, p, q = 0.62e6, 2.8e6, 31, 140
    n = 40960 * 4
    rng = np.random.default_rng(11 + signed)
    iq = np.clip(rng.normal(128, 20, 2 * n), 0, 255).astype(np.uint8)
    y = iq.view(np.int8).astype(np.float64) if signed else iq.astype(np.float64) - 128
    y = y[0::2] + 1j * y[1::2]
    y = y - y.mean()
    r = np.real(y * np.exp(1j * (((2.0 * np.pi * fc) * np.arange(n, dtype=np.float64)) * (1.0 / fs))))
    want = np.packbits((r < 0).astype(np.uint8).reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
    si, sq = sums_of(iq, signed)
    got = np.zeros(n // 8, np.uint8)
    tab = phasor_table(p, q)
    assert emu.emu_iq8_thr(p8(iq), n, 0, int(signed), si, sq, n, tab.ctypes.data_as(C.POINTER(C.c_double)), p, q, 148 * 1024, p8(got)) == 0
    assert int(np.unpackbits(got ^ want).sum()) <= 2


def conv_plain(bits, first_sample, lo, mu, lam, amp):
    """c/conv_1bit_bin_to_hackrf_bin.cpp:62-80 with the phase index taken from a table of one pre-period + one period."""
    b = np.unpackbits(bits, bitorder="little").astype(np.int64)
    i = first_sample + np.arange(b.size, dtype=np.int64)
    k = np.where(i < mu, i, mu + (i - mu) % lam)
    code = lo[k].astype(np.int64)
    out = np.empty(2 * b.size, np.int8)
    out[0::2] = amp * (1 - 2 * (b ^ (code & 1)))
    out[1::2] = amp * (1 - 2 * (b ^ (code >> 1)))
    return out


@pytest.mark.parametrize("mu,lam,first,nbytes,threads,amp", [
    (0, 4, 0, 5000, 256, 30),                 # exact LO rates: period 4
    (0, 4, 8 * 123, 777, 64, 127),
    (37, 1001, 0, 4000, 96, 30),              # pre-period, odd period: groups straddle the wrap, threads cross mu mid-loop
    (37, 1001, 8 * 3, 4000, 1, 1),            # one thread walks the whole stream
    (1000, 13, 0, 600, 32, 100),              # period shorter than two groups
    (5, 262144, 8 * 40000, 3000, 128, 30),    # long period, start deep inside
    (0, 50000, 0, 5000, 8192, 0),             # more threads than bytes; amplitude 0
])
def test_conv_v2_thread_function(emu, mu, lam, first, nbytes, threads, amp):
    rng = np.random.default_rng(mu + lam)
    lo = np.zeros(mu + lam + 16, np.uint8)
    lo[: mu + lam] = rng.integers(0, 4, mu + lam, dtype=np.uint8)
    bits = rng.integers(0, 256, nbytes, dtype=np.uint8)
    out = np.zeros(16 * nbytes, np.uint8)
    assert emu.emu_conv_v2(p8(bits), nbytes, first, p8(lo), mu, lam, amp, threads, p8(out)) == 0
    assert np.array_equal(out.view(np.int8), conv_plain(bits, first, lo, mu, lam, amp))


@pytest.mark.parametrize("fc,fs", [(4.092e6, 5.456e6), (2.6e6, 10e6)])
def test_conv_v2_vs_oracle_restatement(emu, oracle_mod, fc, fs):
    """Real LO rates: the table is the float recurrence of :33,:79-80 itself (no cycle search: mu = 0, lambda = length)."""
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 256, 20000, dtype=np.uint8)
    want = oracle_mod.conv_1bit_iq8(bits, fc, fs, 30)
    n = 8 * bits.size
    rate, ph = np.float32(4 * fc / fs), np.float32(0)
    lo_sin, lo_cos = [1, 1, 0, 0], [1, 0, 0, 1]
    lo = np.zeros(n + 16, np.uint8)
    for i in range(n):
        k = int(ph)
        lo[i] = lo_sin[k] | (lo_cos[k] << 1)
        ph = np.float32(ph + rate)
        if ph >= 4:
            ph = np.float32(ph - np.float32(4))
    out = np.zeros(16 * bits.size, np.uint8)
    assert emu.emu_conv_v2(p8(bits), bits.size, 0, p8(lo), 0, n, 30, 777, p8(out)) == 0
    assert np.array_equal(out.view(np.int8), want)


# ---- randomised sweeps (hypothesis): shapes the hand-picked cases above do not name --------------------------------
from hypothesis import given, settings, strategies as st
from math import gcd


@settings(max_examples=60, deadline=None)
@given(q=st.integers(1, 227), pnum=st.integers(0, 226), signed=st.booleans(), n=st.integers(1, 3000), blk=st.integers(0, 1 << 20),
       threads=st.sampled_from([32, 96, 1024, 5000]), spread=st.sampled_from([1.5, 20.0, 200.0]), seed=st.integers(0, 2 ** 31))
def test_threshold_table_equals_the_double_expression_random(emu, q, pnum, signed, n, blk, threads, spread, seed):
    p = pnum % q
    g = gcd(p, q) if p else q
    p, q = (p // g, q // g) if p else (0, 1)                 # lowest terms, like small_rational()
    if threads < q:
        threads = q
    rng = np.random.default_rng(seed)
    iq = np.clip(rng.normal(128 + rng.uniform(-3, 3), spread, 2 * n), 0, 255).astype(np.uint8)
    si, sq = sums_of(iq, signed)
    tab = phasor_table(p, q, neg=bool(seed & 1))
    a = np.zeros((n + 7) // 8, np.uint8)
    b = np.zeros_like(a)
    dp = tab.ctypes.data_as(C.POINTER(C.c_double))
    assert emu.emu_iq8_thr(p8(iq), n, 8 * blk, int(signed), si, sq, n, dp, p, q, threads, p8(a)) == 0
    assert emu.emu_iq8_direct(p8(iq), n, 8 * blk, int(signed), si, sq, n, dp, p, q, p8(b)) == 0
    assert np.array_equal(a, b)


@settings(max_examples=60, deadline=None)
@given(mu=st.integers(0, 300), lam=st.integers(1, 5000), first=st.integers(0, 100000), nbytes=st.integers(1, 1500),
       threads=st.sampled_from([1, 7, 32, 256, 4096]), amp=st.integers(0, 127), seed=st.integers(0, 2 ** 31))
def test_conv_v2_thread_function_random(emu, mu, lam, first, nbytes, threads, amp, seed):
    rng = np.random.default_rng(seed)
    lo = np.zeros(mu + lam + 16, np.uint8)
    lo[: mu + lam] = rng.integers(0, 4, mu + lam, dtype=np.uint8)
    bits = rng.integers(0, 256, nbytes, dtype=np.uint8)
    out = np.zeros(16 * nbytes, np.uint8)
    assert emu.emu_conv_v2(p8(bits), nbytes, first, p8(lo), mu, lam, amp, threads, p8(out)) == 0
    assert np.array_equal(out.view(np.int8), conv_plain(bits, first, lo, mu, lam, amp))


# ---- host-side table preparation (csrc/ga_frontend_host.h) ----------------------------------------------------------------
@pytest.mark.parametrize("num,den,want", [(0.62e6, 2.8e6, (31, 140)), (2.6e6, 10e6, (13, 50)), (4.092e6, 5.456e6, (3, 4)),
                                          (2.046e6, 8.184e6, (1, 4)), (0.0, 2.8e6, (0, 1)), (1e6 / 3, 2.8e6, (5, 42)),
                                          (3.5e6, 2.8e6, (1, 4)), (4.1304e6, 16.368e6, (1721, 6820))])
def test_small_rational(emu, num, den, want):
    """shift_hz / fs as a fraction in lowest terms, reduced modulo one turn (p < q)."""
    from fractions import Fraction
    p, q = C.c_ulonglong(), C.c_ulonglong()
    emu.emu_small_rational.argtypes = [C.c_double, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    assert emu.emu_small_rational(num / den, C.byref(p), C.byref(q)) == 1
    assert (p.value, q.value) == want
    f = Fraction(num / den).limit_denominator(1 << 20)
    assert (f.numerator % f.denominator, f.denominator) == want


def test_small_rational_rejects_what_is_not_one(emu):
    p, q = C.c_ulonglong(), C.c_ulonglong()
    emu.emu_small_rational.argtypes = [C.c_double, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    for x in (np.pi / 10, 123456.789 / 2.8e6, float("nan"), -0.25, 1e7):
        assert emu.emu_small_rational(x, C.byref(p), C.byref(q)) == 0


@pytest.mark.parametrize("fc,fs,first,n", [(4.092e6, 5.456e6, 0, 100_000), (2.046e6, 8.184e6, 12_345, 100_000),
                                           (2.6e6, 10e6, 0, 3_000_000), (2.6e6, 10e6, 9_000_000, 1_000_000),
                                           (0.62e6, 2.8e6, 0, 2_000_000), (4.1304e6, 16.368e6, 5_000_001, 1_000_000)])
def test_conv_lo_cycle_table_is_the_float_recurrence(emu, fc, fs, first, n):
    """Brent's cycle search + (pre-period, period) table == the float phase NCO of :33,:79-80 run sample by sample, also
    far beyond one period; the sample-by-sample loop itself is pinned to numpy float32 arithmetic on its first samples."""
    emu.emu_conv_lo_cycle.argtypes = [C.c_double, C.c_double, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong),
                                      C.POINTER(C.c_ulonglong), U8P, U8P]
    mu, lam = C.c_ulonglong(), C.c_ulonglong()
    a, b = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    assert emu.emu_conv_lo_cycle(fc, fs, first, n, C.byref(mu), C.byref(lam), p8(a), p8(b)) == 1
    assert np.array_equal(a, b)
    assert lam.value >= 1 and mu.value + lam.value <= (1 << 28)
    if first == 0:
        rate, ph = np.float32(4 * fc / fs), np.float32(0)
        lo_sin, lo_cos = [1, 1, 0, 0], [1, 0, 0, 1]
        for i in range(20_000):
            k = int(ph)
            assert b[i] == lo_sin[k] | (lo_cos[k] << 1)
            ph = np.float32(ph + rate)
            if ph >= 4:
                ph = np.float32(ph - np.float32(4))

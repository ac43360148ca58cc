"""Synthetic 1-bit GPS L1 C/A IF capture generator (host side, numpy).

Restates what the reference's MATLAB tooling does -- ``cacode.m`` (C/A chips, PRN 1-37,
cacode.m:65-120) and ``gps_sig_gen.m`` (code x NAV bits -> IF carrier -> sign -> ``ubit1``
LSB-first file, gps_sig_gen.m:8-41) -- generalised to several satellites with Doppler, code
phase and noise so that bench.py and the tests have inputs of any sampling rate / length.
Signals are generated directly at the target rate with a code NCO (the x8 zero-stuff + FIR of
gps_sig_gen.m only works at 8.184 MHz and is not needed for acquisition inputs).
"""
from __future__ import annotations

import numpy as np

CPS = 1.023e6
# G2 tap pairs, PRN 1..37 (cacode.m:65-101; PRN 1-32 identical to c/search_offline.cpp:20-53)
_TAPS = [(2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6),
         (6, 7), (7, 8), (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6),
         (5, 7), (6, 8), (7, 9), (8, 10), (1, 6), (2, 7), (3, 8), (4, 9), (5, 10), (4, 10), (1, 7), (2, 8),
         (4, 10)]


def cacode(prn: int) -> np.ndarray:
    """1023 C/A chips (0/1) of PRN `prn` (1-based, like cacode.m)."""
    t0, t1 = _TAPS[prn - 1]
    g1 = np.ones(10, np.uint8)
    g2 = np.ones(10, np.uint8)
    out = np.empty(1023, np.uint8)
    for i in range(1023):
        out[i] = g1[9] ^ g2[t0 - 1] ^ g2[t1 - 1]
        f1 = g1[2] ^ g1[9]
        f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1[1:] = g1[:-1]; g1[0] = f1
        g2[1:] = g2[:-1]; g2[0] = f2
    return out


def pack_bits_lsb_first(bits01: np.ndarray) -> np.ndarray:
    """`fwrite(...,'ubit1')` order: sample i -> bit (i & 7) of byte i >> 3 (c/search_offline.cpp:143-146)."""
    n = bits01.size // 8 * 8
    return np.packbits(bits01[:n].astype(np.uint8).reshape(-1, 8), axis=1, bitorder="little").reshape(-1)


def synth_capture(n_samples: int, fs: float, fc: float, sats, seed: int = 1575420000,
                  noise_sigma: float = 1.0, nav_bps: float = 50.0, chunk: int = 1 << 22) -> np.ndarray:
    """Packed 1-bit real-IF capture.

    sats: iterable of dicts {prn, doppler_hz, code_phase_chips, amp}.  Sample n is
    sign(sum_k amp_k * nav_k(t) * ca_k(t) * cos(2 pi (fc + fd_k) t + ph_k) + noise), encoded
    bit = (1 - sign)/2 (gps_sig_gen.m:37), packed LSB first.
    """
    rng = np.random.default_rng(seed)
    sats = list(sats)
    codes = [1.0 - 2.0 * cacode(s["prn"]).astype(np.float64) for s in sats]
    phases = rng.uniform(0, 2 * np.pi, len(sats))
    nbits = int(np.ceil(n_samples / fs * nav_bps)) + 2
    nav = [1.0 - 2.0 * rng.integers(0, 2, nbits) for _ in sats]
    out = np.empty(n_samples // 8, np.uint8)
    for start in range(0, n_samples, chunk):
        n = min(chunk, n_samples - start)
        t = (start + np.arange(n, dtype=np.float64)) / fs
        x = noise_sigma * rng.standard_normal(n) if noise_sigma > 0 else np.zeros(n)
        for k, s in enumerate(sats):
            # code rate follows the carrier Doppler (fd/1540 chips/s), as in a real signal
            chip = (t * (CPS * (1.0 + s["doppler_hz"] / 1575.42e6)) + s["code_phase_chips"]) % 1023.0
            c = codes[k][chip.astype(np.int64)]
            d = nav[k][(t * nav_bps).astype(np.int64)]
            x += s["amp"] * d * c * np.cos(2 * np.pi * (fc + s["doppler_hz"]) * t + phases[k])
        bits = (x < 0).astype(np.uint8)
        out[start // 8:(start + n) // 8] = pack_bits_lsb_first(bits)
    return out


def default_constellation(fs: float, cn0_dbhz: float = 45.0, seed: int = 1575420000, max_doppler: float = 4500.0):
    """8 satellites {1,5,8,13,21,29,30,31} with random Doppler / code phase at a given C/N0
    (real IF, unit-variance noise => amp = sqrt(4*10^(CN0/10)/fs))."""
    rng = np.random.default_rng(seed + 7)
    amp = float(np.sqrt(4.0 * 10 ** (cn0_dbhz / 10.0) / fs))
    return [dict(prn=p, doppler_hz=float(rng.uniform(-max_doppler, max_doppler)),
                 code_phase_chips=float(rng.uniform(0, 1023)), amp=amp)
            for p in (1, 5, 8, 13, 21, 29, 30, 31)]


def synth_iq8(n_samples: int, fs: float, sats, seed: int = 7, noise_sigma: float = 20.0, dc=(3.3, -2.1),
              signed: bool = False, nav_bps: float = 50.0) -> np.ndarray:
    """Interleaved 8-bit IQ baseband capture as an rtl-sdr (uint8, offset 128) or a HackRF (int8) writes
    it: the receiver is tuned to L1, so satellite k sits at its Doppler.  A DC offset is added because
    real dongles have one -- that is what the `y - mean(y)` of proc_rtl_bin_for_gps.m:36 removes."""
    rng = np.random.default_rng(seed)
    sats = list(sats)
    t = np.arange(n_samples, dtype=np.float64) / fs
    y = noise_sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples)) + complex(*dc)
    nbits = int(np.ceil(n_samples / fs * nav_bps)) + 2
    for s in sats:
        code = 1.0 - 2.0 * cacode(s["prn"]).astype(np.float64)
        chip = (t * (CPS * (1.0 + s["doppler_hz"] / 1575.42e6)) + s["code_phase_chips"]) % 1023.0
        d = (1.0 - 2.0 * rng.integers(0, 2, nbits))[(t * nav_bps).astype(np.int64)]
        y += s["amp"] * noise_sigma * d * code[chip.astype(np.int64)] * np.exp(1j * (2 * np.pi * s["doppler_hz"] * t + rng.uniform(0, 2 * np.pi)))
    iq = np.empty(2 * n_samples, np.float64)
    iq[0::2], iq[1::2] = np.round(y.real), np.round(y.imag)
    iq = np.clip(iq, -127, 127)
    return iq.astype(np.int8).view(np.uint8) if signed else (iq + 128).astype(np.uint8)


# ---- counter-based restatement of the GPU generator (csrc/ga_siggen.cuh), for parity tests -----------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def synth_capture_counter(n_samples: int, fs: float, fc: float, sats, seed: int = 1, noise_sigma: float = 1.0,
                          nav_bps: float = 50.0) -> np.ndarray:
    """Same formulas, same splitmix64 counters as synth_bits_kernel: bit-identical except where |x| ~ 1e-15."""
    n = np.arange(n_samples, dtype=np.uint64)
    t = n.astype(np.float64) / fs
    seed64 = np.uint64(seed)
    x = np.zeros(n_samples)
    if noise_sigma > 0:
        a = _splitmix64(seed64 ^ _splitmix64(np.uint64(2) * n))
        b = _splitmix64(seed64 ^ _splitmix64(np.uint64(2) * n + np.uint64(1)))
        u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740993.0)
        u2 = (b >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        x += noise_sigma * np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    for k, s in enumerate(sats):
        code = 1.0 - 2.0 * cacode(s["prn"]).astype(np.float64)
        chip = np.fmod(t * (CPS * (1.0 + s["doppler_hz"] / 1575.42e6)) + s["code_phase_chips"], 1023.0).astype(np.int64)
        nb = (t * nav_bps).astype(np.uint64)
        with np.errstate(over="ignore"):
            key = (np.uint64(0xA5A5000000000000) + (np.uint64(k) << np.uint64(40)) + nb) & _M64
        nav = np.where(_splitmix64(seed64 ^ _splitmix64(key)) & np.uint64(1), -1.0, 1.0)
        cyc = (fc + s["doppler_hz"]) * t + s.get("carrier_phase_cycles", 0.0)
        x += s["amp"] * nav * code[chip] * np.cos(2.0 * np.pi * (cyc - np.floor(cyc)))
    return pack_bits_lsb_first((x < 0).astype(np.uint8))

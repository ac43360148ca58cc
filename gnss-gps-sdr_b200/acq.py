"""ctypes binding of include/gpsacq.h and a Python mirror of the reference's SearchTask().

There is deliberately no numerical fallback here: if ``libgpsacq.so`` is missing or no CUDA
device is usable, construction raises.  (The CPU oracle lives in ``oracle/`` and is test
infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

NUM_SATS = 32            # c/gps_offline.h:16
FFT_LEN = 40000          # c/gps_offline.h:15
SNR_THRESHOLD = 25.0     # c/search_offline.cpp:248

_HERE = Path(__file__).resolve().parent


class GpsAcqError(RuntimeError):
    pass


class _Cfg(C.Structure):
    _fields_ = [("fc", C.c_double), ("fs", C.c_double), ("max_fo", C.c_double),
                ("fft_len", C.c_int32), ("device", C.c_int32), ("max_blocks", C.c_int32),
                ("mode", C.c_int32), ("doppler_step", C.c_double), ("noncoh_blocks", C.c_int32),
                ("dop_first", C.c_int32), ("dop_count", C.c_int32), ("reserved", C.c_int32),
                ("fs_replica", C.c_double)]


class _Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("abi_version", "fft_len", "n1", "n2", "window", "dmax", "n_doppler", "chunk_bytes",
                 "max_blocks", "device", "sm_count", "cell_ctas", "cell_threads", "cell_smem_bytes")] + \
               [("bytes_per_corr", C.c_int64), ("mode", C.c_int32), ("noncoh_blocks", C.c_int32),
                ("block_bytes", C.c_int32), ("max_acq", C.c_int32), ("doppler_step", C.c_double),
                ("dop_first", C.c_int32), ("n_doppler_full", C.c_int32), ("blocks_per_launch", C.c_int32),
                ("reserved", C.c_int32)]


class _Sat(C.Structure):
    _fields_ = [("prn", C.c_int32), ("reserved", C.c_int32), ("amp", C.c_double), ("doppler_hz", C.c_double),
                ("code_phase_chips", C.c_double), ("carrier_phase_cycles", C.c_double)]


class _Handoff(C.Structure):
    _fields_ = [("lo_dop_hz", C.c_double), ("ca_dop_hz", C.c_double), ("lo_rate", C.c_uint32), ("ca_rate", C.c_uint32),
                ("ca_shift", C.c_int32), ("ca_pause", C.c_uint32), ("taps", C.c_int32), ("sv", C.c_int32)]


HANDOFF_DTYPE = np.dtype([("lo_dop_hz", "<f8"), ("ca_dop_hz", "<f8"), ("lo_rate", "<u4"), ("ca_rate", "<u4"),
                          ("ca_shift", "<i4"), ("ca_pause", "<u4"), ("taps", "<i4"), ("sv", "<i4")])

PEAK_DTYPE = np.dtype([("snr", "<f4"), ("max_pwr", "<f4"), ("tot_pwr", "<f4"), ("lo_shift", "<i4"),
                       ("ca_shift", "<i4"), ("sv", "<i4"), ("flags", "<i4"), ("reserved", "<i4")])
EVENT_DTYPE = np.dtype([("chunk_index", "<i8"), ("sv", "<i4"), ("ch", "<i4"), ("peak", PEAK_DTYPE), ("start", HANDOFF_DTYPE)])
CELL_DTYPE = np.dtype([("max_pwr", "<f4"), ("tot_pwr", "<f4"), ("max_idx", "<i4"), ("reserved", "<i4")])

_LIB = None


def lib_path() -> Path:
    return Path(os.environ.get("GPSACQ_LIB", _HERE / "csrc" / "libgpsacq.so"))


def load_library() -> C.CDLL:
    """dlopen libgpsacq.so and declare every symbol of include/gpsacq.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise GpsAcqError(f"{p} not found: build it with __graft_entry__.build() "
                          f"(nvcc, sm_100a); there is no CPU fallback")
    lib = C.CDLL(str(p))
    vp, i32p, u8p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    sigs = {
        "gpsacq_create": (C.c_int, [C.POINTER(_Cfg), C.POINTER(vp)]),
        "gpsacq_destroy": (None, [vp]),
        "gpsacq_last_error": (C.c_char_p, [vp]),
        "gpsacq_get_info": (C.c_int, [vp, C.POINTER(_Info)]),
        "gpsacq_set_stream": (C.c_int, [vp, vp]),
        "gpsacq_synchronize": (C.c_int, [vp]),
        "gpsacq_search_blocks": (C.c_int, [vp, vp, C.c_size_t, vp, vp]),
        "gpsacq_search_blocks_device": (C.c_int, [vp, vp, C.c_size_t, vp, vp]),
        "gpsacq_acquire": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "gpsacq_acquire_device": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "gpsacq_iq8_to_bits": (C.c_int, [vp, vp, C.c_size_t, C.c_int, C.c_double, C.c_double, vp]),
        "gpsacq_iq8_to_bits_device": (C.c_int, [vp, vp, C.c_size_t, C.c_int, C.c_double, C.c_double, vp, vp]),
        "gpsacq_bits_to_iq8": (C.c_int, [C.c_int, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_int, vp]),
        "gpsacq_bits_to_iq8_device": (C.c_int, [C.c_int, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_int, vp, vp]),
        "gpsacq_sig_gen_literal": (C.c_int, [C.c_int, C.c_int, vp, C.c_int, vp, vp]),
        "gpsacq_group_create": (C.c_int, [C.POINTER(_Cfg), C.c_int, i32p, C.c_int, C.POINTER(vp)]),
        "gpsacq_group_destroy": (None, [vp]),
        "gpsacq_group_search_blocks": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "gpsacq_group_acquire": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "gpsacq_group_gather_kind": (C.c_char_p, [vp]),
        "gpsacq_group_last_error": (C.c_char_p, [vp]),
        "gpsacq_group_engine": (vp, [vp, C.c_int]),
        "gpsacq_handoff_compute": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp]),
        "gpsacq_service_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "gpsacq_service_destroy": (None, [vp]),
        "gpsacq_service_feed": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(C.c_size_t), vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "gpsacq_service_enable": (C.c_int, [vp, C.c_int]),
        "gpsacq_service_signal_lost": (C.c_int, [vp, C.c_int]),
        "gpsacq_service_state": (C.c_int, [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]),
        "gpsacq_service_last_error": (C.c_char_p, [vp]),
        "gpsacq_synth_capture": (C.c_int, [C.c_int, C.c_double, C.c_double, C.POINTER(_Sat), C.c_int, C.c_double, C.c_double,
                                           C.c_uint64, C.c_size_t, vp, vp]),
        "gpsacq_stage_times": (C.c_int, [vp, f32p]),
        "gpsacq_get_replica_time": (C.c_int, [vp, C.c_int, vp]),
        "gpsacq_get_replica_spectrum": (C.c_int, [vp, C.c_int, vp]),
        "gpsacq_get_block_spectrum": (C.c_int, [vp, C.c_size_t, vp]),
        "gpsacq_get_cell_stats": (C.c_int, [vp, C.c_size_t, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


ABI_SYMBOLS = ("gpsacq_create", "gpsacq_destroy", "gpsacq_last_error", "gpsacq_get_info",
               "gpsacq_set_stream", "gpsacq_synchronize", "gpsacq_search_blocks",
               "gpsacq_search_blocks_device", "gpsacq_acquire", "gpsacq_acquire_device", "gpsacq_iq8_to_bits", "gpsacq_iq8_to_bits_device",
               "gpsacq_bits_to_iq8", "gpsacq_bits_to_iq8_device", "gpsacq_sig_gen_literal", "gpsacq_group_create",
               "gpsacq_group_destroy", "gpsacq_group_search_blocks", "gpsacq_group_acquire", "gpsacq_group_gather_kind", "gpsacq_group_last_error",
               "gpsacq_group_engine", "gpsacq_handoff_compute", "gpsacq_service_create", "gpsacq_service_destroy",
               "gpsacq_service_feed", "gpsacq_service_enable", "gpsacq_service_signal_lost", "gpsacq_service_state",
               "gpsacq_service_last_error", "gpsacq_synth_capture", "gpsacq_stage_times", "gpsacq_get_replica_time",
               "gpsacq_get_replica_spectrum", "gpsacq_get_block_spectrum", "gpsacq_get_cell_stats")


class Acquisition:
    """One engine instance = what SearchInit() sets up (c/search_offline.cpp:74-110).

    fc, fs, max_fo are the reference's FC / FS / max_fo globals (c/gps_offline.h:23-25).
    """

    def __init__(self, fc: float, fs: float, max_fo: float = 5000.0, device: int = -1, max_blocks: int = 0,
                 mode: int = 0, doppler_step: float = 0.0, noncoh_blocks: int = 1, dop_first: int = 0, dop_count: int = 0):
        """dop_first/dop_count (GRID mode): search only that contiguous shard of the Doppler grid -- records keep
        absolute bin numbers, shards are merged with shard.merge_peaks()."""
        self._lib = load_library()
        self._h = C.c_void_p()
        cfg = _Cfg(fc=fc, fs=fs, max_fo=max_fo, fft_len=0, device=device, max_blocks=max_blocks, mode=mode,
                   doppler_step=doppler_step, noncoh_blocks=noncoh_blocks, dop_first=dop_first, dop_count=dop_count, reserved=0)
        rc = self._lib.gpsacq_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._lib.gpsacq_last_error(None)
            self._h = C.c_void_p()
            raise GpsAcqError(f"gpsacq_create failed ({rc}): {msg.decode() if msg else '?'}")
        info = _Info()
        self._check(self._lib.gpsacq_get_info(self._h, C.byref(info)))
        self.info = {n: getattr(info, n) for n, _ in _Info._fields_}
        self.fc, self.fs, self.max_fo = fc, fs, max_fo

    # -- plumbing -------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.gpsacq_last_error(self._h)
            raise GpsAcqError(f"libgpsacq error {rc}: {msg.decode() if msg else '?'}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.gpsacq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def chunk_bytes(self) -> int:
        return self.info["chunk_bytes"]

    @property
    def n_doppler(self) -> int:
        return self.info["n_doppler"]

    # -- the hot path ------------------------------------------------------------------
    def search_blocks(self, bits, sv_of_block=None) -> np.ndarray:
        """Sample()+Correlate() for every 5120-byte chunk of `bits` (host memory).

        Chunk b is searched for PRN sv_of_block[b]+1 (default: b mod 32, the order
        SearchTask() walks the file, c/search_offline.cpp:239-246).  Returns PEAK_DTYPE records.
        """
        buf = np.ascontiguousarray(np.frombuffer(bits, dtype=np.uint8) if not isinstance(bits, np.ndarray) else bits,
                                   dtype=np.uint8)
        cb = self.chunk_bytes
        n_blocks = buf.size // cb
        out = np.zeros(n_blocks, dtype=PEAK_DTYPE)
        svp = None
        if sv_of_block is not None:
            sv = np.ascontiguousarray(sv_of_block, dtype=np.int32)
            if sv.size != n_blocks:
                raise ValueError("sv_of_block must have one entry per chunk")
            svp = sv.ctypes.data
        self._check(self._lib.gpsacq_search_blocks(self._h, buf.ctypes.data, n_blocks, svp, out.ctypes.data))
        return out

    # -- GRID mode (mode=1): 1 ms blocks, explicit Doppler grid, K-block non-coherent sum --------------
    @property
    def acq_bytes(self) -> int:
        return self.info["block_bytes"] * self.info["noncoh_blocks"]

    def acquire(self, bits) -> np.ndarray:
        """GRID mode: every `acq_bytes` of `bits` is one acquisition of all 32 PRNs; returns 32 records each."""
        buf = np.ascontiguousarray(np.frombuffer(bits, dtype=np.uint8) if not isinstance(bits, np.ndarray) else bits,
                                   dtype=np.uint8)
        n_acq = buf.size // self.acq_bytes
        out = np.zeros(n_acq * NUM_SATS, dtype=PEAK_DTYPE)
        self._check(self._lib.gpsacq_acquire(self._h, buf.ctypes.data, n_acq, out.ctypes.data))
        return out

    def acquire_device(self, d_bits_ptr: int, n_acq: int, d_out_ptr: int):
        self._check(self._lib.gpsacq_acquire_device(self._h, d_bits_ptr, n_acq, d_out_ptr))

    # -- 8-bit IQ front-end (proc_rtl_bin_for_gps.m / proc_hackrf_bin_for_gps.m on the GPU) ----------------
    def iq8_to_bits(self, iq, shift_hz: float, fs: float | None = None, signed: bool = False) -> np.ndarray:
        """Interleaved I,Q bytes (uint8 offset-128 for rtl-sdr, int8 when `signed`) -> mean removal ->
        shift up by shift_hz -> real part -> packed 1-bit samples (LSB first)."""
        buf = np.ascontiguousarray(np.frombuffer(iq, dtype=np.uint8) if not isinstance(iq, np.ndarray) else iq.view(np.uint8))
        n = buf.size // 2
        out = np.zeros((n + 7) // 8, np.uint8)
        self._check(self._lib.gpsacq_iq8_to_bits(self._h, buf.ctypes.data, n, 1 if signed else 0, shift_hz,
                                                 fs if fs is not None else self.fs, out.ctypes.data))
        return out

    def iq8_to_bits_device(self, d_iq_ptr: int, n_samples: int, shift_hz: float, fs: float, d_bits_ptr: int, d_sums_ptr: int,
                           signed: bool = False):
        """Device-buffer form of iq8_to_bits, asynchronous on the handle's stream (ratios shift_hz/fs that are not a small
        fraction p/q, q <= 227, read the mean back once inside the call)."""
        self._check(self._lib.gpsacq_iq8_to_bits_device(self._h, d_iq_ptr, n_samples, 1 if signed else 0, shift_hz, fs,
                                                        d_bits_ptr, d_sums_ptr))

    def search_blocks_device(self, d_bits_ptr: int, n_blocks: int, d_sv_ptr: int | None, d_out_ptr: int):
        """Asynchronous device-pointer variant (raw CUDA device addresses)."""
        self._check(self._lib.gpsacq_search_blocks_device(self._h, d_bits_ptr, n_blocks, d_sv_ptr, d_out_ptr))

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.gpsacq_set_stream(self._h, cuda_stream_ptr))

    def synchronize(self):
        self._check(self._lib.gpsacq_synchronize(self._h))

    def stage_times(self) -> dict:
        ms = (C.c_float * 4)()
        self._check(self._lib.gpsacq_stage_times(self._h, ms))
        return {"fwd_ms": ms[0], "cells_ms": ms[1], "best_ms": ms[2], "total_ms": ms[3]}

    # -- parity probes -----------------------------------------------------------------
    def replica_time(self, sv: int) -> np.ndarray:
        out = np.empty(self.info["fft_len"], np.float32)
        self._check(self._lib.gpsacq_get_replica_time(self._h, sv, out.ctypes.data))
        return out

    def replica_spectrum(self, sv: int) -> np.ndarray:
        out = np.empty(self.info["fft_len"], np.complex64)
        self._check(self._lib.gpsacq_get_replica_spectrum(self._h, sv, out.ctypes.data))
        return out

    def block_spectrum(self, block: int) -> np.ndarray:
        out = np.empty(self.info["fft_len"], np.complex64)
        self._check(self._lib.gpsacq_get_block_spectrum(self._h, block, out.ctypes.data))
        return out

    def cell_stats(self, block: int) -> np.ndarray:
        out = np.zeros(self.n_doppler, CELL_DTYPE)
        self._check(self._lib.gpsacq_get_cell_stats(self._h, block, out.ctypes.data))
        return out


def synth_capture_gpu(n_samples: int, fs: float, fc: float, sats, seed: int = 1, noise_sigma: float = 1.0,
                      nav_bps: float = 50.0, device: int = 0, d_out_ptr: int | None = None) -> np.ndarray | None:
    """Synthetic packed 1-bit IF capture generated on the GPU (gpsacq_synth_capture).  sats: dicts with prn,
    amp, doppler_hz, code_phase_chips and optionally carrier_phase_cycles.  Returns the bytes (host) unless
    d_out_ptr (a device address) is given."""
    lib = load_library()
    arr = (_Sat * len(sats))()
    for i, s in enumerate(sats):
        arr[i] = _Sat(prn=s["prn"], reserved=0, amp=s["amp"], doppler_hz=s["doppler_hz"],
                      code_phase_chips=s["code_phase_chips"], carrier_phase_cycles=s.get("carrier_phase_cycles", 0.0))
    out = None if d_out_ptr else np.zeros((n_samples + 7) // 8, np.uint8)
    rc = lib.gpsacq_synth_capture(device, fs, fc, arr, len(sats), noise_sigma, nav_bps, seed, n_samples,
                                  out.ctypes.data if out is not None else None, d_out_ptr)
    if rc != 0:
        msg = lib.gpsacq_last_error(None)
        raise GpsAcqError(f"gpsacq_synth_capture failed ({rc}): {msg.decode() if msg else '?'}")
    return out


def _free_call(rc: int, what: str):
    if rc != 0:
        msg = load_library().gpsacq_last_error(None)
        raise GpsAcqError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def bits_to_iq8(bits, fc: float, fs: float, amplitude: int = 30, first_sample: int = 0, device: int = 0) -> np.ndarray:
    """The reference's c/conv_1bit_bin_to_hackrf_bin.cpp on the GPU: packed 1-bit real-IF samples -> interleaved int8
    I,Q at baseband (16 output bytes per input byte).  fc / fs: what the reference takes from c/gps.h (2.6e6 / 10e6)."""
    buf = np.ascontiguousarray(np.frombuffer(bits, np.uint8) if not isinstance(bits, np.ndarray) else bits)
    out = np.zeros(16 * buf.size, np.int8)
    _free_call(load_library().gpsacq_bits_to_iq8(device, buf.ctypes.data, buf.size, first_sample, fc, fs, amplitude, out.ctypes.data),
               "gpsacq_bits_to_iq8")
    return out


def bits_to_iq8_device(d_bits_ptr: int, n_bytes: int, fc: float, fs: float, d_out_ptr: int, amplitude: int = 30,
                       first_sample: int = 0, device: int = -1, stream_ptr: int | None = None):
    _free_call(load_library().gpsacq_bits_to_iq8_device(device, d_bits_ptr, n_bytes, first_sample, fc, fs, amplitude, d_out_ptr, stream_ptr),
               "gpsacq_bits_to_iq8_device")


def sig_gen_literal(prn: int, nav_bits01, device: int = 0) -> np.ndarray:
    """gps_sig_gen.m:8-41 on the GPU, literally (x8 zero-stuff, 20 periods per NAV bit, rcosine(1,8), carrier at fs/4,
    sign, 'ubit1'): the bytes of the file it writes for satellite `prn` and the given NAV bits."""
    nav = np.ascontiguousarray(nav_bits01, np.uint8)
    out = np.zeros((nav.size * 163680 + 48 + 7) // 8, np.uint8)
    _free_call(load_library().gpsacq_sig_gen_literal(device, prn, nav.ctypes.data, nav.size, out.ctypes.data, None), "gpsacq_sig_gen_literal")
    return out


def handoff(peak, fc: float, fs: float, bin_num: float, bin_den: float, secs_since_sample: float) -> np.ndarray:
    """CHANNEL::Start() values for one acquisition record (c/channel.cpp:134-171); needs no GPU.
    Doppler = lo_shift*bin_num/bin_den: (FS, FFT_LEN) in REF mode, (doppler_step, 1) in GRID mode."""
    lib = load_library()
    rec = np.zeros(1, PEAK_DTYPE)
    rec[0] = peak
    out = np.zeros(1, HANDOFF_DTYPE)
    rc = lib.gpsacq_handoff_compute(rec.ctypes.data, fc, fs, bin_num, bin_den, secs_since_sample, out.ctypes.data)
    if rc != 0:
        raise GpsAcqError(f"gpsacq_handoff_compute failed ({rc})")
    return out[0]


class SearchService:
    """The receiver's SearchTask() loop over a chunk stream (c/search.cpp:214-239): skip tracked SVs, one chunk per
    searched SV, channel allocation, detection events with their CHANNEL::Start() hand-off (include/gpsacq.h,
    gpsacq_service_*)."""

    def __init__(self, acq: Acquisition, num_chans: int = 0, max_rounds_per_batch: int = 0):
        self._lib = load_library()
        self._acq = acq
        self._s = C.c_void_p()
        rc = self._lib.gpsacq_service_create(acq._h, num_chans, max_rounds_per_batch, C.byref(self._s))
        if rc != 0:
            raise GpsAcqError(f"gpsacq_service_create failed ({rc}): {self._lib.gpsacq_last_error(acq._h).decode()}")

    def feed(self, chunks, max_events: int = 64):
        """Returns (chunks consumed, EVENT_DTYPE records)."""
        buf = np.ascontiguousarray(np.frombuffer(chunks, dtype=np.uint8) if not isinstance(chunks, np.ndarray) else chunks, dtype=np.uint8)
        n = buf.size // self._acq.chunk_bytes
        ev = np.zeros(max_events, EVENT_DTYPE)
        used, nev = C.c_size_t(), C.c_size_t()
        rc = self._lib.gpsacq_service_feed(self._s, buf.ctypes.data, n, C.byref(used), ev.ctypes.data, max_events, C.byref(nev))
        if rc != 0:
            raise GpsAcqError(f"gpsacq_service_feed failed ({rc}): {self._lib.gpsacq_service_last_error(self._s).decode()}")
        return used.value, ev[: nev.value].copy()

    def enable(self, sv: int):
        if self._lib.gpsacq_service_enable(self._s, sv) != 0:
            raise GpsAcqError("bad sv")

    def signal_lost(self, ch: int):
        if self._lib.gpsacq_service_signal_lost(self._s, ch) != 0:
            raise GpsAcqError("bad channel")

    def state(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_int64()
        self._lib.gpsacq_service_state(self._s, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def close(self):
        if self._s and self._s.value:
            self._lib.gpsacq_service_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AcquisitionGroup:
    """Several GPUs in one process: REF mode splits the chunks of a batch, GRID mode splits the Doppler bins of
    every acquisition; one ncclAllGather of the peak records per batch (include/gpsacq.h, gpsacq_group_*).
    `devices` may name a device more than once (then the records are gathered through the host)."""

    def __init__(self, fc: float, fs: float, max_fo: float = 5000.0, n_gpus: int = 2, use_nccl: bool = True, max_blocks: int = 0,
                 mode: int = 0, doppler_step: float = 0.0, noncoh_blocks: int = 1, devices=None):
        self._lib = load_library()
        self._g = C.c_void_p()
        cfg = _Cfg(fc=fc, fs=fs, max_fo=max_fo, fft_len=0, device=-1, max_blocks=max_blocks, mode=mode, doppler_step=doppler_step,
                   noncoh_blocks=noncoh_blocks, dop_first=0, dop_count=0, reserved=0)
        devs = None
        if devices is not None:
            devs = (C.c_int32 * n_gpus)(*devices)
        self.acq_bytes = int(round(fs / 1000)) // 8 * noncoh_blocks
        rc = self._lib.gpsacq_group_create(C.byref(cfg), n_gpus, devs, 1 if use_nccl else 0, C.byref(self._g))
        if rc != 0:
            msg = self._lib.gpsacq_last_error(None)
            raise GpsAcqError(f"gpsacq_group_create failed ({rc}): {msg.decode() if msg else '?'}")
        self.gather_kind = self._lib.gpsacq_group_gather_kind(self._g).decode()
        self.chunk_bytes = 5120

    def search_blocks(self, bits) -> np.ndarray:
        buf = np.ascontiguousarray(np.frombuffer(bits, dtype=np.uint8) if not isinstance(bits, np.ndarray) else bits, dtype=np.uint8)
        n = buf.size // self.chunk_bytes
        out = np.zeros(n, dtype=PEAK_DTYPE)
        rc = self._lib.gpsacq_group_search_blocks(self._g, buf.ctypes.data, n, out.ctypes.data)
        if rc != 0:
            raise GpsAcqError(f"group search failed ({rc}): {self._lib.gpsacq_group_last_error(self._g).decode()}")
        return out

    def acquire(self, bits) -> np.ndarray:
        """GRID group: 32 records per acquisition, identical to a single-GPU handle's."""
        buf = np.ascontiguousarray(np.frombuffer(bits, dtype=np.uint8) if not isinstance(bits, np.ndarray) else bits, dtype=np.uint8)
        n = buf.size // self.acq_bytes
        out = np.zeros(n * 32, dtype=PEAK_DTYPE)
        rc = self._lib.gpsacq_group_acquire(self._g, buf.ctypes.data, n, out.ctypes.data)
        if rc != 0:
            raise GpsAcqError(f"group acquire failed ({rc}): {self._lib.gpsacq_group_last_error(self._g).decode()}")
        return out

    def close(self):
        if self._g and self._g.value:
            self._lib.gpsacq_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- SearchTask() report formatting (c/search_offline.cpp:264-287) -------------------------
def format_run(run_count: int, peaks: np.ndarray) -> str:
    """The six stdout lines SearchTask() prints for one run of 32 chunks."""
    hits = [p for p in peaks if not (p["snr"] < SNR_THRESHOLD)]
    out = []
    out.append("%2d satellite: " % run_count + "".join("%5d " % p["sv"] for p in hits))
    out.append("%2d SNR(>=25): " % run_count + "".join("%5.1f " % p["snr"] for p in hits))
    out.append("%2d  lo_shift: " % run_count + "".join("%5d " % p["lo_shift"] for p in hits))
    out.append("%2d  ca_shift: " % run_count + "".join("%5d " % p["ca_shift"] for p in hits))
    out.append("".join("%2.0f " % p["snr"] for p in peaks))
    out.append("")
    return "\n".join(out) + "\n"


def search_task_text(acq: Acquisition, filename: str, runs_per_batch: int = 16, max_runs: int | None = None) -> str:
    """Python mirror of SearchTask(char*) (c/search_offline.cpp:219-292): same traversal of the
    file (32 consecutive chunks per run, partial run discarded with "run out of file!"),
    same report text.  Returns what the reference would print after the banner."""
    try:
        fp = open(filename, "rb")
    except OSError:
        return "can not open file!\n"
    text = []
    run_bytes = NUM_SATS * acq.chunk_bytes
    run_count = 0
    with fp:
        while max_runs is None or run_count < max_runs:
            want = runs_per_batch if max_runs is None else min(runs_per_batch, max_runs - run_count)
            data = fp.read(want * run_bytes)
            full = len(data) // run_bytes
            if full:
                peaks = acq.search_blocks(data[: full * run_bytes])
                for r in range(full):
                    text.append(format_run(run_count, peaks[r * NUM_SATS:(r + 1) * NUM_SATS]))
                    run_count += 1
            if full < want:
                text.append("run out of file!\n")
                break
    return "".join(text)

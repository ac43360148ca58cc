"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU oracle:
  * ``Oracle``      -- the C restatement (oracle/gpsacq_oracle.c -> oracle/liboracle.so)
  * ``RefHarness``  -- the UNMODIFIED reference translation unit compiled against the
                       fftw3.h stand-in (oracle/_ref/libref_harness.so; built only where
                       /root/reference exists, prebuilt file travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (gnss-gps-sdr_b200/, include/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PEAK_DTYPE = np.dtype([("snr", "<f4"), ("max_pwr", "<f4"), ("tot_pwr", "<f4"), ("lo_shift", "<i4"),
                       ("ca_shift", "<i4"), ("sv", "<i4"), ("flags", "<i4"), ("reserved", "<i4")])
TORCH_MKL = None


def build(quiet: bool = True) -> None:
    """make -C oracle (liboracle.so always; _ref/ only when /root/reference is present)."""
    subprocess.run(["make", "-C", str(HERE)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _lib() -> C.CDLL:
    p = HERE / "liboracle.so"
    if not p.exists():
        build()
    lib = C.CDLL(str(p))
    vp = C.c_void_p
    lib.oracle_create.restype = vp
    lib.oracle_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.oracle_destroy.argtypes = [vp]
    for f in ("oracle_num_doppler", "oracle_window", "oracle_chunk_bytes", "oracle_fft_len"):
        getattr(lib, f).restype = C.c_int
        getattr(lib, f).argtypes = [vp]
    lib.oracle_code_spectrum.restype = vp
    lib.oracle_code_spectrum.argtypes = [vp, C.c_int]
    lib.oracle_sample.argtypes = [vp, vp, vp]
    lib.oracle_cells.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    lib.oracle_search_blocks.restype = C.c_int
    lib.oracle_search_blocks.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.oracle_cacode_chips.argtypes = [C.c_int, vp]
    lib.oracle_search_code.restype = C.c_int
    lib.oracle_search_code.argtypes = [C.c_int, C.c_int]
    lib.oracle_replica_time.argtypes = [C.c_double, C.c_int, C.c_int, vp]
    lib.oracle_lo_table.argtypes = [C.c_double, C.c_double, C.c_int, vp]
    lib.oracle_grid_create.restype = vp
    lib.oracle_grid_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    lib.oracle_grid_destroy.argtypes = [vp]
    lib.oracle_grid_num_doppler.restype = C.c_int
    lib.oracle_grid_num_doppler.argtypes = [vp]
    lib.oracle_grid_window.restype = C.c_int
    lib.oracle_grid_window.argtypes = [vp]
    lib.oracle_grid_acquire.restype = C.c_int
    lib.oracle_grid_acquire.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.oracle_grid_acquire_svs.restype = C.c_int
    lib.oracle_grid_acquire_svs.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp]
    lib.oracle_conv_1bit_iq8.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_int, vp]
    lib.oracle_sig_gen_literal.argtypes = [C.c_int, vp, C.c_int, vp, vp]
    return lib


_L = None


def lib() -> C.CDLL:
    global _L
    if _L is None:
        _L = _lib()
    return _L


def cacode_chips(sv: int) -> np.ndarray:
    out = np.zeros(1023, np.uint8)
    lib().oracle_cacode_chips(sv, out.ctypes.data)
    return out


def search_code(sv: int, g1: int) -> int:
    return lib().oracle_search_code(sv, g1)


def replica_time(fs: float, sv: int, n: int = 40000) -> np.ndarray:
    out = np.zeros(n, np.float32)
    lib().oracle_replica_time(fs, sv, n, out.ctypes.data)
    return out


def lo_table(fc: float, fs: float, n: int = 40960) -> np.ndarray:
    out = np.zeros(n, np.uint8)
    lib().oracle_lo_table(fc, fs, n, out.ctypes.data)
    return out


class Oracle:
    """C restatement of SearchInit()/Sample()/Correlate() (see gpsacq_oracle.c for the line map)."""

    def __init__(self, fc: float, fs: float, max_fo: float = 5000.0, fft_len: int = 40000, fft_f64: bool = True):
        self._l = lib()
        self._h = self._l.oracle_create(fc, fs, max_fo, fft_len, 1 if fft_f64 else 0)
        if not self._h:
            raise MemoryError("oracle_create failed")
        self.n = self._l.oracle_fft_len(self._h)
        self.n_doppler = self._l.oracle_num_doppler(self._h)
        self.dmax = (self.n_doppler - 1) // 2
        self.window = self._l.oracle_window(self._h)
        self.chunk_bytes = self._l.oracle_chunk_bytes(self._h)

    def close(self):
        if self._h:
            self._l.oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def code_spectrum(self, sv: int) -> np.ndarray:
        p = self._l.oracle_code_spectrum(self._h, sv)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(2 * self.n,)).view(np.complex64).copy()

    def sample(self, chunk) -> np.ndarray:
        buf = np.ascontiguousarray(np.frombuffer(chunk, np.uint8))
        out = np.zeros(self.n, np.complex64)
        self._l.oracle_sample(self._h, buf.ctypes.data, out.ctypes.data)
        return out

    def cells(self, chunk, sv: int):
        buf = np.ascontiguousarray(np.frombuffer(chunk, np.uint8))
        mp = np.zeros(self.n_doppler, np.float32)
        mi = np.zeros(self.n_doppler, np.int32)
        tp = np.zeros(self.n_doppler, np.float32)
        self._l.oracle_cells(self._h, buf.ctypes.data, sv, mp.ctypes.data, mi.ctypes.data, tp.ctypes.data)
        return mp, mi, tp

    def search_blocks(self, bits, sv_of_block=None) -> np.ndarray:
        buf = np.ascontiguousarray(np.frombuffer(bits, np.uint8) if not isinstance(bits, np.ndarray) else bits)
        nb = buf.size // self.chunk_bytes
        out = np.zeros(nb, PEAK_DTYPE)
        svp = None
        if sv_of_block is not None:
            sv = np.ascontiguousarray(sv_of_block, np.int32)
            svp = sv.ctypes.data
        rc = self._l.oracle_search_blocks(self._h, buf.ctypes.data, nb, svp, out.ctypes.data)
        if rc:
            raise MemoryError("oracle_search_blocks failed")
        return out


class GridOracle:
    """GRID-mode definition with true W-point FFTs; see gpsacq_oracle.c."""

    def __init__(self, fc: float, fs: float, max_fo: float, step: float, kblocks: int = 1):
        self._l = lib()
        self._h = self._l.oracle_grid_create(fc, fs, max_fo, step, kblocks)
        if not self._h:
            raise MemoryError("oracle_grid_create failed")
        self.n_doppler = self._l.oracle_grid_num_doppler(self._h)
        self.dmax = (self.n_doppler - 1) // 2
        self.window = self._l.oracle_grid_window(self._h)
        self.kblocks = kblocks
        self.acq_bytes = kblocks * self.window // 8

    def __del__(self):
        try:
            if self._h:
                self._l.oracle_grid_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def acquire(self, bits, want_cells: bool = False, svs=None):
        """svs: optional list of 0-based PRN indices (default all 32); records / cell rows come in that order."""
        buf = np.ascontiguousarray(np.frombuffer(bits, np.uint8) if not isinstance(bits, np.ndarray) else bits)
        n_acq = buf.size // self.acq_bytes
        sv = np.arange(32, dtype=np.int32) if svs is None else np.ascontiguousarray(svs, np.int32)
        ns = len(sv)
        out = np.zeros(n_acq * ns, PEAK_DTYPE)
        cells = []
        for a in range(n_acq):
            cm = np.zeros((ns, self.n_doppler), np.float32)
            ci = np.zeros((ns, self.n_doppler), np.int32)
            ct = np.zeros((ns, self.n_doppler), np.float32)
            sub = np.ascontiguousarray(buf[a * self.acq_bytes:(a + 1) * self.acq_bytes])
            one = np.zeros(ns, PEAK_DTYPE)
            rc = self._l.oracle_grid_acquire_svs(self._h, sub.ctypes.data, sv.ctypes.data, ns, one.ctypes.data,
                                                 cm.ctypes.data, ci.ctypes.data, ct.ctypes.data)
            if rc:
                raise MemoryError("oracle_grid_acquire_svs failed")
            out[a * ns:(a + 1) * ns] = one
            cells.append((cm, ci, ct))
        return (out, cells) if want_cells else out


def conv_1bit_iq8(bits, fc: float, fs: float, amp: int = 30, first_sample: int = 0) -> np.ndarray:
    """c/conv_1bit_bin_to_hackrf_bin.cpp:29-86 restated: packed 1-bit IF -> interleaved int8 I,Q."""
    buf = np.ascontiguousarray(np.frombuffer(bits, np.uint8) if not isinstance(bits, np.ndarray) else bits)
    out = np.zeros(16 * buf.size, np.int8)
    lib().oracle_conv_1bit_iq8(buf.ctypes.data, buf.size, first_sample, fc, fs, amp, out.ctypes.data)
    return out


def rcosine_1_8() -> np.ndarray:
    """MATLAB rcosine(1, 8): normal raised-cosine FIR, roll-off 0.5, delay 3 -> 49 taps, in double."""
    n = np.arange(-24, 25) / 8.0
    with np.errstate(divide="ignore", invalid="ignore"):
        h = np.sinc(n) * np.cos(np.pi * 0.5 * n) / (1 - n ** 2)
    h[np.isclose(np.abs(n), 1)] = np.pi / 4 * np.sinc(1.0)
    return h


def sig_gen_literal(sv: int, nav01) -> np.ndarray:
    """gps_sig_gen.m:8-41 restated (see gpsacq_oracle.c): the packed 'ubit1' bytes for satellite index sv (PRN-1)."""
    nav = np.ascontiguousarray(nav01, np.uint8)
    n_out = nav.size * 163680 + 48
    out = np.zeros((n_out + 7) // 8, np.uint8)
    taps = np.ascontiguousarray(rcosine_1_8(), np.float64)
    lib().oracle_sig_gen_literal(sv, nav.ctypes.data, nav.size, taps.ctypes.data, out.ctypes.data)
    return out


def recover_nav_bits(file_bits: np.ndarray, sv: int) -> np.ndarray:
    """The NAV bits gps_sig_gen.m drew with rand, read back from the file it wrote: the sign of each 20 ms stretch
    against a generation with all-zero NAV bits."""
    n = file_bits.size * 8 // 163680
    ref = np.unpackbits(sig_gen_literal(sv, np.zeros(n, np.uint8)), bitorder="little")[24: 24 + n * 163680]
    got = np.unpackbits(file_bits, bitorder="little")[24: 24 + n * 163680]
    agree = (ref == got).reshape(n, 163680).mean(1)
    assert np.all(np.abs(agree - 0.5) > 0.4), "file is not a gps_sig_gen.m product for this satellite"
    return (agree < 0.5).astype(np.uint8)


def mkl_env() -> dict:
    """Environment that makes the fftw3.h stand-in use MKL (exported by torch's libtorch_cpu.so)."""
    env = dict(os.environ)
    try:
        import torch  # only to locate the library
        p = Path(torch.__file__).parent / "lib" / "libtorch_cpu.so"
        if p.exists():
            env.update(ORACLE_FFT="mkl", ORACLE_MKL_LIB=str(p), MKL_NUM_THREADS="1", OMP_NUM_THREADS="1")
    except Exception:
        pass
    return env


def ref_available() -> bool:
    return (HERE / "_ref" / "libref_harness.so").exists()


class RefHarness:
    """The real reference (c/search_offline.cpp, unmodified) behind extern "C" probes.
    One instance per process: the reference keeps its state in file statics."""

    def __init__(self, fc: float, fs: float, max_fo: float = 5000.0):
        p = HERE / "_ref" / "libref_harness.so"
        if not p.exists():
            raise FileNotFoundError(f"{p}: build with `make -C oracle` where /root/reference exists")
        self._l = C.CDLL(str(p))
        vp = C.c_void_p
        self._l.ref_init.restype = C.c_int
        self._l.ref_init.argtypes = [C.c_double, C.c_double, C.c_double]
        self._l.ref_fft_backend.restype = C.c_char_p
        self._l.ref_get_code.argtypes = [C.c_int, vp]
        self._l.ref_sample.restype = C.c_int
        self._l.ref_sample.argtypes = [vp, C.c_size_t, vp]
        self._l.ref_search_blocks.restype = C.c_int
        self._l.ref_search_blocks.argtypes = [vp, C.c_size_t, C.c_int, vp, vp, vp, vp]
        self._l.ref_search_code.restype = C.c_int
        self._l.ref_search_code.argtypes = [C.c_int, C.c_int]
        if self._l.ref_init(fc, fs, max_fo) != 0:
            raise RuntimeError("reference SearchInit() failed")
        self.n = 40000
        self.chunk_bytes = 5120

    @property
    def fft_backend(self) -> str:
        return self._l.ref_fft_backend().decode()

    def code_spectrum(self, sv: int) -> np.ndarray:
        out = np.zeros(self.n, np.complex64)
        self._l.ref_get_code(sv, out.ctypes.data)
        return out

    def sample(self, chunk) -> np.ndarray:
        buf = np.ascontiguousarray(np.frombuffer(chunk, np.uint8))
        out = np.zeros(self.n, np.complex64)
        if self._l.ref_sample(buf.ctypes.data, buf.size, out.ctypes.data) != 0:
            raise RuntimeError("reference Sample() ran out of data")
        return out

    def search_blocks(self, bits, sv_of_block=None) -> np.ndarray:
        buf = np.ascontiguousarray(np.frombuffer(bits, np.uint8) if not isinstance(bits, np.ndarray) else bits)
        nb = buf.size // self.chunk_bytes
        snr = np.zeros(nb, np.float32)
        lo = np.zeros(nb, np.int32)
        ca = np.zeros(nb, np.int32)
        svp = None
        if sv_of_block is not None:
            sv = np.ascontiguousarray(sv_of_block, np.int32)
            svp = sv.ctypes.data
        done = self._l.ref_search_blocks(buf.ctypes.data, buf.size, nb, svp, snr.ctypes.data, lo.ctypes.data, ca.ctypes.data)
        if done != nb:
            raise RuntimeError(f"reference stopped after {done} of {nb} chunks")
        out = np.zeros(nb, PEAK_DTYPE)
        out["snr"], out["lo_shift"], out["ca_shift"] = snr, lo, ca
        out["sv"] = np.arange(nb) % 32 if sv_of_block is None else np.asarray(sv_of_block)
        out["flags"] = (~(snr < 25)).astype(np.int32)
        return out

    def search_code(self, sv: int, g1: int) -> int:
        return self._l.ref_search_code(sv, g1)


# ---- consumers of the acquisition records -----------------------------------------------
# The receiver's own c/search.cpp + c/channel.cpp, UNMODIFIED, run here behind a stand-in for their SPI / coroutine /
# clock layer (ref_target_harness.cpp -> oracle/_ref/libref_target.so): RefTarget below.  Its event log
# (tests/golden/ref_target_events.json, made by tests/golden/make_golden_target.py) pins the plain-Python
# restatements that follow, and through them -- and directly -- the engine's hand-off and service loop.
SPI_CMDS = ("CmdSample", "CmdSetMask", "CmdSetRateCA", "CmdSetRateLO", "CmdSetGainCA", "CmdSetGainLO", "CmdSetSV", "CmdPause",
            "CmdSetVCO", "CmdGetSamples", "CmdGetChan", "CmdGetClocks", "CmdGetGlitches", "CmdSetDAC", "CmdSetLCD", "CmdGetJoy")   # c/spi.h:9-26


class RefTarget:
    """One run of the receiver's SearchTask() + 12 ChanTask()s (the reference's own code) over a chunk stream.
    One instance per process.  fc / fs are the macros of c/gps.h (2.6 MHz / 10 MHz)."""

    def __init__(self):
        p = HERE / "_ref" / "libref_target.so"
        if not p.exists():
            raise FileNotFoundError(f"{p}: build with `make -C oracle` where /root/reference exists")
        self._l = C.CDLL(str(p))
        self._l.reft_fc.restype = C.c_double
        self._l.reft_fs.restype = C.c_double
        self._l.reft_run.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_int]
        self.fc, self.fs, self.num_chans = self._l.reft_fc(), self._l.reft_fs(), self._l.reft_num_chans()

    def run(self, bits, us_per_yield: int = 2000, max_rec: int = 1 << 16):
        """Returns the parsed event list: dicts {"type": "start", chunk, ch, sv, taps, lo_shift, ca_shift, secs, lo_rate,
        ca_rate, ca_pause, mask} and {"type": "lost", chunk, ch, sv, mask} in the order they happened."""
        buf = np.ascontiguousarray(np.frombuffer(bits, np.uint8) if not isinstance(bits, np.ndarray) else bits)
        rec = np.zeros((max_rec, 8), np.int32)
        n = self._l.reft_run(buf.ctypes.data, buf.size, us_per_yield, rec.ctypes.data, max_rec)
        if n < 0 or n > max_rec:
            raise RuntimeError(f"reft_run returned {n}")
        return parse_target_log(rec[:n])


def parse_target_log(rec) -> list:
    """One pass over the log.  CHANNEL::Start() (c/channel.cpp:134-171) spans a TimerWait(3) during which other tasks run,
    so what it sends is matched by channel number: CmdSetRateLO / CmdSetRateCA / CmdPause / CmdSetSV with wparam == ch
    up to the CmdSetMask that sets the channel's bit; a CmdSetMask that clears bits is CHANNEL::SignalLost() (:245-249)."""
    events, mask, sv_of_ch, pending = [], 0, {}, {}
    for r in rec:
        kind, chunk, a, b, c, d, e, u = (int(v) for v in r)
        u &= 0xFFFFFFFF
        if kind == 0:                                   # ChanStart(ch, sv, t_sample, taps, lo_shift, ca_shift)
            ev = dict(type="start", chunk=chunk, ch=a, sv=b, taps=c, lo_shift=d, ca_shift=e, secs=u / 1e6, ca_pause=0)
            pending[a] = ev
            sv_of_ch[a] = b
            events.append(ev)
            continue
        name, w = SPI_CMDS[a], b
        if name == "CmdSetMask":
            for ch in range(32):
                if (mask >> ch) & 1 and not (w >> ch) & 1:
                    events.append(dict(type="lost", chunk=chunk, ch=ch, sv=sv_of_ch.get(ch, -1), mask=w))
            for ch in list(pending):
                if (w >> ch) & 1:
                    pending.pop(ch)["mask"] = w
            mask = w
        elif w in pending:
            ev = pending[w]
            if name == "CmdSetRateLO": ev["lo_rate"] = u
            elif name == "CmdSetRateCA": ev["ca_rate"] = u
            elif name == "CmdPause": ev["ca_pause"] = u + 1                      # spi_set(CmdPause, ch, ca_pause-1), :165
            elif name == "CmdSetSV": ev["taps_sent"] = u
    return events
L1_HZ = 1575.42e6          # c/gps.h:22
CPS_HZ = 1.023e6           # c/gps.h:25
SATS_TAPS = [(2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6), (6, 7),
             (7, 8), (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6), (5, 7), (6, 8),
             (7, 9), (8, 10), (1, 6), (2, 7), (3, 8), (4, 9)]      # Sats[], c/search.cpp:16-53 (T1, T2)


def channel_start(sv: int, lo_shift: int, ca_shift: int, fc: float, fs: float, fft_len: int, secs: float) -> dict:
    """CHANNEL::Start(), c/channel.cpp:134-171, as arithmetic: what it writes to the NCOs.  `secs` is
    (Microseconds()-t_sample)/1e6 (:155).  The constants 20000/10000 of :164 are 2 and 1 code periods at the
    receiver's FS = 10 MHz; W = ceil(fs/1000) generalises them."""
    import math
    lo_dop = lo_shift * fs / fft_len                                  # :147
    ca_dop = lo_dop / L1_HZ * CPS_HZ                                  # :148
    lo_rate = int((fc + lo_dop) / fs * math.pow(2, 32)) & 0xFFFFFFFF   # :151 (uint32_t conversion truncates)
    ca_rate = int((CPS_HZ + ca_dop) / fs * math.pow(2, 32)) & 0xFFFFFFFF   # :152
    ca_shift = ca_shift + int(np.rint(ca_dop * secs * fs / CPS_HZ))    # :161 nearbyint = round-half-even
    w = math.ceil(fs / 1000.0)
    ca_pause = int(math.fmod(2 * w - ca_shift, w))                     # :164 C's % truncates toward zero
    t1, t2 = SATS_TAPS[sv]
    return dict(lo_dop_hz=lo_dop, ca_dop_hz=ca_dop, lo_rate=lo_rate, ca_rate=ca_rate, ca_shift=ca_shift,
                ca_pause=ca_pause & 0xFFFFFFFF, taps=(t1 << 4) + t2, sv=sv)      # taps: c/search.cpp:236-237


def search_task_on_target(engine, data: bytes, num_chans: int = 12, busy=None, chan_busy: int = 0, next_sv: int = 0):
    """SearchTask() of the receiver, c/search.cpp:214-239, ONE chunk at a time over `data`: for sv in 0..31 round
    robin; skip Busy[sv]; stop sampling while all channels are busy (ChanReset() < 0, c/channel.cpp:396-404);
    Sample() = the next chunk; Correlate(); snr < 25 -> continue; else Busy[sv] = true and ChanStart(ch, ...).
    `engine` is an Oracle (or RefHarness).  Returns (events, chunks consumed, busy, chan_busy, next_sv)."""
    busy = [False] * 32 if busy is None else list(busy)
    cb = engine.chunk_bytes
    n = len(data) // cb
    pos, sv, events = 0, next_sv, []
    while pos < n:
        if all(busy):
            break
        if busy[sv]:
            sv = (sv + 1) % 32
            continue
        free = [c for c in range(num_chans) if not (chan_busy >> c) & 1]
        if not free:
            break
        ch = free[0]
        p = engine.search_blocks(data[pos * cb:(pos + 1) * cb], sv_of_block=[sv])[0]
        pos += 1
        if not (p["snr"] < 25):
            busy[sv] = True
            chan_busy |= 1 << ch
            events.append(dict(chunk_index=pos - 1, sv=sv, ch=ch, snr=float(p["snr"]), lo_shift=int(p["lo_shift"]),
                               ca_shift=int(p["ca_shift"])))
        sv = (sv + 1) % 32
    return events, pos, busy, chan_busy, sv

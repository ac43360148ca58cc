"""CPU-only: host logic, the C ABI's exported symbols, the C++ host tree, and the CPU replay of
the kernel arithmetic (tests/emu) against numpy."""
import ctypes
import re
import subprocess

import numpy as np
import pytest

from conftest import CAPTURES, ROOT, parse_stdout, strip_banner

PKG = ROOT / "gnss-gps-sdr_b200"


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.run(["make", "-s", "lib", "gps_test", "emu"], cwd=ROOT, check=True)


def test_c_abi_exports_every_declared_symbol(ga):
    header = (ROOT / "include" / "gpsacq.h").read_text()
    declared = set(re.findall(r"\b(gpsacq_[a-z0-9_]+)\s*\(", header))
    declared -= {"gpsacq_cfg", "gpsacq_peak", "gpsacq_cell", "gpsacq_info"}
    lib = ctypes.CDLL(str(ga.lib_path()))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/gpsacq.h but not exported"
    from gnss_gps_sdr_b200 import acq
    assert set(acq.ABI_SYMBOLS) == declared
    ga.load_library()


def test_struct_layouts_match_header(ga, tmp_path):
    assert ga.PEAK_DTYPE.itemsize == 32 and ga.CELL_DTYPE.itemsize == 16
    assert ga.PEAK_DTYPE.fields["ca_shift"][1] == 16 and ga.PEAK_DTYPE.fields["flags"][1] == 24
    # the ctypes / numpy mirrors against what a C compiler makes of include/gpsacq.h
    from gnss_gps_sdr_b200 import acq
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gpsacq.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %d\\n",'
                   'sizeof(gpsacq_cfg), sizeof(gpsacq_info), sizeof(gpsacq_peak), sizeof(gpsacq_cell), sizeof(gpsacq_handoff),'
                   'sizeof(gpsacq_event), offsetof(gpsacq_event, start), sizeof(gpsacq_sat), GPSACQ_ABI_VERSION);return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(acq._Cfg), ctypes.sizeof(acq._Info), ga.PEAK_DTYPE.itemsize, ga.CELL_DTYPE.itemsize,
            ga.HANDOFF_DTYPE.itemsize, ga.EVENT_DTYPE.itemsize, ga.EVENT_DTYPE.fields["start"][1], ctypes.sizeof(acq._Sat), 4]
    assert got == want, (got, want)
    assert ctypes.sizeof(acq._Handoff) == ga.HANDOFF_DTYPE.itemsize


def test_no_cpu_fallback(ga, gpu_available):
    if gpu_available:
        pytest.skip("a GPU is present")
    with pytest.raises(ga.GpsAcqError, match="no CUDA device|CUDA"):
        ga.Acquisition(4.092e6, 5.456e6)


def test_product_never_imports_oracle():
    """The oracle is the checker; nothing under the package or include/ may import, link or dlopen it."""
    banned = re.compile(r"import\s+oracle|from\s+oracle|liboracle|libref_harness|gps_test_ref|fft_mixed|libemu")
    files = [p for ext in ("*.py", "*.cu", "*.cuh", "*.h", "*.cpp", "akefile") for p in PKG.rglob(ext)]
    files += [ROOT / "include" / "gpsacq.h", ROOT / "gpsacq_loader.py"]
    assert len(files) > 10
    for p in files:
        assert not banned.search(p.read_text()), p


def test_format_run_reproduces_reference_stdout(ga):
    """SearchTask()'s report (c/search_offline.cpp:264-287) from the reference's own peak records."""
    for name, c in CAPTURES.items():
        ref = np.load(c["peaks"])
        pk = np.zeros(len(ref), ga.PEAK_DTYPE)
        for f in ("snr", "lo_shift", "ca_shift", "sv"):
            pk[f] = ref[f]
        text = "".join(ga.format_run(r, pk[32 * r:32 * r + 32]) for r in range(c["runs"]))
        golden = strip_banner(c["stdout"].read_text())
        assert golden.startswith(text), name          # byte-identical for the runs covered by the fixture


class _FakeAcq:
    """Stands in for the engine to test SearchTask()'s file traversal without a GPU."""
    chunk_bytes = 5120

    def __init__(self, ga, peaks):
        self.ga, self.peaks, self.calls = ga, peaks, []

    def search_blocks(self, bits, sv_of_block=None):
        n = len(bits) // 5120
        start = sum(self.calls)
        self.calls.append(n)
        out = np.zeros(n, self.ga.PEAK_DTYPE)
        for f in ("snr", "lo_shift", "ca_shift", "sv"):
            out[f] = self.peaks[f][start:start + n]
        return out


def test_search_task_traversal(ga, tmp_path):
    c = CAPTURES["nottingham"]
    ref = np.load(c["peaks"])
    data = c["bin"].read_bytes()
    # 3 full runs + a ragged tail: the partial run is discarded with "run out of file!"
    f = tmp_path / "ragged.bin"
    f.write_bytes(data[: 3 * 163840 + 7777])
    fake = _FakeAcq(ga, ref)
    text = ga.search_task_text(fake, str(f), runs_per_batch=2)
    runs, tail = parse_stdout(text)
    assert [r["run"] for r in runs] == [0, 1, 2] and tail == ["run out of file!"]
    assert fake.calls == [64, 32]
    assert strip_banner(c["stdout"].read_text()).startswith(text[: text.index("run out")])
    # exact multiple of a run: still ends with the message (fread returns 0 on the next Sample())
    f.write_bytes(data[: 2 * 163840])
    text = ga.search_task_text(_FakeAcq(ga, ref), str(f), runs_per_batch=2)
    assert text.endswith("run out of file!\n") and len(parse_stdout(text)[0]) == 2
    # empty file and missing file
    f.write_bytes(b"")
    assert ga.search_task_text(_FakeAcq(ga, ref), str(f)) == "run out of file!\n"
    assert ga.search_task_text(_FakeAcq(ga, ref), str(tmp_path / "nope.bin")) == "can not open file!\n"


def test_cxx_host_tree(tmp_path):
    exe = tmp_path / "host_check"
    subprocess.run(["g++", "-O1", "-std=c++17", f"-I{PKG / 'c'}", str(ROOT / "tests/host/host_check.cpp"),
                    str(PKG / "c/search_offline.cpp"), "-o", str(exe), f"-L{PKG / 'csrc'}", "-lgpsacq",
                    f"-Wl,-rpath,{PKG / 'csrc'}", "-lpthread"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "first10 1440 ones 512 g1_after_period 1023 g1_seed 1023"     # IS-GPS-200 PRN 1
    assert out[1] == "first10 1712 ones 512 g1_after_period 1023 g1_seed 1023"     # PRN 32
    sc = [int(v) for v in out[2].split()[1:]]
    assert sc[0] == 0 and 0 < sc[1] < 1023 and sc[2] == -1 and sc[3] == -1


def test_gps_test_cli_contract(gpu_available):
    """Banner, argument rules and exit codes of gps_test (c/test_search_offline.cpp:24-44)."""
    exe = PKG / "c" / "gps_test"
    golden_banner = "\n".join(CAPTURES["nottingham"]["stdout"].read_text().split("\n")[:6]) + "\n"
    r = subprocess.run([str(exe), "a", "b"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == golden_banner + "Please run with 3 arguments or without argument!\n"
    if not gpu_available:
        r = subprocess.run([str(exe), "x.bin", "4.092e6", "5.456e6", "5000"], capture_output=True, text=True)
        assert r.returncode != 0 and r.stdout.startswith(golden_banner) and "SearchInit() returned" in r.stdout


# ---- CPU replay of the kernel arithmetic ---------------------------------------------------------
def _emu():
    L = ctypes.CDLL(str(ROOT / "tests/emu/libemu.so"))
    return L


@pytest.mark.parametrize("gid", [0, 1, 2])
def test_forward_gather_tables_replay(gid):
    """fwd_kernel's table-driven pass-A input (ga_fft3.h fwd_lut_entry / fwd_lut_gather, one table load per group of N1
    one-bit samples) against the sample-by-sample unpack + XOR mix + multiply-add of Sample() (c/search_offline.cpp:143-153)
    on random bits and a random LO table: bit-identical for N1 = 5 and N1 = 4 (same operation order), float rounding
    only for N1 = 10 (two groups of five, summed)."""
    L = _emu()
    n1, n2 = ctypes.c_int(), ctypes.c_int()
    assert L.emu_geom(gid, ctypes.byref(n1), ctypes.byref(n2)) == 0
    N1, N2 = n1.value, n2.value
    rng = np.random.default_rng(100 + gid)
    chunk = rng.integers(0, 256, 5120, dtype=np.uint8)
    lo = rng.integers(0, 4, 40960, dtype=np.uint8)
    up = ctypes.POINTER(ctypes.c_ubyte)
    fp = ctypes.POINTER(ctypes.c_float)
    for s in range(N1):
        a, b = np.zeros(N2, np.complex64), np.zeros(N2, np.complex64)
        assert L.emu_gather(gid, chunk.ctypes.data_as(up), lo.ctypes.data_as(up), s, 0, a.ctypes.data_as(fp)) == 0
        assert L.emu_gather(gid, chunk.ctypes.data_as(up), lo.ctypes.data_as(up), s, 1, b.ctypes.data_as(fp)) == 0
        if N1 <= 5:
            assert a.tobytes() == b.tobytes(), f"s={s}"
        else:
            assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()
        # and against the definition in double
        bits = ((chunk[:, None] >> np.arange(8)) & 1).reshape(-1)[: N1 * N2].astype(np.int64)
        x = np.where(bits ^ (lo[: N1 * N2] & 1), -1.0, 1.0) + 1j * np.where(bits ^ (lo[: N1 * N2] >> 1), -1.0, 1.0)
        z = (x.reshape(N1, N2) * np.exp(-2j * np.pi * np.arange(N1) * s / N1)[:, None]).sum(0)
        assert np.abs(b - z).max() <= 1e-5


# gid 0-2: the three REF geometries (N = 40000); 3: a GRID embedding geometry (N1 = 2, L = 12800); 4: a GRID exact-length
# geometry (N1 = 1, L = W = 4096)
@pytest.mark.parametrize("gid,W", [(0, 5456), (1, 8184), (2, 2800), (3, 5456), (4, 4096)])
def test_kernel_math_replay_matches_numpy(gid, W):
    L = _emu()
    fp = ctypes.POINTER(ctypes.c_float)
    P = lambda a: a.ctypes.data_as(fp)
    n1, n2 = ctypes.c_int(), ctypes.c_int()
    assert L.emu_geom(gid, ctypes.byref(n1), ctypes.byref(n2)) == 0
    N1, N2 = n1.value, n2.value
    N = N1 * N2
    assert N == (40000 if gid < 3 else 12800 if gid == 3 else 4096) and W <= N2
    rng = np.random.default_rng(gid)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    X = np.fft.fft(x.astype(np.complex128))
    for s in (0, N1 - 1):
        out = np.zeros(N2, np.complex64)
        L.emu_fwd(gid, P(x.view(np.float32)), s, P(out.view(np.float32)))
        assert np.abs(out - X[s::N1]).max() <= 2e-6 * np.abs(X).max()
    Xs = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    Cc = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    xd = np.conj(Xs).reshape(N2, N1).T.copy()
    cd = Cc.reshape(N2, N1).T
    cext = np.concatenate([cd, cd], axis=1).copy()
    for dop in ((-36, -1, 0, 5, 36) if gid < 3 else (0, -3, 7)):
        # prod[k] = conj(X[k]) * C[(k - dop) mod N]  (c/search_offline.cpp:181-185), backward FFT (:187)
        y = np.fft.ifft(np.conj(Xs.astype(np.complex128)) * np.roll(Cc.astype(np.complex128), dop)) * N
        yy = np.zeros(N2, np.complex64)
        best, bi, sm = ctypes.c_float(), ctypes.c_int(), ctypes.c_float()
        L.emu_cell(gid, P(xd.view(np.float32)), P(cext.view(np.float32)), dop, W, P(yy.view(np.float32)),
                   ctypes.byref(best), ctypes.byref(bi), ctypes.byref(sm))
        assert np.abs(yy - y[:N2]).max() <= 2e-6 * np.abs(y).max()
        pw = np.abs(y[:W]) ** 2
        assert bi.value == int(pw.argmax())
        assert abs(best.value / pw.max() - 1) < 1e-5 and abs(sm.value / pw.sum() - 1) < 1e-5


def test_segmented_window_replay_matches_numpy():
    """Sampling rates above 10 MHz: the W > 10000 search window is covered by output segments of the 4 x 10000 geometry
    (ga_kernels.cuh c_ktab_seg).  Replay of every segment against the full 40000-point backward FFT."""
    L = _emu()
    fp = ctypes.POINTER(ctypes.c_float)
    P = lambda a: a.ctypes.data_as(fp)
    N1, N2, N, W = 4, 10000, 40000, 36368
    rng = np.random.default_rng(44)
    Xs = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    Cc = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    xd = np.conj(Xs).reshape(N2, N1).T.copy()
    cd = Cc.reshape(N2, N1).T
    cext = np.concatenate([cd, cd], axis=1).copy()
    for dop in (-12, 0, 7):
        y = np.fft.ifft(np.conj(Xs.astype(np.complex128)) * np.roll(Cc.astype(np.complex128), dop)) * N
        for seg in range(4):
            yy = np.zeros(N2, np.complex64)
            best, bi, sm = ctypes.c_float(), ctypes.c_int(), ctypes.c_float()
            L.emu_cell_seg(seg, P(xd.view(np.float32)), P(cext.view(np.float32)), dop, W, P(yy.view(np.float32)),
                           ctypes.byref(best), ctypes.byref(bi), ctypes.byref(sm))
            assert np.abs(yy - y[seg * N2:(seg + 1) * N2]).max() <= 2e-6 * np.abs(y).max()
            pw = np.abs(y[seg * N2:min(W, (seg + 1) * N2)]) ** 2
            assert bi.value == seg * N2 + int(pw.argmax())
            assert abs(best.value / pw.max() - 1) < 1e-5 and abs(sm.value / pw.sum() - 1) < 1e-5


@pytest.mark.parametrize("W", [5456, 8184, 2800])
def test_pfa_kernel_math_replay_matches_numpy(W):
    """Native W-point prime-factor transforms (csrc/ga_pfa.h) replayed per thread on the CPU: forward transform
    into the (a,b,c)-linear order, then product + backward transform + statistics of one GRID cell."""
    L = _emu()
    fp = ctypes.POINTER(ctypes.c_float)
    P = lambda a: a.ctypes.data_as(fp)
    order = np.zeros(W, np.int32)
    assert L.emu_pfa_order(W, order.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
    assert sorted(order.tolist()) == list(range(W))                # Good's map is a permutation
    rng = np.random.default_rng(W)
    x = (rng.standard_normal(W) + 1j * rng.standard_normal(W)).astype(np.complex64)
    c = (rng.standard_normal(W)).astype(np.complex64)
    X = np.fft.fft(x.astype(np.complex128))
    C = np.fft.fft(c.astype(np.complex128))
    xs = np.zeros(W, np.complex64)
    cs = np.zeros(W, np.complex64)
    assert L.emu_pfa_fwd(W, P(x.view(np.float32)), 1, P(xs.view(np.float32))) == 0
    assert L.emu_pfa_fwd(W, P(c.view(np.float32)), 0, P(cs.view(np.float32))) == 0
    assert np.abs(xs - np.conj(X)[order]).max() <= 2e-6 * np.abs(X).max()
    assert np.abs(cs - C[order]).max() <= 2e-6 * np.abs(C).max()
    inv = np.argsort(order)                                        # natural spectral index -> position
    y = np.fft.ifft(xs.astype(np.complex128)[inv] * cs.astype(np.complex128)[inv]) * W     # conj(X)*C, backward
    yy = np.zeros(W, np.complex64)
    best, bi, sm, slow = ctypes.c_float(), ctypes.c_int(), ctypes.c_float(), ctypes.c_int()
    assert L.emu_pfa_cell(W, P(xs.view(np.float32)), P(cs.view(np.float32)), 0, P(yy.view(np.float32)),
                          ctypes.byref(best), ctypes.byref(bi), ctypes.byref(sm), ctypes.byref(slow)) == 0
    assert np.abs(yy - y).max() <= 3e-6 * np.abs(y).max()
    pw = np.abs(y) ** 2
    assert bi.value == int(pw.argmax()) and slow.value == 0
    assert abs(best.value / pw.max() - 1) < 1e-5 and abs(sm.value / pw.sum() - 1) < 1e-5
    # Doppler bins a whole DFT bin apart share one forward transform: the cell of bin r + R*q multiplies X_r with the
    # replica spectrum rotated by -q; its powers equal those of conj(X_r[k+q]) C[k] (the product is only shifted)
    for q in (1, -1, 7, -100, 100):
        pq = np.abs(np.fft.ifft(np.roll(xs.astype(np.complex128)[inv], -q) * cs.astype(np.complex128)[inv]) * W) ** 2
        yr = np.zeros(W, np.complex64)
        assert L.emu_pfa_cell(W, P(xs.view(np.float32)), P(cs.view(np.float32)), q, P(yr.view(np.float32)),
                              ctypes.byref(best), ctypes.byref(bi), ctypes.byref(sm), ctypes.byref(slow)) == 0
        assert np.abs(np.abs(yr.astype(np.complex128)) ** 2 - pq).max() <= 1e-5 * pq.max(), q
        assert bi.value == int(pq.argmax())
    # exact ties: "first maximum wins" (c/search_offline.cpp:192) must not depend on the order the butterflies
    # produce their outputs in.  A flat product spectrum gives one peak at lag 0 and exact zeros elsewhere
    # is not guaranteed in floats, so use the two degenerate inputs whose powers are exactly equal:
    for fill, want in ((0.0, 0), (1.0, 0)):
        xs[:] = fill
        cs[:] = 0.0 if fill == 0.0 else 1.0
        if fill == 1.0:
            cs[:] = 0.0
            cs[np.argsort(order)[0]] = 1.0            # only spectral bin k = 0 set: y[t] = 1 for every lag
        L.emu_pfa_cell(W, P(xs.view(np.float32)), P(cs.view(np.float32)), 0, P(yy.view(np.float32)),
                       ctypes.byref(best), ctypes.byref(bi), ctypes.byref(sm), ctypes.byref(slow))
        assert bi.value == want and slow.value > 0
        assert best.value == fill and sm.value == fill * W


# ---- argument validation happens before any device is touched: checkable without a GPU ---------------------------------
def test_argument_validation_precedes_the_device(ga):
    """Bad configurations are GPSACQ_EINVAL (-1) with a message -- never GPSACQ_ECUDA (-2), which is what a VALID configuration
    gets on a box without a GPU.  REF: FS / fft_len / max_fo limits of c/search_offline.cpp:176,190 and c/gps_offline.h:15;
    GRID: block-length and grid rules of SURVEY App. E; stream converters and generator: argument checks."""
    import ctypes as C
    bad = [dict(fc=4e6, fs=-1.0), dict(fc=4e6, fs=0.0), dict(fc=-1.0, fs=5.456e6), dict(fc=4e6, fs=40.5e6),     # FS/1000 > FFT_LEN
           dict(fc=4e6, fs=5.456e6, max_fo=-5.0), dict(fc=4e6, fs=5.456e6, max_fo=2e6),                         # dmax >= N2
           dict(fc=4e6, fs=float("nan")), dict(fc=4e6, fs=5.456e6, mode=7),
           dict(fc=4e6, fs=5.456e6, mode=1, doppler_step=0.0), dict(fc=4e6, fs=5.4561e6, mode=1, doppler_step=500.0),   # W not a multiple of 8
           dict(fc=4e6, fs=5.456e6, mode=1, doppler_step=333.0),                                                       # FS/step not an integer
           dict(fc=4e6, fs=16.368e6, mode=1, doppler_step=500.0),                                                      # GRID: W > 10000
           dict(fc=4e6, fs=5.456e6, mode=1, doppler_step=500.0, dop_first=20, dop_count=5)]                           # shard outside the grid
    for kw in bad:
        with pytest.raises(ga.GpsAcqError, match=r"failed \(-1\)"):
            ga.Acquisition(**kw)
    lib = ga.load_library()
    assert lib.gpsacq_create(None, None) == -1
    # converters / generator: bad arguments are rejected before the device is selected
    out = (C.c_ubyte * 64)()
    assert lib.gpsacq_bits_to_iq8(0, None, C.c_size_t(8), C.c_size_t(0), C.c_double(2.6e6), C.c_double(10e6), 30, out) == -1
    lib.gpsacq_bits_to_iq8_device.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    assert lib.gpsacq_bits_to_iq8_device(0, C.c_void_p(16), 8, 0, 2.6e6, 10e6, 200, C.c_void_p(16), None) == -1      # amplitude > 127
    assert lib.gpsacq_bits_to_iq8(0, out, C.c_size_t(0), C.c_size_t(0), C.c_double(2.6e6), C.c_double(10e6), 30, out) == 0     # empty input: nothing to do
    with pytest.raises(ga.GpsAcqError):
        ga.synth_capture_gpu(4096, 5.456e6, 4.092e6, [dict(prn=40, doppler_hz=0.0, code_phase_chips=0.0, amp=1.0)])
    with pytest.raises(ga.GpsAcqError):
        ga.synth_capture_gpu(4096, 5.456e6, 4.092e6, [dict(prn=3, doppler_hz=0.0, code_phase_chips=-2.0, amp=1.0)])
    with pytest.raises(ga.GpsAcqError):
        ga.sig_gen_literal(0, np.zeros(4, np.uint8))

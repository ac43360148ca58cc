"""GPU parity tests (run with `-m gpu` on a B200): the CUDA engine, called through the C ABI,
against (a) golden vectors produced by the unmodified reference and (b) the CPU oracle."""
import importlib
import json
import subprocess

import numpy as np
import pytest

from conftest import (CAPTURES, GOLD, ROOT, SNR_RTOL, compare_peaks, compare_runs, parse_stdout, rates_case, rates_golden,
                      strip_banner)

pytestmark = pytest.mark.gpu
PKG = ROOT / "gnss-gps-sdr_b200"


@pytest.fixture(scope="module")
def engines(ga):
    made = {}

    def get(fc, fs, max_fo=5000.0, **kw):
        key = (fc, fs, max_fo, tuple(sorted(kw.items())))
        if key not in made:
            made[key] = ga.Acquisition(fc, fs, max_fo, **kw)
        return made[key]

    yield get
    for a in made.values():
        a.close()


@pytest.fixture(scope="module")
def siggen(ga):
    return importlib.import_module("gnss_gps_sdr_b200.siggen")


# ---- SearchInit(): replicas -------------------------------------------------------------------
@pytest.mark.parametrize("fs", [5.456e6, 8.184e6, 2.8e6, 10e6])
def test_replica_time_bit_exact(engines, oracle_mod, fs):
    """Code NCO + chip-edge blend (c/search_offline.cpp:84-103) is float-for-float identical."""
    acq = engines(fs / 4, fs)
    for sv in range(32):
        a, b = acq.replica_time(sv), oracle_mod.replica_time(fs, sv)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"PRN {sv + 1}"


@pytest.mark.parametrize("name", list(CAPTURES))
def test_replica_spectra_vs_reference(engines, oracle_mod, name):
    c = CAPTURES[name]
    acq = engines(c["fc"], c["fs"])
    p = np.load(c["probe"])
    o = oracle_mod.Oracle(c["fc"], c["fs"])
    for sv in range(32):
        s = acq.replica_spectrum(sv)
        scale = np.abs(s).max()
        assert np.abs(s[p["idx"]] - p["code"][sv]).max() <= 3e-6 * scale          # reference code[sv]
        assert abs(np.abs(s).astype(np.float64).sum() / p["code_abs_sum"][sv] - 1) < 1e-6
        if sv % 8 == 0:
            assert np.abs(s - o.code_spectrum(sv)).max() <= 3e-6 * scale           # full spectrum vs oracle


# ---- Sample(): unpack + mix + forward FFT ----------------------------------------------------------
@pytest.mark.parametrize("name", list(CAPTURES))
def test_block_spectrum_vs_reference(engines, oracle_mod, name):
    c = CAPTURES[name]
    acq = engines(c["fc"], c["fs"])
    data = c["bin"].read_bytes()
    acq.search_blocks(data[: 40 * 5120])
    p = np.load(c["probe"])
    o = oracle_mod.Oracle(c["fc"], c["fs"])
    x0 = acq.block_spectrum(0)
    assert np.abs(x0[p["idx"]] - p["block0"]).max() <= 3e-6 * np.abs(x0).max()    # reference fwd_buf
    for b in (0, 17, 39):
        xo = o.sample(data[b * 5120:(b + 1) * 5120])
        assert np.abs(acq.block_spectrum(b) - xo).max() <= 3e-6 * np.abs(xo).max()


# ---- Correlate(): per-cell statistics ------------------------------------------------------------
@pytest.mark.parametrize("name", list(CAPTURES))
def test_cell_stats_vs_oracle(engines, oracle_mod, name):
    c = CAPTURES[name]
    acq = engines(c["fc"], c["fs"])
    data = c["bin"].read_bytes()[: 32 * 5120]
    acq.search_blocks(data)
    o = oracle_mod.Oracle(c["fc"], c["fs"])
    for b in (0, 7, 20, 31):
        cs = acq.cell_stats(b)
        mp, mi, tp = o.cells(data[b * 5120:(b + 1) * 5120], b)
        assert np.abs(cs["max_pwr"] / mp - 1).max() <= 2e-5
        assert np.abs(cs["tot_pwr"] / tp - 1).max() <= 2e-5
        # argmax exact; a differing index is tolerated only for a genuine float tie
        diff = cs["max_idx"] != mi
        assert diff.sum() <= 1 and np.all(np.abs(cs["max_pwr"][diff] / mp[diff] - 1) < 1e-6)


# ---- whole path vs the reference's own outputs -----------------------------------------------------
@pytest.mark.parametrize("name", list(CAPTURES))
def test_peaks_vs_reference_golden(engines, name):
    """(snr, lo_shift, ca_shift) of every chunk of the fixture vs what the unmodified reference's
    Sample()+Correlate() returned (tests/golden/ref_peaks_*.npy)."""
    c = CAPTURES[name]
    acq = engines(c["fc"], c["fs"])
    got = acq.search_blocks(c["bin"].read_bytes())
    ref = np.load(c["peaks"])
    assert len(got) == len(ref) == c["runs"] * 32
    worst = compare_peaks(got, ref)
    assert worst < 1e-3
    assert np.array_equal(got["sv"], np.arange(len(got)) % 32)


@pytest.mark.parametrize("name", list(CAPTURES))
def test_search_task_text_vs_golden_stdout(ga, engines, name):
    c = CAPTURES[name]
    acq = engines(c["fc"], c["fs"])
    text = ga.search_task_text(acq, str(c["bin"]), runs_per_batch=3)
    got_runs, tail = parse_stdout(text)
    ref_runs, _ = parse_stdout(strip_banner(c["stdout"].read_text()))
    assert tail == ["run out of file!"] and len(got_runs) == c["runs"]
    compare_runs(got_runs, ref_runs[: c["runs"]])


@pytest.mark.parametrize("name", list(CAPTURES))
def test_gps_test_binary_vs_golden_stdout(name):
    """The C++ drop-in `gps_test` (reference CLI) end to end."""
    c = CAPTURES[name]
    subprocess.run(["make", "-s", "gps_test"], cwd=ROOT, check=True)
    r = subprocess.run([str(PKG / "c" / "gps_test"), str(c["bin"]), repr(c["fc"]), repr(c["fs"]), "5000"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    golden = c["stdout"].read_text()
    assert r.stdout.split("\n")[:6] == golden.split("\n")[:6]                       # banner byte-identical
    got_runs, tail = parse_stdout(strip_banner(r.stdout))
    ref_runs, _ = parse_stdout(strip_banner(golden))
    assert tail == ["run out of file!"]
    compare_runs(got_runs, ref_runs[: c["runs"]])


@pytest.mark.parametrize("which", ["fc", "max_fo"])
def test_cxx_host_rereads_the_globals_like_the_reference(tmp_path, which):
    """FC / max_fo changed between SearchInit() and SearchTask() (tests/host/host_globals.cpp): the reference reads them in
    Sample() / Correlate() only (c/search_offline.cpp:127,176), so its output is the golden stdout of the constant-globals
    run; the C++ host here has to notice the change and rebuild its engine."""
    c = CAPTURES["nottingham"]
    exe = tmp_path / "host_globals"
    subprocess.run(["g++", "-O1", "-std=c++17", f"-I{PKG / 'c'}", str(ROOT / "tests/host/host_globals.cpp"),
                    str(PKG / "c/search_offline.cpp"), "-o", str(exe), f"-L{PKG / 'csrc'}", "-lgpsacq",
                    f"-Wl,-rpath,{PKG / 'csrc'}", "-lpthread"], check=True)
    r = subprocess.run([str(exe), str(c["bin"]), which], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got_runs, tail = parse_stdout(r.stdout)
    ref_runs, _ = parse_stdout(strip_banner(c["stdout"].read_text()))
    assert tail == ["run out of file!"] and len(got_runs) == c["runs"]
    compare_runs(got_runs, ref_runs[: c["runs"]])


# ---- the whole bundled capture (north_star acceptance; README.md:61, c/search_offline.cpp:219-292) ----------
FULL_CAPTURE = ROOT / "oracle" / "_ref" / "data" / "gps.samples.1bit.I.fs5456.if4092.bin"      # staged by `make -C oracle`
FULL_PEAKS = GOLD / "ref_peaks_nottingham_full.npy"          # the reference's Sample()+Correlate() on all 10,880 chunks
STRIDED = GOLD / "nottingham_strided_runs.bin"               # whole runs 100-103, 200-203, 336-339 (always in the repo)
RUN_BYTES = 32 * 5120


def _capture_slices():
    """(label, bytes, run numbers): the whole 340-run capture when it travelled with the repository (it is
    git-ignored but NOT gpurun-ignored), and in any case the committed strided runs -- never a skip."""
    runs = json.loads((GOLD / "nottingham_strided_runs.json").read_text())["runs"]
    out = [("strided", STRIDED.read_bytes(), runs)]
    if FULL_CAPTURE.exists():
        out.append(("full", FULL_CAPTURE.read_bytes()[: 340 * RUN_BYTES], list(range(340))))   # + a partial 341st run ("run out of file!")
    return out


def _renumber(runs, numbers):
    return [dict(r, run=numbers[i]) for i, r in enumerate(runs)]


def test_whole_capture_peaks_vs_reference(engines):
    """Every chunk's (snr, lo_shift, ca_shift) against what the UNMODIFIED reference returned for it
    (tests/golden/ref_peaks_nottingham_full.npy): all 10,880 chunks when the capture is staged."""
    c = CAPTURES["nottingham"]
    acq = engines(c["fc"], c["fs"])
    ref_all = np.load(FULL_PEAKS)
    assert len(ref_all) == 340 * 32 and int((ref_all["snr"] >= 25).sum()) == 3582
    for label, data, runs in _capture_slices():
        got = acq.search_blocks(data)
        ref = np.concatenate([ref_all[r * 32:(r + 1) * 32] for r in runs])
        assert len(got) == len(ref)
        worst = compare_peaks(got, ref)
        print(f"whole-capture peaks [{label}]: {len(got)} chunks, {int((ref['snr'] >= 25).sum())} reference hits, worst SNR rel err {worst:.2e}")


def test_marginal_hits_vs_reference(engines):
    """The 103 chunks of the capture whose reference SNR lies in [24, 26] -- every hit and miss hugging the
    SNR-25 threshold (24 hits within +-0.5) -- searched for their own PRN: SNR within 1e-4,
    the same cell wins, and the detection decision agrees outside the 1e-3 band around 25."""
    c = CAPTURES["nottingham"]
    meta = json.loads((GOLD / "nottingham_marginal_chunks.json").read_text())
    data = (GOLD / "nottingham_marginal_chunks.bin").read_bytes()
    idx, sv = np.array(meta["chunk"]), np.array(meta["sv"], np.int32)
    ref = np.load(FULL_PEAKS)[idx]
    assert len(idx) >= 100 and int((np.abs(ref["snr"] - 25) <= 0.5).sum()) >= 24
    got = engines(c["fc"], c["fs"]).search_blocks(data, sv)
    rel = np.abs(got["snr"].astype(np.float64) / ref["snr"] - 1)
    assert rel.max() <= SNR_RTOL, rel.max()
    assert np.array_equal(got["lo_shift"], ref["lo_shift"]) and np.array_equal(got["ca_shift"], ref["ca_shift"])
    decided = np.abs(ref["snr"] / 25.0 - 1) > 1e-3
    assert np.array_equal((got["snr"] >= 25)[decided], (ref["snr"] >= 25)[decided])
    assert np.array_equal(got["flags"] & 1, (got["snr"] >= 25).astype(np.int32))


def test_whole_capture_search_task_text(ga, engines):
    """SearchTask()'s printed report (Python mirror over the C ABI) against the reference's stdout."""
    c = CAPTURES["nottingham"]
    ref_runs, _ = parse_stdout(strip_banner(c["stdout"].read_text()))
    assert len(ref_runs) == 340
    for label, data, runs in _capture_slices():
        f = FULL_CAPTURE if label == "full" else STRIDED
        text = ga.search_task_text(engines(c["fc"], c["fs"]), str(f))
        got_runs, tail = parse_stdout(text)
        assert tail == ["run out of file!"] and len(got_runs) == len(runs)
        compare_runs(_renumber(got_runs, runs), [ref_runs[r] for r in runs])


def test_whole_capture_gps_test_binary():
    """The C++ drop-in `gps_test` (reference CLI) over the whole capture / the strided runs."""
    c = CAPTURES["nottingham"]
    subprocess.run(["make", "-s", "gps_test"], cwd=ROOT, check=True)
    ref_runs, _ = parse_stdout(strip_banner(c["stdout"].read_text()))
    for label, data, runs in _capture_slices():
        f = FULL_CAPTURE if label == "full" else STRIDED
        r = subprocess.run([str(PKG / "c" / "gps_test"), str(f), repr(c["fc"]), repr(c["fs"]), "5000"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        got_runs, tail = parse_stdout(strip_banner(r.stdout))
        assert tail == ["run out of file!"] and len(got_runs) == len(runs)
        compare_runs(_renumber(got_runs, runs), [ref_runs[r] for r in runs])


# ---- batching, ragged sizes, explicit PRN maps --------------------------------------------------------
def test_batching_and_ragged_sizes(ga, engines):
    c = CAPTURES["nottingham"]
    data = c["bin"].read_bytes()
    base = engines(c["fc"], c["fs"]).search_blocks(data)                # 128 chunks, one batch
    small = ga.Acquisition(c["fc"], c["fs"], max_blocks=48)             # forces 48+48+32
    try:
        again = small.search_blocks(data)
        assert again.tobytes() == base.tobytes()                        # bit-identical regardless of batching
        assert len(small.search_blocks(b"")) == 0                       # empty input
        one = small.search_blocks(data[:5120 + 100])                    # ragged tail ignored
        assert len(one) == 1 and one.tobytes() == base[:1].tobytes()
        # explicit PRN map: search chunk 0 for every PRN; entry 0 equals the REF result
        sv = np.arange(32, dtype=np.int32)
        allsv = small.search_blocks(data[:5120] * 32, sv)
        assert allsv[0].tobytes() == base[0].tobytes() and np.array_equal(allsv["sv"], sv)
        with pytest.raises(ga.GpsAcqError):
            small.search_blocks(data[:5120], np.array([32], np.int32))
    finally:
        small.close()


def test_rejects_unsupported_configs(ga):
    with pytest.raises(ga.GpsAcqError, match="fft_len"):
        ga.Acquisition(4e6, 40.5e6)                 # FS/1000 > FFT_LEN: the reference's window would run past its buffer
    with pytest.raises(ga.GpsAcqError):
        ga.Acquisition(4e6, -1.0)


# ---- sampling rates above 10 MHz: W > 10000, search window covered by output segments (c/search_offline.cpp:190 takes any
# FS with FS/1000 <= FFT_LEN; a 16.368 MHz front-end is the common case) -----------------------------------------------
@pytest.mark.parametrize("fs,fc,seed", [(16.368e6, 4.092e6, 13), (25e6, 6.25e6, 14), (40e6, 10e6, 15)])
def test_high_sampling_rates_vs_oracle(engines, oracle_mod, siggen, fs, fc, seed):
    sats = siggen.default_constellation(fs, cn0_dbhz=57.0, seed=seed)
    bits = siggen.synth_capture(40960 * 32, fs, fc, sats, seed=seed)
    acq = engines(fc, fs)
    W = int(np.ceil(fs / 1000))
    assert acq.info["window"] == W and acq.info["n2"] == 10000 and acq.n_doppler == 2 * int(5000.0 * 40000 / fs) + 1
    got = acq.search_blocks(bits)
    compare_peaks(got, rates_golden(rates_case(fs, fc, seed), bits))       # what the UNMODIFIED reference returned for this input
    o = oracle_mod.Oracle(fc, fs)
    ref = o.search_blocks(bits)
    compare_peaks(got, ref)
    assert (ref["snr"] >= 25).sum() >= 6            # (the coherent window shrinks with FS: 40000 samples are 1 ms at 40 MHz)
    for b in (0, 20):                               # per-cell statistics over the whole window (all segments merged)
        cs = acq.cell_stats(b)
        mp, mi, tp = o.cells(bits[b * 5120:(b + 1) * 5120], b)
        assert np.abs(cs["max_pwr"] / mp - 1).max() <= 2e-5 and np.abs(cs["tot_pwr"] / tp - 1).max() <= 2e-5
        assert (cs["max_idx"] != mi).sum() <= 1
    for sv in (0, 31):
        assert np.array_equal(acq.replica_time(sv).view(np.uint32), oracle_mod.replica_time(fs, sv).view(np.uint32))


# ---- other sampling rates, synthetic captures: the reference bundles no capture at these rates, so its records for
# these synthetic inputs were taken by running it unmodified (tests/golden/make_golden_rates.py) ---------------------
# (8 MHz and 4 MHz exercise the wide-window variants of the 8000- and 4000-point geometries: 20 / 10 accumulators per
# butterfly, 448-thread and 128-thread CTA shapes)
@pytest.mark.parametrize("fs,fc,seed", [(2.8e6, 0.62e6, 1575420001), (10e6, 2.6e6, 3), (5.456e6, 4.092e6, 1575420000),
                                        (8.0e6, 2.0e6, 11), (4.0e6, 1.0e6, 12)])
def test_synthetic_vs_oracle(engines, oracle_mod, siggen, fs, fc, seed):
    sats = siggen.default_constellation(fs, seed=seed)
    bits = siggen.synth_capture(40960 * 32, fs, fc, sats, seed=seed)
    acq = engines(fc, fs)
    got = acq.search_blocks(bits)
    # the unmodified reference's records for this very input (tests/golden/ref_peaks_rates.npz), then the C restatement
    compare_peaks(got, rates_golden(rates_case(fs, fc, seed), bits))
    ref = oracle_mod.Oracle(fc, fs).search_blocks(bits)
    compare_peaks(got, ref)
    for s in sats:                      # every generated satellite is found at its Doppler bin
        p = got[s["prn"] - 1]
        assert p["snr"] >= 25 and abs(p["lo_shift"] - s["doppler_hz"] * 40000 / fs) <= 1.0


def test_max_fo_changes_the_doppler_grid(engines, oracle_mod, siggen):
    fs, fc = 5.456e6, 4.092e6
    sats = [dict(prn=9, doppler_hz=7300.0, code_phase_chips=100.25, amp=0.3)]
    bits = siggen.synth_capture(40960 * 9, fs, fc, sats, noise_sigma=1.0, seed=5)
    acq = engines(fc, fs, 10000.0)
    assert acq.n_doppler == 2 * int(10000.0 * 40000 / fs) + 1
    got = acq.search_blocks(bits)
    compare_peaks(got, rates_golden("maxfo_10000", bits))          # the unmodified reference with max_fo = 10000 on this input
    ref = oracle_mod.Oracle(fc, fs, 10000.0).search_blocks(bits)
    compare_peaks(got, ref)
    assert got[8]["lo_shift"] == round(7300.0 * 40000 / fs)


# ---- size-independent properties at full batch size ------------------------------------------------------
def test_full_batch_properties(engines, siggen):
    """512 chunks x 73 bins: identical (chunk, PRN) pairs give bit-identical records wherever they
    sit in the batch; the complement of a capture (all bits flipped = signal negated) gives the
    same powers; and records do not depend on neighbours."""
    c = CAPTURES["nottingham"]
    data = np.frombuffer(c["bin"].read_bytes(), np.uint8).reshape(128, 5120)
    rng = np.random.default_rng(11)
    pick = rng.integers(0, 128, 512)
    sv = rng.integers(0, 32, 512).astype(np.int32)
    acq = engines(c["fc"], c["fs"])
    got = acq.search_blocks(np.ascontiguousarray(data[pick]).reshape(-1), sv)
    key = {}
    for i in range(512):
        k = (int(pick[i]), int(sv[i]))
        if k in key:
            j = key[k]
            assert got[i].tobytes() == got[j].tobytes()
        key[k] = i
    neg = acq.search_blocks(np.ascontiguousarray(~data[pick[:64]]).reshape(-1), sv[:64])
    assert np.array_equal(neg["lo_shift"], got["lo_shift"][:64]) and np.array_equal(neg["ca_shift"], got["ca_shift"][:64])
    assert np.allclose(neg["snr"], got["snr"][:64], rtol=1e-5)


# ---- the cell kernel's variants: work queue vs round robin, operands by __ldg vs staged by TMA -----------------
def test_ticket_scheduler_matches_round_robin(ga, monkeypatch):
    """Cells drawn from the device-wide ticket counter (default) or dealt round robin (GPSACQ_STATIC_SCHED=1) are the
    same cells: byte-identical records, over several back-to-back launches (the counter rewinds itself)."""
    c = CAPTURES["nottingham"]
    data = c["bin"].read_bytes()
    with ga.Acquisition(c["fc"], c["fs"], max_blocks=128) as acq:
        a = [acq.search_blocks(data).copy() for _ in range(3)]
    monkeypatch.setenv("GPSACQ_STATIC_SCHED", "1")
    with ga.Acquisition(c["fc"], c["fs"], max_blocks=128) as acq:
        b = acq.search_blocks(data).copy()
    for x in a:
        assert x.tobytes() == b.tobytes()


def test_tma_staged_kernel_matches_ldg_kernel(ga, monkeypatch):
    """GPSACQ_CELL_TMA=1 (ga_cell_tma.cuh: cp.async.bulk.tensor + mbarrier ring, replica spectra in the halo layout)
    against the default kernel on the Nottingham fixture and on ragged batch sizes: same integers, per-cell powers
    within 2e-6 (only the summation order of a thread's power sums differs), and the reference's golden returns."""
    c = CAPTURES["nottingham"]
    data = c["bin"].read_bytes()
    with ga.Acquisition(c["fc"], c["fs"], max_blocks=128) as acq:
        ref = acq.search_blocks(data).copy()
        ref_cells = acq.cell_stats(5).copy()
    monkeypatch.setenv("GPSACQ_CELL_TMA", "1")
    with ga.Acquisition(c["fc"], c["fs"], max_blocks=128) as acq:
        assert acq.info["cell_threads"] == 256                     # 7 consumer warps + the producer warp
        got = acq.search_blocks(data).copy()
        cells = acq.cell_stats(5).copy()
        small = [acq.search_blocks(data[: n * 5120]).copy() for n in (1, 3, 37)]
    same = (got["lo_shift"] == ref["lo_shift"]) & (got["ca_shift"] == ref["ca_shift"])
    assert same[ref["snr"] >= 20].all() and same.mean() > 0.98       # (a noise-only chunk may flip between two bins whose snr ties to 1e-7)
    assert np.array_equal(got["sv"], ref["sv"]) and np.array_equal(got["flags"], ref["flags"])
    assert np.allclose(got["snr"], ref["snr"], rtol=2e-6) and np.allclose(got["max_pwr"][same], ref["max_pwr"][same], rtol=2e-6)
    assert np.array_equal(cells["max_idx"], ref_cells["max_idx"])
    assert np.allclose(cells["tot_pwr"], ref_cells["tot_pwr"], rtol=2e-6)
    for n, s_ in zip((1, 3, 37), small):
        assert s_.tobytes() == got[:n].tobytes()
    gold = np.load(c["peaks"])
    compare_peaks(got, gold)


def test_pinned_and_pageable_host_buffers_agree(ga, engines):
    """gpsacq_search_blocks() hands a page-locked caller buffer straight to the copy engine and stages pageable memory
    through its own pinned buffer; the batch is cut into 1/8, 1/8, 1/4, 1/2 slices either way.  Same records, also
    for an unaligned view into the pinned buffer and a ragged chunk count."""
    import torch
    c = CAPTURES["nottingham"]
    raw = np.frombuffer(c["bin"].read_bytes(), np.uint8)
    data = np.concatenate([raw, raw, raw])[: 331 * 5120]            # 331 chunks: four unequal slices
    acq = engines(c["fc"], c["fs"])
    want = acq.search_blocks(data.copy()).copy()
    pinned = torch.empty(data.size + 5120, dtype=torch.uint8).pin_memory()
    view = pinned.numpy()[5120:]
    view[:] = data
    got = acq.search_blocks(view)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(want["sv"], np.arange(331) % 32)


# ---- device-pointer API on torch's stream ---------------------------------------------------------------
def test_device_api_on_torch_stream(ga, engines):
    import torch
    c = CAPTURES["nottingham"]
    acq = engines(c["fc"], c["fs"])
    data = c["bin"].read_bytes()[: 64 * 5120]
    host = acq.search_blocks(data)
    dev = torch.device("cuda", acq.info["device"])
    bits = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    out = torch.zeros(64 * 32, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        acq.set_stream(stream.cuda_stream)
        acq.search_blocks_device(bits.data_ptr(), 64, None, out.data_ptr())
    stream.synchronize()
    acq.set_stream(None)
    got = np.frombuffer(out.cpu().numpy().tobytes(), ga.PEAK_DTYPE)
    assert got.tobytes() == host.tobytes()
    t = acq.stage_times()
    assert t["cells_ms"] > 0 and t["total_ms"] >= t["cells_ms"]


# ---- several GPUs in one process: chunk ranges per device + ncclAllGather of the peak records ----------------
@pytest.mark.parametrize("use_nccl", [True, False])
def test_group_api_matches_single_gpu(ga, engines, use_nccl):
    import torch
    n = min(torch.cuda.device_count(), 2)
    c = CAPTURES["nottingham"]
    data = c["bin"].read_bytes()
    base = engines(c["fc"], c["fs"]).search_blocks(data)
    grp = ga.AcquisitionGroup(c["fc"], c["fs"], n_gpus=n, use_nccl=use_nccl, max_blocks=40)   # 128 chunks -> 2 batches
    try:
        assert grp.gather_kind == ("nccl" if (use_nccl and n > 1) else "host")
        got = grp.search_blocks(data)
        assert got.tobytes() == base.tobytes()
        assert len(grp.search_blocks(data[: 3 * 5120])) == 3        # fewer chunks than GPUs*... ragged shares
    finally:
        grp.close()

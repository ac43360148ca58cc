"""CPU-only: pins the oracle (C restatement) to the reference's golden vectors."""
import subprocess
import sys

import numpy as np
import pytest

from conftest import CAPTURES, ROOT, compare_peaks, parse_stdout, strip_banner, compare_runs

# IS-GPS-200 Table 3-Ia, "First 10 chips octal C/A", PRN 1..32
IS_GPS_200_FIRST10 = [0o1440, 0o1620, 0o1710, 0o1744, 0o1133, 0o1455, 0o1131, 0o1454, 0o1626, 0o1504, 0o1642,
                      0o1750, 0o1764, 0o1772, 0o1775, 0o1776, 0o1156, 0o1467, 0o1633, 0o1715, 0o1746, 0o1763,
                      0o1063, 0o1706, 0o1743, 0o1761, 0o1770, 0o1774, 0o1127, 0o1453, 0o1625, 0o1712]


def test_cacode_known_answers(oracle_mod):
    for sv in range(32):
        chips = oracle_mod.cacode_chips(sv)
        first10 = int("".join(str(int(b)) for b in chips[:10]), 2)
        assert first10 == IS_GPS_200_FIRST10[sv], f"PRN {sv + 1}"
        assert int(chips.sum()) == 512          # balanced Gold code: 512 ones, 511 zeros
    # distinct codes, period exactly 1023 (autocorrelation of a Gold code is {-1, 63, -65, 1023})
    c = 1 - 2 * oracle_mod.cacode_chips(0).astype(int)
    ac = np.array([np.dot(c, np.roll(c, k)) for k in range(1023)])
    assert ac[0] == 1023 and set(ac[1:]) <= {-1, 63, -65}


def test_search_code_roundtrip(oracle_mod):
    # SearchCode(sv, g1): chips until the G1 register equals g1 (c/search_offline.cpp:205-209).
    assert oracle_mod.search_code(0, 0x3FF) == 0            # all-ones seed
    seen = {oracle_mod.search_code(5, g1) for g1 in range(1, 1024)}
    assert seen == set(range(1023))                           # G1 is maximal length: every state once
    assert oracle_mod.search_code(5, 0) == -1                 # unreachable state (reference would spin)


def test_replica_nco_facts(oracle_mod):
    # at 8.184 MHz chip edges land on ca_phase = 0 -> replica is exactly +-1
    r = oracle_mod.replica_time(8.184e6, 7)
    assert set(np.unique(r)) == {-1.0, 1.0}
    # at 5.456 MHz ca_rate = 3/16 exactly: 40000 samples advance exactly 7500 chips
    r = oracle_mod.replica_time(5.456e6, 0)
    assert np.abs(r).max() <= 1.0 and (np.abs(r) < 1.0).any()
    lo = oracle_mod.lo_table(4.092e6, 5.456e6)
    assert np.array_equal(lo, (3 * np.arange(40960)) % 4)     # lo_rate = 3.0 exactly


@pytest.mark.parametrize("name", list(CAPTURES))
@pytest.mark.parametrize("fft_f64", [True, False])
def test_oracle_matches_reference_peaks(oracle_mod, name, fft_f64):
    c = CAPTURES[name]
    ref = np.load(c["peaks"])
    o = oracle_mod.Oracle(c["fc"], c["fs"], fft_f64=fft_f64)
    nb = 64 if fft_f64 else 32                               # keep the CPU suite short
    got = o.search_blocks(c["bin"].read_bytes()[: nb * 5120])
    compare_peaks(got, ref[:nb], snr_rtol=2e-5)


@pytest.mark.parametrize("name", list(CAPTURES))
def test_oracle_matches_reference_spectra(oracle_mod, name):
    c = CAPTURES[name]
    p = np.load(c["probe"])
    o = oracle_mod.Oracle(c["fc"], c["fs"])
    for sv in range(32):
        s = o.code_spectrum(sv)
        assert np.abs(s[p["idx"]] - p["code"][sv]).max() <= 2e-6 * np.abs(s).max()
        assert abs(np.abs(s).astype(np.float64).sum() / p["code_abs_sum"][sv] - 1) < 1e-6
    x = o.sample(c["bin"].read_bytes()[:5120])
    assert np.abs(x[p["idx"]] - p["block0"]).max() <= 2e-6 * np.abs(x).max()


@pytest.mark.parametrize("name", list(CAPTURES))
def test_oracle_reproduces_golden_stdout(oracle_mod, ga, name):
    """Oracle -> SearchTask() report formatter (host logic) -> compare with gps_test's own stdout."""
    c = CAPTURES[name]
    ref_runs, tail = parse_stdout(strip_banner(c["stdout"].read_text()))
    assert tail == ["run out of file!"]
    o = oracle_mod.Oracle(c["fc"], c["fs"])
    pk = o.search_blocks(c["bin"].read_bytes()[: 2 * 32 * 5120])
    text = "".join(ga.format_run(r, pk[32 * r: 32 * r + 32]) for r in range(2))
    got_runs, _ = parse_stdout(text)
    compare_runs(got_runs, ref_runs[:2])


def test_golden_stdout_is_complete():
    runs, _ = parse_stdout(strip_banner(CAPTURES["nottingham"]["stdout"].read_text()))
    assert len(runs) == 340 and runs[-1]["run"] == 339      # 55,791,616 B / 163,840 B
    assert runs[0]["ca"] == [1057, 4848, 4963, 5029, 478, 1148, 3283, 1331, 3005, 4049]
    runs, _ = parse_stdout(strip_banner(CAPTURES["gps_sig"]["stdout"].read_text()))
    assert len(runs) == 12 and 7 in runs[0]["sv"]


def test_full_capture_goldens_agree_with_each_other(ga):
    """tests/golden/ref_peaks_nottingham_full.npy (the reference's Sample()+Correlate() returns for all 10,880 chunks,
    made by make_golden_full.py) formatted like SearchTask() prints them reproduces the reference's own stdout
    (nottingham_full.stdout.txt, made by gps_test_ref) over all 340 runs; the strided / marginal fixtures are the
    slices of the capture their json files say."""
    import json
    from conftest import GOLD
    pk = np.load(GOLD / "ref_peaks_nottingham_full.npy")
    assert len(pk) == 340 * 32 and np.array_equal(pk["sv"], np.arange(len(pk)) % 32)
    assert int((pk["snr"] >= 25).sum()) == 3582 and int((np.abs(pk["snr"] - 25) <= 0.5).sum()) >= 24
    ref_runs, _ = parse_stdout(strip_banner(CAPTURES["nottingham"]["stdout"].read_text()))
    full = np.zeros(len(pk), ga.PEAK_DTYPE)
    for k in ("snr", "lo_shift", "ca_shift", "sv"):
        full[k] = pk[k]
    text = "".join(ga.format_run(r, full[32 * r: 32 * r + 32]) for r in range(340))
    assert text == strip_banner(CAPTURES["nottingham"]["stdout"].read_text()).replace("run out of file!\n", "")
    first = np.load(CAPTURES["nottingham"]["peaks"])             # made with the built-in FFT behind the shim, this one with MKL
    assert all(np.array_equal(pk[k][:128], first[k]) for k in ("lo_shift", "ca_shift", "sv"))
    assert np.abs(pk["snr"][:128] / first["snr"] - 1).max() < 2e-6
    runs = json.loads((GOLD / "nottingham_strided_runs.json").read_text())["runs"]
    assert (GOLD / "nottingham_strided_runs.bin").stat().st_size == len(runs) * 32 * 5120
    m = json.loads((GOLD / "nottingham_marginal_chunks.json").read_text())
    assert (GOLD / "nottingham_marginal_chunks.bin").stat().st_size == len(m["chunk"]) * 5120
    assert np.all((pk["snr"][m["chunk"]] >= 24) & (pk["snr"][m["chunk"]] <= 26))
    # the first strided run is run 100 of the capture: its first chunk is the one the marginal list may also hold
    assert runs[:4] == [100, 101, 102, 103]


def test_oracle_on_threshold_hugging_chunks(oracle_mod):
    """The C restatement against the reference on chunks whose SNR hugs the threshold (a sample of the 103 with
    reference SNR in [24, 26]): same cell, SNR within 2e-5."""
    import json
    from conftest import GOLD
    m = json.loads((GOLD / "nottingham_marginal_chunks.json").read_text())
    data = (GOLD / "nottingham_marginal_chunks.bin").read_bytes()
    ref = np.load(GOLD / "ref_peaks_nottingham_full.npy")[m["chunk"]]
    pick = np.argsort(np.abs(ref["snr"] - 25))[:12]                 # the 12 closest to 25
    o = oracle_mod.Oracle(4.092e6, 5.456e6)
    got = o.search_blocks(b"".join(data[i * 5120:(i + 1) * 5120] for i in pick), np.array(m["sv"], np.int32)[pick])
    assert np.array_equal(got["lo_shift"], ref["lo_shift"][pick]) and np.array_equal(got["ca_shift"], ref["ca_shift"][pick])
    assert np.abs(got["snr"] / ref["snr"][pick] - 1).max() < 2e-5


def test_live_reference_agrees_with_oracle(oracle_mod):
    """When oracle/_ref was built here (or travelled prebuilt), drive the real reference TU."""
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref/libref_harness.so not built (needs /root/reference)")
    code = (
        "import sys; sys.path.insert(0, %r); import numpy as np, oracle\n"
        "d = open(%r,'rb').read()[:8*5120]\n"
        "r = oracle.RefHarness(4.092e6, 5.456e6); o = oracle.Oracle(4.092e6, 5.456e6)\n"
        "sv = np.array([0, 4, 12, 15, 20, 22, 24, 28], np.int32)\n"
        "a = r.search_blocks(d, sv); b = o.search_blocks(d, sv)\n"
        "assert np.array_equal(a['lo_shift'], b['lo_shift']) and np.array_equal(a['ca_shift'], b['ca_shift'])\n"
        "assert np.abs(a['snr']/b['snr']-1).max() < 1e-5\n"
        "assert r.search_code(3, 0x2AA) == oracle.search_code(3, 0x2AA)\n"
    ) % (str(ROOT / "oracle"), str(CAPTURES["nottingham"]["bin"]))
    subprocess.run([sys.executable, "-c", code], check=True)


@pytest.mark.parametrize("fs,fc,max_fo", [(2.8e6, 0.62e6, 5000.0),        # rtl-sdr rate: inexact float LO step, 143 bins, 10 x 4000 geometry
                                          (10e6, 2.6e6, 5000.0),           # the receiver's own rate (c/gps.h:23-24)
                                          (16.368e6, 4.1304e6, 5000.0),    # W = 16368 > 10000: two output segments on the GPU
                                          (4e6, 1e6, 5000.0),
                                          (5.456e6, 4.092e6, 2000.0),      # another max_fo: dmax = (int)(max_fo*N/FS) (:176)
                                          (8.184e6, 2.046e6, 12345.6),
                                          (8e6, 2e6, 5000.0), (16.368e6, 4.092e6, 5000.0), (25e6, 6.25e6, 5000.0), (40e6, 10e6, 5000.0)])
def test_live_reference_agrees_with_oracle_at_other_rates(oracle_mod, ga, fs, fc, max_fo):
    """The GPU tests at these rates (test_synthetic_vs_oracle, test_high_sampling_rates_vs_oracle,
    test_max_fo_changes_the_doppler_grid) compare the engine with the C restatement; here the restatement is pinned to the
    UNMODIFIED reference TU at the same rates: synthetic capture, eight chunks searched for the satellites that are in it.
    (One reference instance per process: it keeps its state in file statics.)"""
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref/libref_harness.so not built (needs /root/reference)")
    code = (
        "import sys, importlib; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np, oracle, gpsacq_loader\n"
        "gpsacq_loader.load(); sg = importlib.import_module('gnss_gps_sdr_b200.siggen')\n"
        "fs, fc, max_fo = %r, %r, %r\n"
        "sats = sg.default_constellation(fs, cn0_dbhz=48.0 if fs <= 10e6 else 57.0, seed=11, max_doppler=0.9 * min(max_fo, 4500.0))\n"
        "d = sg.synth_capture(8 * 40960, fs, fc, sats, seed=5).tobytes()\n"
        "sv = np.array([s['prn'] - 1 for s in sats], np.int32)\n"
        "r = oracle.RefHarness(fc, fs, max_fo); o = oracle.Oracle(fc, fs, max_fo)\n"
        "a = r.search_blocks(d, sv); b = o.search_blocks(d, sv)\n"
        "assert (a['snr'] >= 25).sum() >= 6, a['snr']\n"
        "assert np.array_equal(a['lo_shift'], b['lo_shift']) and np.array_equal(a['ca_shift'], b['ca_shift'])\n"
        "assert np.abs(a['snr'] / b['snr'] - 1).max() < 1e-5\n"
        "for s_, p in zip(sats, a): assert p['snr'] < 25 or abs(p['lo_shift'] - s_['doppler_hz'] * 40000 / fs) <= 1.0\n"
    ) % (str(ROOT / "oracle"), str(ROOT), fs, fc, max_fo)
    subprocess.run([sys.executable, "-c", code], check=True, env=oracle_mod.mkl_env())


@pytest.mark.parametrize("name", ["syn_2800", "syn_10000", "syn_5456", "syn_8000", "syn_4000", "hi_16368", "hi_25000", "hi_40000",
                                  "maxfo_10000"])
def test_oracle_matches_reference_goldens_at_other_rates(oracle_mod, ga, name):
    """tests/golden/ref_peaks_rates.npz holds what the UNMODIFIED reference returned for the synthetic captures the GPU tests
    search at 2.8 ... 40 MHz (32 chunks each, all 32 PRNs): the C restatement reproduces it -- integers, and SNR far inside
    the tolerance the GPU tests then apply to the engine against the same golden."""
    import importlib
    import json
    from conftest import GOLD, compare_peaks, rates_golden
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    c = json.loads((GOLD / "ref_peaks_rates.json").read_text())["cases"][name]
    sats = c.get("sats") or sg.default_constellation(c["fs"], cn0_dbhz=c["cn0"], seed=c["seed"])
    bits = sg.synth_capture(40960 * c["chunks"], c["fs"], c["fc"], sats, seed=c["seed"])
    ref = rates_golden(name, bits)
    got = oracle_mod.Oracle(c["fc"], c["fs"], c["max_fo"]).search_blocks(bits.tobytes())
    assert compare_peaks(got, ref, snr_rtol=2e-5) < 2e-4
    assert np.array_equal(got["lo_shift"], ref["lo_shift"]) and np.array_equal(got["ca_shift"], ref["ca_shift"])


# ---- GRID mode: the C oracle against an independent numpy restatement of the definition -------
@pytest.mark.parametrize("fs,fc,step,K", [(2.8e6, 0.62e6, 250.0, 2), (5.456e6, 4.092e6, 500.0, 1)])
def test_grid_oracle_matches_numpy_definition(oracle_mod, ga, fs, fc, step, K):
    """The reference has no GRID mode ("parity unpinned by the reference"); what CAN be pinned is that the C oracle
    computes the written definition: mixer and LO table as Sample(), wipe-off exp(-j2pi((d n) mod M)/M), one-period
    replica through the SearchInit() code NCO, y = IFFT_W(conj(FFT_W(x_d)) FFT_W(c_p)) unnormalised, powers summed
    over the K blocks, first max / sum / snr / ascending strictly-greater scan as Correlate()."""
    import importlib
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    W, max_fo = int(round(fs / 1000)), 2000.0
    M = int(round(fs / step))
    sats = sg.default_constellation(fs, cn0_dbhz=50.0, seed=4, max_doppler=1800.0)
    bits = sg.synth_capture(W * K, fs, fc, sats, seed=8)
    g = oracle_mod.GridOracle(fc, fs, max_fo, step, K)
    ref, cells = g.acquire(bits, want_cells=True)
    cm, ci, ct = cells[0]
    dmax = int(max_fo // step)
    assert g.n_doppler == 2 * dmax + 1 and g.window == W
    b = np.unpackbits(np.frombuffer(bits, np.uint8), bitorder="little").astype(np.int64).reshape(K, W)
    k = oracle_mod.lo_table(fc, fs, W).astype(np.int64)                    # int(lo_phase) per sample (pinned by REF mode)
    lo_sin, lo_cos = np.array([1, 1, 0, 0]), np.array([0, 1, 1, 0])        # c/search_offline.cpp:124-125
    x = (1 - 2 * (b ^ lo_cos[k])) + 1j * (1 - 2 * (b ^ lo_sin[k]))         # Bipolar(bit ^ lo_cos) + j Bipolar(bit ^ lo_sin)
    n = np.arange(W)
    for prn in (sats[0]["prn"], sats[3]["prn"], 2):
        C = np.fft.fft(oracle_mod.replica_time(fs, prn - 1, W).astype(np.float64))
        best = (0.0, 0, 0)
        for di, d in enumerate(range(-dmax, dmax + 1)):
            wipe = np.exp(-2j * np.pi * ((d * n) % M) / M)
            P = np.zeros(W)
            for k in range(K):
                y = np.fft.ifft(np.conj(np.fft.fft(x[k] * wipe)) * C) * W
                P += np.abs(y) ** 2
            assert abs(cm[prn - 1, di] / P.max() - 1) < 2e-5 and abs(ct[prn - 1, di] / P.sum() - 1) < 2e-5
            if P.max() > 1.001 * np.partition(P, -2)[-2]:                  # a clear maximum: the index must agree
                assert ci[prn - 1, di] == int(P.argmax())
            snr = P.max() / (P.sum() / W)
            if snr > best[0]:
                best = (snr, d, int(P.argmax()))
        r = ref[prn - 1]
        assert abs(r["snr"] / best[0] - 1) < 2e-5
        if best[0] >= 25:
            assert (int(r["lo_shift"]), int(r["ca_shift"])) == best[1:]


# ---- GRID mode pinned to what the reference holds (CPU: the oracle's definition; the GPU twins are in test_gpu_grid.py) ----
def test_grid_oracle_agrees_with_the_reference_on_the_capture(oracle_mod):
    """Every detection of the UNMODIFIED reference in run 0 of the capture (10 SVs) is found by the GRID definition
    (7 non-coherent 1 ms blocks of the same chunk, 500 Hz bins) at the reference's code phase (+-1 sample) and within
    one bin of its Doppler -- SURVEY App. D's acceptance criterion for GRID vs REF."""
    c = CAPTURES["nottingham"]
    data = c["bin"].read_bytes()
    ref = np.load(c["peaks"])[:32]
    g = oracle_mod.GridOracle(c["fc"], c["fs"], 5000.0, 500.0, 7)
    hits = np.nonzero(ref["snr"] >= 25)[0]
    assert len(hits) == 10
    for i in hits:
        p = g.acquire(data[i * 5120: i * 5120 + 7 * 682], svs=[i % 32])[0]
        d = abs(int(p["ca_shift"]) - int(ref[i]["ca_shift"]))
        assert min(d, 5456 - d) <= 1 and abs(p["lo_shift"] * 500.0 - ref[i]["lo_shift"] * c["fs"] / 40000) <= 500.0


def test_grid_definition_equals_the_reference_where_the_two_coincide(oracle_mod, ga):
    """At FS = 40 MHz the reference's 40000-sample window IS one 1 ms block, and with doppler_step = FS/N = 1 kHz its
    spectrum rotation by whole bins IS the time-domain wipe-off of the GRID definition: GRID mode (SURVEY App. E -- new
    semantics) and Correlate() then describe the same computation.  The GRID oracle run that way reproduces the UNMODIFIED
    reference's records for the 40 MHz golden input (tests/golden/ref_peaks_rates.npz, case hi_40000): every Doppler bin and
    code phase, SNR to 1e-6 -- mixer, LO restart, replica NCO, wipe-off sign, conjugate on the data, statistics and the
    best-over-Doppler scan of the definition are thereby pinned to the reference itself."""
    import importlib
    import json
    from conftest import GOLD, rates_golden
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    c = json.loads((GOLD / "ref_peaks_rates.json").read_text())["cases"]["hi_40000"]
    sats = sg.default_constellation(c["fs"], cn0_dbhz=c["cn0"], seed=c["seed"])
    bits = sg.synth_capture(40960 * c["chunks"], c["fs"], c["fc"], sats, seed=c["seed"])
    ref = rates_golden("hi_40000", bits)
    g = oracle_mod.GridOracle(c["fc"], c["fs"], c["max_fo"], c["fs"] / 40000.0, 1)
    assert g.window == 40000 and g.n_doppler == 2 * int(c["max_fo"] * 40000 / c["fs"]) + 1
    chunks = bits.reshape(c["chunks"], 5120)
    got = np.concatenate([g.acquire(chunks[b, :5000], svs=[b % 32]) for b in range(c["chunks"])])   # samples 40000.. are discarded there too
    assert (ref["snr"] >= 25).sum() >= 6
    assert np.array_equal(got["lo_shift"], ref["lo_shift"]) and np.array_equal(got["ca_shift"], ref["ca_shift"])
    assert np.abs(got["snr"].astype(np.float64) / ref["snr"] - 1).max() < 1e-6
    # a 500 Hz grid (wipe-off period M = FS/step = 80000 != W): its EVEN bins are the reference's bins; the reference's
    # scan (ascending, strictly greater, :173,:198) over the per-bin statistics of those bins gives the same records
    g2 = oracle_mod.GridOracle(c["fc"], c["fs"], c["max_fo"], c["fs"] / 80000.0, 1)
    assert g2.n_doppler == 2 * g.n_doppler - 1
    for b in range(0, c["chunks"], 3):
        _, cells = g2.acquire(chunks[b, :5000], want_cells=True, svs=[b % 32])
        cm, ci, ct = cells[0]
        best = (np.float32(0), 0, 0)
        for di in range(0, g2.n_doppler, 2):
            snr = np.float32(cm[0, di]) / (np.float32(ct[0, di]) / np.float32(g2.window))
            if snr > best[0]:
                best = (snr, (di - g2.dmax) // 2, int(ci[0, di]))
        assert (best[1], best[2]) == (ref["lo_shift"][b], ref["ca_shift"][b]) and abs(float(best[0]) / float(ref["snr"][b]) - 1) < 1e-6


def test_grid_oracle_vs_the_dataset_page_known_answer(oracle_mod):
    """250 Hz bins on the first 20 ms of the capture: PRN 1/21/29/30/31 within +-1 bin of lo_shift 6/8/-9/-9/-8
    ("Raw GPS signal samples data set for testing GPS receivers.html", Holme's search on this file)."""
    c = CAPTURES["nottingham"]
    known = {0: 6, 20: 8, 28: -9, 29: -9, 30: -8}
    g = oracle_mod.GridOracle(c["fc"], c["fs"], 5000.0, 250.0, 10)
    out = g.acquire(c["bin"].read_bytes()[: 2 * 6820], svs=sorted(known)).reshape(2, 5)
    for a in range(2):
        for n, sv in enumerate(sorted(known)):
            assert out[a, n]["snr"] >= 25 and abs(int(out[a, n]["lo_shift"]) - known[sv]) <= 1


# ---- "next" rows with reference-made artifacts: f1 reverse converter, f2 signal generator --------------------------
def test_conv_1bit_oracle_vs_the_reference_program(oracle_mod):
    """c/conv_1bit_bin_to_hackrf_bin.cpp restated (oracle_conv_1bit_iq8) against the SHA-256 of what the UNMODIFIED program
    (oracle/_ref/conv_1bit_ref, FC/FS from c/gps.h) wrote for the prefix of the capture that is the 4-run fixture."""
    import hashlib, json
    from conftest import GOLD
    g = json.loads((GOLD / "f1f2_golden.json").read_text())["conv_1bit_bin_to_hackrf_bin"]
    out = oracle_mod.conv_1bit_iq8((GOLD / g["fixture"]).read_bytes(), g["fc"], g["fs"], g["amplitude"])
    assert out[:32].tolist() == g["first_32_out_bytes"] and set(np.unique(out)) == {-30, 30}
    assert hashlib.sha256(out.tobytes()).hexdigest() == g["sha256_fixture_prefix"]


def test_sig_gen_oracle_reproduces_the_bundled_file(oracle_mod):
    """gps_sig_gen.m:8-41 restated in double with MATLAB's operation order, fed the NAV bits that the script drew,
    reproduces the reference's bundled gps_sig_tmp.bin bit for bit (SHA-256 of all 2,046,006 bytes; the committed
    fixture is its first 327,680 bytes)."""
    import hashlib, json
    from conftest import GOLD
    g = json.loads((GOLD / "f1f2_golden.json").read_text())["gps_sig_gen"]
    out = oracle_mod.sig_gen_literal(g["prn"] - 1, g["nav_bits01"])
    assert out.size == g["n_bytes"] and hashlib.sha256(out.tobytes()).hexdigest() == g["sha256_file"]
    fx = CAPTURES["gps_sig"]["bin"].read_bytes()
    assert out[: len(fx)].tobytes() == fx


# ---- f3 / f4: the restatements of CHANNEL::Start() and the receiver's SearchTask() against the reference's own code -------
def test_channel_start_restatement_vs_the_reference_log(oracle_mod):
    """Every ChanStart() of the golden run (UNMODIFIED c/channel.cpp behind oracle/ref_target_harness.cpp): the NCO rate
    words, the code-generator pause and the tap word the reference sent to the FPGA equal oracle.channel_start()."""
    from conftest import target_golden
    g, _ = target_golden()
    starts = [e for e in g["events"] if e["type"] == "start"]
    assert len(starts) >= 25 and any(e["lo_shift"] < 0 for e in starts)
    for e in starts:
        r = oracle_mod.channel_start(e["sv"], e["lo_shift"], e["ca_shift"], g["input"]["fc"], g["input"]["fs"], g["fft_len"], e["secs"])
        assert r["taps"] == e["taps"] == e.get("taps_sent", e["taps"])
        if "lo_rate" in e:
            assert (r["lo_rate"], r["ca_rate"]) == (e["lo_rate"], e["ca_rate"]), e
        if "mask" in e:                                  # Start() ran to its end: a missing CmdPause means ca_pause == 0
            assert r["ca_pause"] == e["ca_pause"], (e, r)


def test_search_task_restatement_vs_the_reference_log(oracle_mod):
    """The receiver's SearchTask() restated (oracle.search_task_on_target, one chunk at a time on the C oracle) replays
    the golden run: same detections on the same chunks, handed to the same channels, with the same bins -- across 25
    signal losses and re-acquisitions."""
    from conftest import target_golden, replay_target
    g, bits = target_golden()
    o = oracle_mod.Oracle(g["input"]["fc"], g["input"]["fs"])
    st = dict(busy=[False] * 32, chan_busy=0, sv=0)

    def feed(chunk):
        ev, pos, st["busy"], st["chan_busy"], st["sv"] = oracle_mod.search_task_on_target(o, chunk, 12, st["busy"], st["chan_busy"], st["sv"])
        assert pos == 1
        return [(e["sv"], e["ch"], e["lo_shift"], e["ca_shift"], e) for e in ev]

    def lost(ch, sv):
        st["chan_busy"] &= ~(1 << ch)
        st["busy"][sv] = False

    got = replay_target(g, bits, feed, lost)
    want = [(e["chunk"], e["sv"], e["ch"], e["lo_shift"], e["ca_shift"]) for e in g["events"] if e["type"] == "start"]
    assert [x[:5] for x in got] == want

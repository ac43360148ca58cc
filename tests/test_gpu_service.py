"""the consumers of the acquisition records.

  * hand-off arithmetic of CHANNEL::Start() (c/channel.cpp:134-171) -- CPU, bit-exact against the words the reference's own
    CHANNEL::Start() sent (golden run of the unmodified c/channel.cpp) and against the oracle restatement
  * the receiver's SearchTask() loop (c/search.cpp:214-239) -- GPU, batched and speculative: replays the golden run of the
    unmodified c/search.cpp + c/channel.cpp event for event, and equals the oracle's one-chunk-at-a-time loop on the capture
"""
import numpy as np
import pytest

from conftest import CAPTURES


def test_handoff_matches_channel_start(ga, oracle_mod):
    rng = np.random.default_rng(3)
    cases = [(0, 0, 0, 2.6e6, 10e6, 0.0), (5, -13, 9999, 2.6e6, 10e6, 0.25), (31, 36, 5455, 4.092e6, 5.456e6, 0.0075)]
    for _ in range(300):
        fs = float(rng.choice([10e6, 5.456e6, 8.184e6, 2.8e6]))
        fc = float(rng.choice([2.6e6, 4.092e6, 2.046e6, 0.62e6]))
        cases.append((int(rng.integers(0, 32)), int(rng.integers(-40, 41)), int(rng.integers(0, int(np.ceil(fs / 1000)))),
                      fc, fs, float(rng.uniform(0, 2.0))))
    for sv, lo, ca, fc, fs, secs in cases:
        peak = np.zeros(1, ga.PEAK_DTYPE)[0]
        peak["sv"], peak["lo_shift"], peak["ca_shift"], peak["snr"] = sv, lo, ca, 30.0
        got = ga.handoff(peak, fc, fs, fs, 40000, secs)
        want = oracle_mod.channel_start(sv, lo, ca, fc, fs, 40000, secs)
        for k in ("lo_rate", "ca_rate", "ca_shift", "ca_pause", "taps", "sv"):
            assert int(got[k]) == want[k], (k, sv, lo, ca, fc, fs, secs)
        assert got["lo_dop_hz"] == want["lo_dop_hz"] and got["ca_dop_hz"] == want["ca_dop_hz"]
    # the receiver's own numbers (FS = 10 MHz, FC = 2.6 MHz, 250 Hz bins): +4 bins = +1 kHz
    h = oracle_mod.channel_start(0, 4, 1234, 2.6e6, 10e6, 40000, 0.0)
    assert h["lo_dop_hz"] == 1000.0 and h["ca_pause"] == (20000 - 1234) % 10000 and h["taps"] == (2 << 4) + 6


def test_handoff_vs_the_reference_log(ga):
    """gpsacq_handoff_compute() against what the reference's own CHANNEL::Start() (UNMODIFIED c/channel.cpp, run behind
    oracle/ref_target_harness.cpp) sent to the FPGA for every ChanStart() of the golden run: carrier / code NCO words,
    code-generator pause (after the code creep over the logged 1.118 s), tap word.  Host code: needs no GPU."""
    from conftest import GOLD
    import json
    g = json.loads((GOLD / "ref_target_events.json").read_text())
    starts = [e for e in g["events"] if e["type"] == "start" and "mask" in e]
    assert len(starts) >= 25
    for e in starts:
        peak = np.zeros(1, ga.PEAK_DTYPE)[0]
        peak["sv"], peak["lo_shift"], peak["ca_shift"], peak["snr"] = e["sv"], e["lo_shift"], e["ca_shift"], 30.0
        h = ga.handoff(peak, g["input"]["fc"], g["input"]["fs"], g["input"]["fs"], g["fft_len"], e["secs"])
        assert (int(h["lo_rate"]), int(h["ca_rate"]), int(h["ca_pause"]), int(h["taps"])) == (e["lo_rate"], e["ca_rate"], e["ca_pause"], e["taps"]), e


@pytest.mark.gpu
@pytest.mark.parametrize("max_rounds", [1, 0])
def test_service_loop_vs_the_reference_log(ga, max_rounds):
    """gpsacq_service_* replays the golden run of the receiver's own SearchTask() + ChanTask()s (UNMODIFIED c/search.cpp and
    c/channel.cpp): 29 detections handed to the same channels on the same chunks with the same bins, across 25 signal
    losses (CHANNEL::SignalLost() -> channel freed, SearchEnable(sv)) and the re-acquisitions that follow."""
    from conftest import target_golden, replay_target
    g, bits = target_golden()
    with ga.Acquisition(g["input"]["fc"], g["input"]["fs"]) as acq:
        svc = ga.SearchService(acq, num_chans=12, max_rounds_per_batch=max_rounds)

        def feed(chunk):
            used, ev = svc.feed(chunk)
            assert used == 1
            return [(int(e["sv"]), int(e["ch"]), int(e["peak"]["lo_shift"]), int(e["peak"]["ca_shift"]), e) for e in ev]

        def lost(ch, sv):
            svc.signal_lost(ch)
            svc.enable(sv)

        got = replay_target(g, bits, feed, lost)
        svc.close()
    want = [(e["chunk"], e["sv"], e["ch"], e["lo_shift"], e["ca_shift"]) for e in g["events"] if e["type"] == "start"]
    assert [x[:5] for x in got] == want
    for x in got:
        assert int(x[5]["chunk_index"]) == x[0] and int(x[5]["start"]["taps"]) == next(e["taps"] for e in g["events"] if e["type"] == "start" and e["chunk"] == x[0])


def _events_equal(ev, want):
    assert len(ev) == len(want)
    for e, w in zip(ev, want):
        assert int(e["chunk_index"]) == w["chunk_index"] and int(e["sv"]) == w["sv"] and int(e["ch"]) == w["ch"]
        assert int(e["peak"]["lo_shift"]) == w["lo_shift"] and int(e["peak"]["ca_shift"]) == w["ca_shift"]
        assert abs(float(e["peak"]["snr"]) / w["snr"] - 1) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("num_chans", [12, 3])
def test_service_loop_matches_sequential_oracle(ga, oracle_mod, num_chans):
    c = CAPTURES["nottingham"]
    data = c["bin"].read_bytes()                      # 4 runs = 128 chunks
    ora = oracle_mod.Oracle(c["fc"], c["fs"])
    want, used_w, busy_w, chans_w, next_w = oracle_mod.search_task_on_target(ora, data, num_chans=num_chans)
    assert len(want) >= min(num_chans, 5)
    results = []
    for max_rounds in (1, 2, 0):                      # speculation depth must not change the outcome
        with ga.Acquisition(c["fc"], c["fs"]) as acq:
            svc = ga.SearchService(acq, num_chans=num_chans, max_rounds_per_batch=max_rounds)
            used, ev = svc.feed(data)
            busy, chans, seen = svc.state()
            svc.close()
        _events_equal(ev, want)
        assert used == used_w == seen
        assert busy == sum(1 << i for i, b in enumerate(busy_w) if b) and chans == chans_w
        results.append(ev.tobytes())
        # hand-off of every event = CHANNEL::Start() on its record, one chunk after the sample
        for e in ev:
            h = oracle_mod.channel_start(int(e["sv"]), int(e["peak"]["lo_shift"]), int(e["peak"]["ca_shift"]), c["fc"], c["fs"],
                                         40000, 40960 / c["fs"])
            for k in ("lo_rate", "ca_rate", "ca_shift", "ca_pause", "taps"):
                assert int(e["start"][k]) == h[k]
    assert results[0] == results[1] == results[2]


@pytest.mark.gpu
def test_service_reacquires_after_signal_lost(ga, oracle_mod):
    """CHANNEL::SignalLost() (c/channel.cpp:245-254): the channel is freed and the SV searched again."""
    c = CAPTURES["nottingham"]
    data = c["bin"].read_bytes()
    cb = 5120
    half = 64 * cb
    ora = oracle_mod.Oracle(c["fc"], c["fs"])
    w1, used1, busy, chans, nxt = oracle_mod.search_task_on_target(ora, data[:half])
    assert used1 == 64 and len(w1) >= 3
    lost = w1[0]
    busy[lost["sv"]] = False
    chans &= ~(1 << lost["ch"])
    w2, used2, _, _, _ = oracle_mod.search_task_on_target(ora, data[half:], busy=busy, chan_busy=chans, next_sv=nxt)
    with ga.Acquisition(c["fc"], c["fs"]) as acq:
        svc = ga.SearchService(acq)
        u1, e1 = svc.feed(data[:half])
        _events_equal(e1, w1)
        svc.signal_lost(int(e1[0]["ch"]))
        svc.enable(int(e1[0]["sv"]))
        u2, e2 = svc.feed(data[half:])
        svc.close()
    for w in w2:
        w["chunk_index"] += used1
    _events_equal(e2, w2)
    assert (u1, u2) == (used1, used2)
    assert any(int(e["sv"]) == lost["sv"] for e in e2)          # it is found again

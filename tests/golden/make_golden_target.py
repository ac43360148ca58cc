#!/usr/bin/env python
"""Golden event log of the receiver's own search loop + channel hand-off (development container only).

The UNMODIFIED c/search.cpp and c/channel.cpp of the reference run over a synthetic 10 MHz / IF 2.6 MHz chunk stream
(the constants of c/gps.h) behind oracle/ref_target_harness.cpp; every ChanStart() call and everything CHANNEL::Start()
and CHANNEL::SignalLost() send to the FPGA is logged.  Writes tests/golden/ref_target_events.json:
  input    how to regenerate the chunk stream (numpy generator, seeds) + its SHA-256
  events   in order: {"type": "start", chunk, ch, sv, taps, lo_shift, ca_shift, secs, lo_rate, ca_rate, ca_pause, mask}
                     {"type": "lost", chunk, ch, sv, mask}
"""
import hashlib
import importlib
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
import gpsacq_loader  # noqa: E402

N_CHUNKS, US_PER_YIELD, CN0, SEED_SATS, SEED_NOISE = 120, 2000, 50.0, 21, 6


def make_input():
    gpsacq_loader.load()
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    sats = sg.default_constellation(10e6, cn0_dbhz=CN0, seed=SEED_SATS)
    return sg.synth_capture(40960 * N_CHUNKS, 10e6, 2.6e6, sats, seed=SEED_NOISE), sats


def main():
    oracle.build()
    bits, sats = make_input()
    t = oracle.RefTarget()
    assert (t.fc, t.fs, t.num_chans) == (2.6e6, 10e6, 12)
    ev = t.run(bits, US_PER_YIELD)
    out = {"input": {"generator": "gnss_gps_sdr_b200.siggen.synth_capture (numpy)", "fs": 10e6, "fc": 2.6e6, "n_chunks": N_CHUNKS,
                     "cn0_dbhz": CN0, "seed_constellation": SEED_SATS, "seed_noise": SEED_NOISE,
                     "sha256": hashlib.sha256(bits.tobytes()).hexdigest()},
           "us_per_yield": US_PER_YIELD, "fft_len": 40000, "events": ev}
    json.dump(out, open(HERE / "ref_target_events.json", "w"), indent=0)
    print(len([e for e in ev if e["type"] == "start"]), "starts,", len([e for e in ev if e["type"] == "lost"]), "losses")
    for e in ev[:12]:
        print(e)


if __name__ == "__main__":
    main()

"""Throughput of the re-acquisition service loop (gpsacq_service_feed) on one GPU:
a synthetic stream with 8 satellites (C/N0 50 dB-Hz so that they clear the snr >= 25 rule) is fed through the
speculative batched loop; cold start (detections force re-batching) and steady state (the found SVs are tracked,
the other 24 are searched round robin) are timed end to end from host memory (wall clock around the C call)."""
import importlib, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
sg = importlib.import_module("gnss_gps_sdr_b200.siggen")

FS, FC = 5.456e6, 4.092e6
n_chunks = 2048
sats = sg.default_constellation(FS, cn0_dbhz=50.0, seed=1575420000)
bits = ga.synth_capture_gpu(n_chunks * 40960, FS, FC, sats, seed=21)
with ga.Acquisition(FC, FS, max_blocks=512) as acq:
    acq.search_blocks(bits[: 64 * 5120])                       # warm-up (context, clocks)
    svc = ga.SearchService(acq)
    t0 = time.perf_counter()
    used_cold, ev = svc.feed(bits[: 256 * 5120])
    t_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    used, ev2 = svc.feed(bits[used_cold * 5120:])
    t_steady = time.perf_counter() - t0
    busy, chans, seen = svc.state()
    svc.close()
    D = acq.n_doppler
print(json.dumps({"workload": "service loop, fs=5.456MHz REF grid (73 bins), 8 SVs at 50 dB-Hz, 12 channels",
                  "cold_start": {"chunks": used_cold, "events": len(ev), "ms": t_cold * 1e3, "correlations_per_s": used_cold * D / t_cold},
                  "steady_state": {"chunks": used, "events": len(ev2), "ms": t_steady * 1e3, "correlations_per_s": used * D / t_steady},
                  "detected_svs": sorted(int(e["sv"]) + 1 for e in list(ev) + list(ev2)), "generated_prns": sorted(s["prn"] for s in sats),
                  "busy_sv_mask": busy, "busy_channels": bin(chans).count("1")}))

"""Throughput of the stream converters either side of the search path, device buffers, CUDA events:
  iq8 -> bits   gpsacq_iq8_to_bits_device   (proc_rtl_bin_for_gps.m / proc_hackrf_bin_for_gps.m): 2 B in + 1/8 B out per sample,
                                            the input is read twice (mean, then conversion)
  bits -> iq8   gpsacq_bits_to_iq8_device   (c/conv_1bit_bin_to_hackrf_bin.cpp:29-86): 1/8 B in, 2 B out per sample
Prints one JSON line per converter with GB/s against the measured HBM copy peak."""
import json, sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
dev = torch.device("cuda", 0)
n = 1 << 29                                    # 512 Mi complex samples: 1 GiB of IQ bytes (>> L2)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
acq = ga.Acquisition(0.62e6, 2.8e6, 5000.0, device=0, max_blocks=32)
acq.set_stream(stream.cuda_stream)
iq = torch.randint(0, 256, (2 * n,), dtype=torch.uint8, device=dev)
bits = torch.zeros(n // 8, dtype=torch.uint8, device=dev)
sums = torch.zeros(16, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


import os
for version in (1, 2):          # 1: double-precision table kernel / compare-select expander; 2: threshold table / byte permutes
    os.environ["GPSACQ_FRONTEND_V2"] = "1" if version == 2 else "0"
    ms = timed(lambda: acq.iq8_to_bits_device(iq.data_ptr(), n, 0.62e6, 2.8e6, bits.data_ptr(), sums.data_ptr()))
    moved = 2 * (2 * n) + n // 8                   # two reads of the IQ bytes + the packed bits
    print(json.dumps({"converter": "iq8_to_bits", "version": version, "samples": n, "ms": ms, "gsamples_per_s": n / ms / 1e6,
                      "gbs_moved": moved / ms / 1e6, "frac_of_hbm_peak": moved / ms / 1e6 / peak,
                      "note": "2 B read twice (exact integer mean, then shift + sign) + 1/8 B written per sample"}), flush=True)
    out = torch.zeros(2 * n, dtype=torch.int8, device=dev)
    ms = timed(lambda: ga.bits_to_iq8_device(bits.data_ptr(), n // 8, 2.6e6, 10e6, out.data_ptr(), device=0, stream_ptr=stream.cuda_stream))
    moved = n // 8 + 2 * n
    print(json.dumps({"converter": "bits_to_iq8", "version": version, "samples": n, "ms": ms, "gsamples_per_s": n / ms / 1e6,
                      "gbs_moved": moved / ms / 1e6, "frac_of_hbm_peak": moved / ms / 1e6 / peak,
                      "note": "1/8 B read + 2 B written per sample"}), flush=True)
    del out
del os.environ["GPSACQ_FRONTEND_V2"]
acq.close()

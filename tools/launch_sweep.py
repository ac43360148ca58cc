"""Scratch: whole-batch device-path throughput of the REF engine against the launch size (GPSACQ_SUB_BLOCKS) and the
cell scheduler (GPSACQ_STATIC_SCHED), for the library named by GPSACQ_LIB.
    python tools/launch_sweep.py [n_chunks] [sub,sub,...] [static]"""
import os
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 3584
subs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128, 512, 3584]
if len(sys.argv) > 3 and sys.argv[3] == "static":
    os.environ["GPSACQ_STATIC_SCHED"] = "1"
fc, fs = float(os.environ.get("SWEEP_FC", 4.092e6)), float(os.environ.get("SWEEP_FS", 5.456e6))
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
d_bits = torch.from_numpy(rng.integers(0, 256, nb * 5120, dtype=np.uint8)).to(dev)
d_out = torch.zeros(nb * 32, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
for sub in subs:
    os.environ["GPSACQ_SUB_BLOCKS"] = str(sub)
    acq = ga.Acquisition(fc, fs, 5000.0, device=0, max_blocks=nb)
    acq.set_stream(stream.cuda_stream)
    for _ in range(2):
        acq.search_blocks_device(d_bits.data_ptr(), nb, None, d_out.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        acq.search_blocks_device(d_bits.data_ptr(), nb, None, d_out.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    pk = np.frombuffer(d_out.cpu().numpy().tobytes(), ga.PEAK_DTYPE)
    print("%s %s fs=%.4g T=%d sub=%d: %.3f ms per %d chunks -> %.3f Mcorr/s (contract frac %.3f)  last launch cells %.3f ms fwd %.3f ms  checksum %.6e" % (
        os.environ.get("GPSACQ_LIB", "default"), "static" if os.environ.get("GPSACQ_STATIC_SCHED") else "ticket", fs, acq.info["cell_threads"], sub, ms, nb,
        nb * acq.n_doppler / ms / 1e3, nb * acq.n_doppler / ms * 1e3 * 640016 / 6550.1e9, acq.stage_times()["cells_ms"], acq.stage_times()["fwd_ms"],
        float(pk["snr"].astype(np.float64).sum())), flush=True)
    acq.close()

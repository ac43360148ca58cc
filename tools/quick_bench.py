"""Scratch: time the cell kernel of a given libgpsacq build (GPSACQ_LIB) on a 512-chunk batch."""
import sys, os, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
fc, fs = (4.092e6, 5.456e6) if len(sys.argv) < 3 else (float(sys.argv[1]), float(sys.argv[2]))
rng = np.random.default_rng(0)
bits = rng.integers(0, 256, 512 * 5120, dtype=np.uint8)
acq = ga.Acquisition(fc, fs)
ref = None
for i in range(3):
    pk = acq.search_blocks(bits)
ts = []
for i in range(10):
    acq.search_blocks(bits); ts.append(acq.stage_times())
c = np.median([t["cells_ms"] for t in ts]); f = np.median([t["fwd_ms"] for t in ts]); tot = np.median([t["total_ms"] for t in ts])
n = 512 * acq.n_doppler // 4          # the host-buffer call is cut into 4 slices; stage_times() are those of the last slice (128 chunks)
print("%s fs=%.3g: cells %.3f ms -> %.3f Mcorr/s (%.1f%% of 6489.9 GB/s) fwd %.3f total %.3f  checksum %.6e" % (
    os.environ.get("GPSACQ_LIB", "default"), fs, c, n / c / 1e3, n / c * 1e3 * 640016 / 6489.9e9 * 100, f, tot, float(pk["snr"].astype(np.float64).sum())))

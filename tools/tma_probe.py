"""Scratch: one small REF search through the default cell kernel (TMA-staged) against the __ldg kernel."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bits = np.random.default_rng(1).integers(0, 256, nb * 5120, dtype=np.uint8)
os.environ["GPSACQ_CELL_TMA"] = "1"
acq = ga.Acquisition(4.092e6, 5.456e6, 5000.0, max_blocks=nb)
a = acq.search_blocks(bits).copy()
st = [acq.cell_stats(b).copy() for b in (0, nb - 1)]
acq.close()
os.environ["GPSACQ_CELL_TMA"] = "0"
acq = ga.Acquisition(4.092e6, 5.456e6, 5000.0, max_blocks=nb)
b = acq.search_blocks(bits).copy()
st2 = [acq.cell_stats(bk).copy() for bk in (0, nb - 1)]
acq.close()
print("records identical:", a.tobytes() == b.tobytes(), " cell stats identical:", all(x.tobytes() == y.tobytes() for x, y in zip(st, st2)))
if a.tobytes() != b.tobytes():
    bad = np.nonzero(a["snr"] != b["snr"])[0]
    print("differing records:", bad[:10], a[bad[:3]], b[bad[:3]])

"""Small REF + GRID runs for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
data = (ROOT / "tests/golden/nottingham_fs5456_if4092_runs0-3.bin").read_bytes()
with ga.Acquisition(4.092e6, 5.456e6, max_blocks=8) as a:
    pk = a.search_blocks(data[: 11 * 5120])
    print("REF", pk["lo_shift"][:4], pk["ca_shift"][:4])
with ga.Acquisition(2.046e6, 8.184e6, max_blocks=4) as a:
    pk = a.search_blocks((ROOT / "tests/golden/gps_sig_fs8184_if2046_runs0-1.bin").read_bytes()[: 5 * 5120])
    print("REF 8.184", pk["lo_shift"][:4])
with ga.Acquisition(4.092e6, 5.456e6, 2000.0, mode=1, doppler_step=500.0, noncoh_blocks=2, max_blocks=1) as a:
    pk = a.acquire(data[: 2 * 682])
    print("GRID", pk["lo_shift"][:4], pk["ca_shift"][:4])

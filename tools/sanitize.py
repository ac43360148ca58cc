"""Small REF + GRID runs for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once --
REF cells (TMEM), GRID native prime-factor cells (K = 1 and K > 1 / TMEM), GRID embedding, front-end, generator."""
import os
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
os.environ["GPSACQ_CELL_TMA"] = "1"            # the TMA-staged REF cell kernel (ga_cell_tma.cuh): mbarrier ring, UTMALDG
with ga.Acquisition(4.092e6, 5.456e6, max_blocks=8) as a:
    pk2 = a.search_blocks(data[: 11 * 5120])
    print("REF tma", a.info["cell_threads"], bool((pk2["ca_shift"] == pk["ca_shift"]).all()))
os.environ["GPSACQ_CELL_TMA"] = "0"
with ga.Acquisition(16.368e6 / 4, 16.368e6, max_blocks=2) as a:      # FS > 10 MHz: segmented windows
    pk3 = a.search_blocks(data[: 2 * 5120])
    print("REF 16.368", pk3["ca_shift"][:2])
with ga.Acquisition(2.046e6, 8.184e6, max_blocks=4) as a:
    pk = a.search_blocks((ROOT / "tests/golden/gps_sig_fs8184_if2046_runs0-1.bin").read_bytes()[: 5 * 5120])
    print("REF 8.184", pk["lo_shift"][:4])
with ga.Acquisition(4.092e6, 5.456e6, 2000.0, mode=1, doppler_step=500.0, noncoh_blocks=2, max_blocks=1) as a:
    pk = a.acquire(data[: 2 * 682])
    print("GRID", pk["lo_shift"][:4], pk["ca_shift"][:4])
with ga.Acquisition(4.092e6, 5.456e6, 1500.0, mode=1, doppler_step=500.0, noncoh_blocks=1, max_blocks=1) as a:
    pk = a.acquire(data[: 682])
    print("GRID native K=1", a.info["fft_len"], pk["lo_shift"][:4], pk["ca_shift"][:4])
with ga.Acquisition(0.62e6, 2.8e6, 1000.0, mode=1, doppler_step=250.0, noncoh_blocks=3, max_blocks=1) as a:
    pk = a.acquire(data[: 3 * 350])
    print("GRID native 2800 K=3", a.info["fft_len"], pk["ca_shift"][:4])
with ga.Acquisition(2.046e6, 8.184e6, 1000.0, mode=1, doppler_step=500.0, noncoh_blocks=1, max_blocks=1, dop_first=1, dop_count=3) as a:
    pk = a.acquire(data[: 1023])
    print("GRID native 8184 shard", a.info["fft_len"], pk["lo_shift"][:4])
os.environ["GPSACQ_GRID_EMBED"] = "1"
with ga.Acquisition(4.092e6, 5.456e6, 1000.0, mode=1, doppler_step=500.0, noncoh_blocks=2, max_blocks=1) as a:
    pk = a.acquire(data[: 2 * 682])
    print("GRID embedding", a.info["fft_len"], pk["ca_shift"][:4])
del os.environ["GPSACQ_GRID_EMBED"]
with ga.Acquisition(4.092e6, 5.456e6, max_blocks=32) as a:
    svc = ga.SearchService(a, max_rounds_per_batch=1)
    used, ev = svc.feed(data[: 40 * 5120])
    print("service", used, [int(e["sv"]) for e in ev])
    svc.close()
    iq = np.random.default_rng(1).integers(0, 256, 2 * 50000, dtype=np.uint8)
    bits = a.iq8_to_bits(iq, 0.62e6, 2.8e6)
    print("front-end", None if bits is None else int(bits[:4].sum()))
    # both generations of both converters (threshold-table / permute kernels, and the ones they replaced), odd lengths
    for v in ("0", "1"):
        os.environ["GPSACQ_FRONTEND_V2"] = v
        b1 = a.iq8_to_bits(iq[: 2 * 49997], 2.6e6, 10e6, signed=True)
        b2 = a.iq8_to_bits(iq[: 2 * 1234], 123456.789, 2.8e6)          # not a small fraction: the sincos kernel
        out = ga.bits_to_iq8(data[: 4099], 2.6e6, 10e6, 30, first_sample=8 * 777)
        print("converters v" + v, int(b1[:4].sum()), int(b2[:4].sum()), int(out[:8].astype(int).sum()))
    del os.environ["GPSACQ_FRONTEND_V2"]

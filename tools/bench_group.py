"""One process, several GPUs: gpsacq_group_search_blocks() (what the C++ host uses with GPSACQ_GPUS=n) on the bench
workload -- 7168 chunks per GPU per step from ONE pinned host buffer, chunk ranges per device, one ncclAllGather of the
peak records per step, records back on the host.  Wall clock around the C call (it returns with the records).
    python tools/bench_group.py [n_gpus] [steps]"""
import json, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
ng = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
per = 7168
host = torch.empty(ng * per * 5120, dtype=torch.uint8).pin_memory()
host.numpy()[:] = np.random.default_rng(5).integers(0, 256, host.numel(), dtype=np.uint8)
grp = ga.AcquisitionGroup(4.092e6, 5.456e6, 5000.0, n_gpus=ng, use_nccl=True, max_blocks=per)
for _ in range(2):
    pk = grp.search_blocks(host.numpy())
t0 = time.perf_counter()
for _ in range(steps):
    pk = grp.search_blocks(host.numpy())
dt = (time.perf_counter() - t0) / steps
print(json.dumps({"what": "gpsacq_group_search_blocks, one process", "n_gpus": ng, "gather": grp.gather_kind, "chunks_per_step": ng * per,
                  "ms_per_step": dt * 1e3, "correlations_per_s": ng * per * 73 / dt, "per_gpu": per * 73 / dt,
                  "checksum": float(pk["snr"].astype(np.float64).sum())}))
grp.close()

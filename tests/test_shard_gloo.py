"""CPU-only, world_size 2 over gloo: the N>1 host logic (run sharding + peak-record gather)."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import CAPTURES, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import gpsacq_loader, importlib
ga = gpsacq_loader.load(); shard = importlib.import_module("gnss_gps_sdr_b200.shard")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
ref = np.load(sys.argv[2]); n_runs = len(ref) // 32                      # golden per-chunk records of the fixture
full = np.zeros(len(ref), ga.PEAK_DTYPE)
for f in ("snr", "lo_shift", "ca_shift", "sv"): full[f] = ref[f]
lo, hi = shard.run_range(n_runs, rank, world)
mine = torch.from_numpy(np.frombuffer(full[lo * 32: hi * 32].tobytes(), np.uint8).copy())   # "this rank searched its runs"
allp = shard.gather_peaks(mine, n_runs, world)
got = shard.peaks_from_bytes(allp.numpy().tobytes(), ga.PEAK_DTYPE)
assert got.tobytes() == full.tobytes(), "gathered records are not in stream order"
text = "".join(ga.format_run(r, got[32 * r: 32 * r + 32]) for r in range(n_runs))
if rank == 0: open(sys.argv[3], "w").write(text)
dist.barrier(); dist.destroy_process_group()
"""


def test_run_range_is_a_partition(ga):
    import importlib
    shard = importlib.import_module("gnss_gps_sdr_b200.shard")
    for n in (0, 1, 3, 4, 16, 17, 340):
        for w in (1, 2, 3, 4, 8):
            parts = [shard.run_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_reproduces_golden_report(tmp_path):
    c = CAPTURES["nottingham"]          # 4 runs; world 2 -> 2+2; also try 3 "ranks worth" through uneven n below
    port = _free_port()
    out = tmp_path / "report.txt"
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER, str(ROOT), str(c["peaks"]), str(out)], env=env))
    assert all(p.wait(timeout=120) == 0 for p in procs)
    golden = "\n".join(c["stdout"].read_text().split("\n")[6:])
    assert golden.startswith(out.read_text())

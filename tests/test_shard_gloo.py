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


# ---- GRID mode: the Doppler grid of one acquisition split over ranks ------------------------------------------
GRID_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import gpsacq_loader, importlib
ga = gpsacq_loader.load(); shard = importlib.import_module("gnss_gps_sdr_b200.shard")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n_rec, n_bins, dmax = 64, 21, 10
rng = np.random.default_rng(5)                                   # same table on every rank
snr = rng.uniform(1, 30, (n_rec, n_bins)).astype(np.float32)
snr[3, 4] = snr[3, 17] = 99.0                                    # exact tie across shards: the lower bin must win
snr[7, :] = 0.0                                                  # nothing above 0: record stays empty
idx = rng.integers(0, 5456, (n_rec, n_bins))
def best_over(lo, n):                                            # what best_kernel does on bins [lo, lo+n)
    out = np.zeros(n_rec, ga.PEAK_DTYPE)
    for i in range(n_rec):
        m = 0.0
        for b in range(lo, lo + n):
            if snr[i, b] > m:
                m = snr[i, b]; out[i]["snr"] = m; out[i]["lo_shift"] = b - dmax; out[i]["ca_shift"] = idx[i, b]
        out[i]["sv"] = i % 32
    return out
lo, n = shard.bin_range(n_bins, rank, world)
mine = torch.from_numpy(np.frombuffer(best_over(lo, n).tobytes(), np.uint8).copy())
got = shard.gather_merge_peaks(mine, world, ga.PEAK_DTYPE)
want = best_over(0, n_bins)
assert got.tobytes() == want.tobytes(), "merged shards differ from the full-grid scan"
assert got[3]["lo_shift"] == 4 - dmax and got[7]["snr"] == 0
dist.barrier(); dist.destroy_process_group()
"""


def test_bin_range_is_a_partition(ga):
    import importlib
    shard = importlib.import_module("gnss_gps_sdr_b200.shard")
    for n in (1, 21, 801, 2001):
        for w in (1, 2, 3, 4, 8):
            if w > n:
                continue
            parts = [shard.bin_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][0] + parts[-1][1] == n
            assert all(a[0] + a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1


def test_two_rank_doppler_shards_merge_to_the_full_grid():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", GRID_WORKER, str(ROOT)], env=env))
    assert all(p.wait(timeout=120) == 0 for p in procs)

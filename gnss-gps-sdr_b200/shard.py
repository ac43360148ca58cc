"""Multi-GPU sharding of the acquisition stream (one process per GPU, torch.distributed).

Cells (chunk x Doppler bin) are independent; the only coupling is the best-over-Doppler scan
inside one chunk (c/search_offline.cpp:196-198), so the stream is split by RUN (32 chunks, one per
PRN): rank r of `world` owns a contiguous range of runs, searches it alone, and the 32-byte peak
records are all-gathered once per step.  No collective touches the data path.
"""
from __future__ import annotations

import numpy as np

PEAK_BYTES = 32


def run_range(n_runs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of runs [lo, hi) -- first (n_runs % world) ranks get one more."""
    base, extra = divmod(n_runs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_peaks(local_u8, n_runs: int, world: int, group=None):
    """All-gather per-rank peak records (uint8 tensor of len runs_r*32*32) into stream order.

    Ranks may own different numbers of runs; records are padded to the largest share for the
    collective and trimmed afterwards.  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist
    per = [run_range(n_runs, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in per) * 32 * PEAK_BYTES
    pad = torch.zeros(cap, dtype=torch.uint8, device=local_u8.device)
    pad[: local_u8.numel()] = local_u8
    out = torch.empty(world * cap, dtype=torch.uint8, device=local_u8.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * cap: r * cap + (hi - lo) * 32 * PEAK_BYTES] for r, (lo, hi) in enumerate(per)]
    return torch.cat(parts)


# ---- GRID mode: ONE acquisition's (PRN x Doppler) grid split by Doppler bin ----------------------------
def bin_range(n_bins: int, rank: int, world: int) -> tuple[int, int]:
    """(dop_first, dop_count) of rank's contiguous shard of the n_bins-bin Doppler grid -- the same split as
    gpsacq_group_create() (ascending ranges, so a lower rank always holds lower bins)."""
    lo = n_bins * rank // world
    return lo, n_bins * (rank + 1) // world - lo


def merge_peaks(per_rank: np.ndarray) -> np.ndarray:
    """per_rank[r, i] = record i of rank r (rank order = ascending bin ranges).  Keeps, per record, the
    shard with the highest snr; on equal snr the lowest rank, i.e. the lower Doppler bin -- exactly the
    winner of the reference's ascending strictly-greater scan (c/search_offline.cpp:173,198)."""
    per_rank = np.asarray(per_rank)
    best = per_rank[0].copy()
    for r in range(1, per_rank.shape[0]):
        take = per_rank[r]["snr"] > best["snr"]
        best[take] = per_rank[r][take]
    return best


def gather_merge_peaks(local_u8, world: int, dtype, group=None) -> np.ndarray:
    """All-gather every rank's records for the same acquisitions (uint8 tensor, CUDA -> NCCL, CPU -> gloo)
    and merge them with merge_peaks()."""
    import torch
    import torch.distributed as dist
    out = torch.empty(world * local_u8.numel(), dtype=torch.uint8, device=local_u8.device)
    dist.all_gather_into_tensor(out, local_u8.contiguous(), group=group)
    rec = np.frombuffer(out.cpu().numpy().tobytes(), dtype=dtype).reshape(world, -1)
    return merge_peaks(rec)


def peaks_from_bytes(buf, dtype) -> np.ndarray:
    return np.frombuffer(bytes(buf), dtype=dtype)

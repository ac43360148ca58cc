#!/usr/bin/env python
"""bench.py -- PRN x Doppler correlations/sec of the B200 GPS L1 C/A acquisition engine.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's engine
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's own CPU path

Workload (config.workload): BASELINE.json configs[1] input parameters -- 32 PRN, +-5 kHz,
fs = 5.456 MHz, IF = 4.092 MHz, synthetic 1-bit IQ -- searched with the REFERENCE's grid
semantics (N = 40000-point coherent window, 73 Doppler bins of fs/N = 136.4 Hz, one 5120-byte
chunk per PRN; c/search_offline.cpp:176,239-246), because that is the only grid on which parity
with gps_test is defined and the only one the reference arm can run.

A "step" is one batch of 224 runs = 7168 chunks x 73 bins = 523,264 correlations per GPU (~57 ms, so
that 20 timed steps run for more than a second): forward FFT kernel, cell kernel (shifted conj-multiply
+ pruned backward FFT + |.|^2 + peak), best-over-Doppler kernel -- ONE launch of each per step (the cell
kernel's CTAs draw their cells from a device-wide ticket counter, so a block spectrum is fetched from HBM
once and shared through L2 by its 73 cells however long the launch is) -- and for N > 1 one NCCL
all-gather of the 7168 32-byte peak records per rank -- inside the timed region of BOTH value and e2e.
Ranks work on different runs of the stream (weak scaling, no data-path collective).

value   : device-resident inputs (packed bits already in HBM), CUDA events on the launching stream.
e2e     : same step through the host-buffer C-ABI call gpsacq_search_blocks(): pinned host bits ->
          H2D -> kernels -> D2H peak records, host synchronised, every step; for N > 1 followed by the
          all-gather of every rank's records and their read-back to the host.
roofline: the cell kernel alone.  achieved = 640,016 algorithmic bytes per correlation (one read
          of the 40000-point complex64 block spectrum + one of the replica spectrum + a 16-byte
          record) x correlations per launch / average launch time from CUDA
          events recorded around the launch inside libgpsacq (gpsacq_stage_times()).
          peak = MEASURED_PEAKS.json hbm_gbs.  The operands are L2-resident after first touch, so
          the real DRAM traffic is far below the algorithmic bytes (roofline.traffic, from ncu) and the
          kernel is bound by FP32 issue: roofline.fp32_frac = executed FP32 flops (ncu instruction counts
          per correlation, profiles/cell_kernel_ncu.json) / live launch time / 74.4 TFLOP/s, next to the
          ncu issue-slot and FMA-pipe utilisation of the same capture.
Secondary blocks in the same line (N = 1): GRID-mode throughput for BASELINE.json configs[1], [2], [3]
(grid_mode_configs1/2/3) and the Doppler-sharded configs[4] acquisition (grid_mode_configs4_sharded).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FC, FS, MAX_FO = 4.092e6, 5.456e6, 5000.0
RUNS_PER_STEP = 224                  # 7168 chunks x 73 bins = 523,264 correlations (~60 ms) per GPU per step
CHUNK = 5120
N_BATCHES = 4                       # distinct input batches cycled through the steps
FP32_PEAK_TFLOPS = 74.4             # 148 SMs x 128 FP32 lanes x 2 flops x 1.965 GHz (non-tensor FP32 peak of a B200)
METRIC = "PRN×Doppler correlations/sec"
UNIT = "correlations/s"


def bench_sats():
    import gpsacq_loader
    import importlib
    gpsacq_loader.load()
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    return sg, sg.default_constellation(FS, seed=1575420000)


def make_batches(n_batches: int, seed: int, device: int):
    """Synthetic captures of the bench workload (8 satellites at 45 dB-Hz + noise), made on the GPU by the library's
    own generator (gpsacq_synth_capture; the numpy generator would need minutes for 4 x 294 M samples)."""
    import gpsacq_loader
    ga = gpsacq_loader.load()
    _, sats = bench_sats()
    n_samples = RUNS_PER_STEP * 32 * CHUNK * 8
    return [ga.synth_capture_gpu(n_samples, FS, FC, sats, seed=seed * 1000 + b, device=device) for b in range(n_batches)]


def make_cpu_sample(runs: int, seed: int):
    """The same workload for the CPU legs (numpy generator: no GPU needed by the reference arm)."""
    sg, sats = bench_sats()
    return sg.synth_capture(runs * 32 * CHUNK * 8, FS, FC, sats, seed=seed)


def workload_config(n_gpus: int):
    return {"workload": "C1-REF: synthetic 1-bit IF fs=5.456MHz if=4.092MHz, 32 PRN x 73 Doppler bins "
                        f"(+-5 kHz @ 136.4 Hz), N=40000 coherent (7.33 ms), {RUNS_PER_STEP} runs ({RUNS_PER_STEP * 32} chunks) per GPU per step",
            "correlations_per_step_per_gpu": RUNS_PER_STEP * 32 * 73,
            "sharding": f"runs of the stream split over {n_gpus} GPU(s); NCCL all-gather of peak records every step, overlapped with the next step's kernels (double-buffered records)"
                        if n_gpus > 1 else "single GPU",
            "l2_policy": f"inputs larger than L2: each step's cell-kernel operands are {RUNS_PER_STEP * 32 * 320 // 1000} MB of block spectra "
                         "+ 20 MB of replica spectra (> 126 MB L2); 4 distinct input batches are cycled"}


# -------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed region (NVML, ~2 ms period;
    falls back to polling nvidia-smi)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nv is not None:
                    nv = self.nv
                    self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                      nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                      nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
                    self._halt.wait(0.002)
                else:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                                          "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.max_mhz = int(out[1])
                    self.rows.append((int(out[0]), float(out[2]), int(out[3], 16)))
                    self._halt.wait(0.05)
            except Exception:
                self._halt.wait(0.05)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        reasons = sorted(n for n, bit in self.BAD.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_mhz_min": min(r[0] for r in self.rows), "sm_max_mhz": self.max_mhz,
                "power_w_max": max(r[1] for r in self.rows), "samples": len(self.rows), "reasons": reasons}


# -------------------------------------------------------------------------------------------------
def cpu_worker(args):
    """One process of the CPU baseline: the reference's Sample()+Correlate() (oracle/_ref) or the
    C port (oracle/liboracle.so) over `runs` runs of the bench workload."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    bits = np.fromfile(args.cpu_worker, np.uint8)
    nb = args.cpu_runs * 32
    off = (args.cpu_index * nb * CHUNK) % max(1, bits.size - nb * CHUNK + 1)
    data = bits[off: off + nb * CHUNK].tobytes()
    if args.cpu_kind == "reference":
        eng = oracle.RefHarness(FC, FS, MAX_FO)
        backend = eng.fft_backend
    else:
        eng = oracle.Oracle(FC, FS, MAX_FO, fft_f64=False)
        backend = "builtin-f32"
    eng.search_blocks(data[: 2 * CHUNK])              # warm-up (plans, page-in)
    t0 = time.perf_counter()
    eng.search_blocks(data)
    dt = time.perf_counter() - t0
    print(json.dumps({"corr": nb * 73, "seconds": dt, "backend": backend}))


def usable_cores() -> int:
    """Host cores this process may actually use: affinity mask, capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


_CPU_SAMPLE = None


def cpu_sample_file() -> str:
    """The bench workload's synthetic capture (one batch), written once for the CPU workers."""
    global _CPU_SAMPLE
    if _CPU_SAMPLE is None:
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            make_cpu_sample(16, seed=77).tofile(f)
            _CPU_SAMPLE = f.name
    return _CPU_SAMPLE


def cpu_baseline(runs_per_core: int, cores: int | None = None):
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    kind = "reference" if oracle.ref_available() else "port"
    cores = cores or usable_cores()
    path = cpu_sample_file()
    env = oracle.mkl_env() if kind == "reference" else dict(os.environ)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, __file__, "--cpu-worker", path, "--cpu-kind", kind, "--cpu-runs",
                               str(runs_per_core), "--cpu-index", str(i)], stdout=subprocess.PIPE, text=True, env=env)
             for i in range(cores)]
    res = [json.loads(p.communicate()[0].strip().split("\n")[-1]) for p in procs]
    wall = time.perf_counter() - t0
    busy = max(r["seconds"] for r in res)
    total = sum(r["corr"] for r in res)
    return {"value": total / busy, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{runs_per_core} run(s) = {runs_per_core * 32 * 73} correlations per core of the bench workload, "
                      f"{cores} single-threaded processes of the reference's Sample()+Correlate() "
                      f"(FFT backend: {res[0]['backend']}); {busy:.1f} s busy, {wall:.1f} s wall incl. start-up",
            "single_core": res[0]["corr"] / res[0]["seconds"]}, total, busy


# -------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    # size the sample so K timed steps finish in a few minutes: one "step" = runs_per_core runs on every core
    runs_per_core = 2
    for _ in range(args.warmup):
        pass                                   # worker processes warm themselves up (plans, page-in)
    vals, ms = [], []
    for _ in range(max(1, args.steps)):
        cb, total, busy = cpu_baseline(runs_per_core)
        vals.append(total / busy)
        ms.append(busy * 1e3)
    v = statistics.median(vals)
    os.unlink(cpu_sample_file())
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": statistics.median(ms), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_engine(args, rank, world, local_rank):
    import torch
    import gpsacq_loader
    ga = gpsacq_loader.load()
    dist = None
    json_fd = None
    if world > 1:
        # NCCL writes its version banner (NCCL_DEBUG=VERSION/WARN) to file descriptor 1.  The contract wants ONE JSON
        # line on stdout: send everything else that lands on fd 1 to stderr and keep a private handle for the line.
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    nb = RUNS_PER_STEP * 32
    batches = make_batches(N_BATCHES, seed=rank + 1, device=local_rank)   # every rank searches different runs of the stream
    acq = ga.Acquisition(FC, FS, MAX_FO, device=local_rank, max_blocks=nb)
    ndop = acq.n_doppler
    corr_per_step = nb * ndop
    d_bits = [torch.from_numpy(b).to(dev) for b in batches]
    h_bits = [torch.from_numpy(b).pin_memory() for b in batches]
    # records double-buffered: the NCCL all-gather of step i runs (on NCCL's stream) while step i+1 computes
    d_outs = [torch.zeros(nb * 32, dtype=torch.uint8, device=dev) for _ in range(2)]
    d_alls = [torch.zeros(world * nb * 32, dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None
    works = [None, None]
    # a non-default stream: libgpsacq launches on it and the CUDA events below are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    acq.set_stream(stream.cuda_stream)

    def step_device(i):
        k = i & 1
        if works[k] is not None:
            works[k].wait()                    # the gather that last used this buffer pair (two steps ago) is done
        acq.search_blocks_device(d_bits[i % N_BATCHES].data_ptr(), nb, None, d_outs[k].data_ptr())
        if world > 1:
            works[k] = dist.all_gather_into_tensor(d_alls[k], d_outs[k], async_op=True)

    def drain():
        for k in range(2):
            if works[k] is not None:
                works[k].wait()                # makes the current stream wait for the collective
                works[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing ------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cell_ms = []
    e0.record(stream)
    for i in range(args.steps):
        step_device(i)
    drain()
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    # per-launch time of the dominant kernel: a second pass with a read-back of the library's events
    for i in range(min(args.steps, 8)):
        step_device(i)
        cell_ms.append(acq.stage_times()["cells_ms"])
    drain()
    barrier()
    clocks = sampler.finish()
    stage = acq.stage_times()

    # ---- end to end through the host-buffer C-ABI call --------------------------------------------
    # every step: pinned host bits -> gpsacq_search_blocks() (H2D, kernels, D2H, host sync) -> records on the host;
    # N > 1: followed by the peak gather (records back to the device buffer NCCL reads, all-gather, read-back of all
    # ranks' records), so that the exchange is inside this timed region too
    acq.set_stream(None)
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    h_all = torch.zeros(world * nb * 32, dtype=torch.uint8).pin_memory() if world > 1 else None

    def step_e2e(i):
        pk = acq.search_blocks(h_bits[i % N_BATCHES].numpy())
        if world > 1:
            d_outs[0].copy_(torch.from_numpy(pk.view(np.uint8).reshape(-1)), non_blocking=True)
            dist.all_gather_into_tensor(d_alls[0], d_outs[0])
            h_all.copy_(d_alls[0], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        return pk

    peaks = None
    for i in range(min(args.warmup, 3)):
        peaks = step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        peaks = step_e2e(i)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    detected = int((peaks["snr"] >= 25).sum())
    grid_c4 = None if args.no_grid else grid_c4_sharded(ga, dev, rank, world, dist)      # collective: every rank takes part

    if rank == 0:
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            peak_gbs, peak_src = json.loads(peaks_file.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bpc = acq.info["bytes_per_corr"]
        cell_avg_ms = statistics.mean(cell_ms)
        # a step is blocks_per_launch chunks per launch triple (the whole step unless GPSACQ_SUB_BLOCKS cuts it); the stage
        # events bracket the last launch triple of a step
        launches_per_step = -(-nb // acq.info["blocks_per_launch"])
        corr_per_launch = acq.info["blocks_per_launch"] * ndop
        achieved = corr_per_launch * bpc / (cell_avg_ms * 1e-3) / 1e9
        # ncu-derived per-correlation figures of the same kernel (profiles/cell_kernel_ncu.json, made by tools/ncu_summary.py
        # + tools/ncu_flops.py from the committed capture): DRAM bytes, executed FP32 flops, issue-slot / FMA-pipe utilisation
        traffic = fp32 = None
        ncu = {}
        prof = ROOT / "profiles" / "cell_kernel_ncu.json"
        if prof.exists():
            ncu = json.loads(prof.read_text())
            traffic = ncu["dram_bytes_per_corr"] * corr_per_launch
            fp32 = ncu["fp32_flops_per_corr"] * corr_per_launch / (cell_avg_ms * 1e-3) / 1e12
        value = world * corr_per_step * args.steps / (total_ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                             "frac": achieved / peak_gbs, "traffic": traffic, "kernel": "cell_kernel",
                             "launch_ms": cell_avg_ms, "bytes_per_launch": corr_per_launch * bpc,
                             "correlations_per_launch": corr_per_launch, "launches_per_step": launches_per_step, "peak_source": peak_src,
                             "whole_step_frac": value / world * bpc / 1e9 / peak_gbs,
                             "fp32_tflops": fp32, "fp32_peak_tflops": FP32_PEAK_TFLOPS,
                             "fp32_frac": None if fp32 is None else fp32 / FP32_PEAK_TFLOPS,
                             "issue_active_pct_ncu": ncu.get("issue_active_pct"), "fma_pipe_active_pct_ncu": ncu.get("fma_pipe_active_pct"),
                             "ncu_capture": ncu.get("source")},
                "e2e": {"value": world * corr_per_step * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": nb * CHUNK + nb * 4, "d2h_bytes_per_step": nb * 32,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": 3 * launches_per_step * args.steps,      # fwd_kernel, cell_kernel, best_kernel per launch triple
                "clocks": clocks,
                "stage_ms": {k: round(v, 4) for k, v in stage.items()},
                "detected_prns_last_step": detected}
        if grid_c4 is not None:
            grid_c4["contract_gbs"] = grid_c4["value"] * grid_c4["bytes_per_corr"] / 1e9
            grid_c4["frac_of_hbm_peak_per_gpu"] = grid_c4["contract_gbs"] / world / peak_gbs
            line["grid_mode_configs4_sharded"] = grid_c4
        if world == 1 and not args.no_grid:
            for name in ("configs1", "configs2", "configs3"):
                line["grid_mode_" + name] = grid_measure(ga, dev, peak_gbs, name)
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_baseline(runs_per_core=4)
            os.unlink(cpu_sample_file())
            line["cpu_baseline"] = cb
        if json_fd is None:
            print(json.dumps(line))
        else:
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(line) + "\n").encode())
    acq.close()
    if world > 1:
        dist.destroy_process_group()


def grid_c4_sharded(ga, dev, rank, world, dist):
    """Secondary: BASELINE.json configs[4] -- ONE acquisition of 32 PRN x 2001 Doppler bins (+-100 kHz @ 100 Hz),
    fs = 8.184 MHz, 10 ms non-coherent (640,320 coherent correlations), its Doppler grid split into contiguous
    shards over the ranks (cfg.dop_first/dop_count), one NCCL all-gather of the 32 x 32-byte peak records per
    acquisition, merged by max snr / lower bin.  STRONG scaling: total work fixed.  Every rank reads the same
    10,230-byte input.  Timed with CUDA events on the launching stream, max over ranks."""
    import importlib
    import torch
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    shard = importlib.import_module("gnss_gps_sdr_b200.shard")
    fs, fc, max_fo, step, K = 8.184e6, 2.046e6, 100000.0, 100.0, 10
    W = int(round(fs / 1000))
    nbins = 2 * int(max_fo // step) + 1
    sats = sg.default_constellation(fs, cn0_dbhz=50.0, seed=1575420002, max_doppler=0.9 * max_fo)
    bits = sg.synth_capture(W * K, fs, fc, sats, seed=4)
    lo, n = shard.bin_range(nbins, rank, world)
    acq = ga.Acquisition(fc, fs, max_fo, device=dev.index, mode=1, doppler_step=step, noncoh_blocks=K, max_blocks=1,
                         dop_first=lo, dop_count=n)
    d_bits = torch.from_numpy(bits).to(dev)
    d_out = torch.zeros(32 * 32, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(world * 32 * 32, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)          # kernels, NCCL and the timing events all on this stream
    torch.cuda.set_stream(stream)
    acq.set_stream(stream.cuda_stream)

    def one():
        acq.acquire_device(d_bits.data_ptr(), 1, d_out.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_out)

    for _ in range(3):
        one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        one()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    rec = np.frombuffer((d_all if world > 1 else d_out).cpu().numpy().tobytes(), ga.PEAK_DTYPE).reshape(world, 32)
    merged = shard.merge_peaks(rec)
    found = sum(1 for s_ in sats if abs(merged[s_["prn"] - 1]["lo_shift"] * step - s_["doppler_hz"]) <= step and merged[s_["prn"] - 1]["snr"] >= 25)
    out = {"workload": "GRID C4: fs=8.184MHz, 32 PRN x 2001 bins (+-100 kHz @ 100 Hz), 10 ms non-coherent, ONE acquisition "
                       f"sharded by Doppler bin over {world} GPU(s), NCCL all-gather of the peak records",
           "value": 32 * nbins * K / ms * 1e3, "unit": UNIT, "scaling": "strong", "ms_per_acquisition": ms,
           "bins_per_gpu": n, "bytes_per_corr": acq.info["bytes_per_corr"], "generated_svs_found": f"{found}/{len(sats)}",
           "stage_ms_rank0": {k: round(v, 4) for k, v in acq.stage_times().items()}}
    acq.close()
    return out


GRID_CONFIGS = {
    # BASELINE.json configs[1..3] in GRID semantics: n_acq acquisitions per launch, sized so that the
    # block spectra of a launch exceed the 126 MB L2 where the shape allows it
    "configs1": dict(fs=5.456e6, fc=4.092e6, max_fo=5000.0, step=500.0, K=1, n_acq=192, seed=1575420000,
                     what="fs=5.456MHz, 32 PRN x 21 bins (+-5 kHz @ 500 Hz), 1 ms coherent"),
    "configs2": dict(fs=8.184e6, fc=2.046e6, max_fo=5000.0, step=500.0, K=1, n_acq=128, seed=1575420000,
                     what="fs=8.184MHz (gps_sig_gen.m rate), 32 PRN x 21 bins (+-5 kHz @ 500 Hz), 1 ms coherent"),
    "configs3": dict(fs=2.8e6, fc=0.62e6, max_fo=100000.0, step=250.0, K=10, n_acq=4, seed=1575420001,
                     what="fs=2.8MHz (rtl-sdr rate), 32 PRN x 801 bins (+-100 kHz @ 250 Hz), 10 ms non-coherent"),
}


def grid_measure(ga, dev, peak_gbs, name):
    """Secondary: one BASELINE.json config in GRID semantics (1 ms coherent blocks, explicit Doppler grid, K-block
    non-coherent sums, all 32 PRNs on the same blocks) -- a mode the reference does not have (no reference arm; parity
    against the oracle's definition and the reference-held vectors, tests/test_gpu_grid.py).  Device-resident input,
    CUDA events on the launching stream; contract bytes 2*W*8+16 per coherent correlation."""
    import importlib
    import torch
    sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
    c = GRID_CONFIGS[name]
    W = int(round(c["fs"] / 1000))
    sats = sg.default_constellation(c["fs"], seed=c["seed"], max_doppler=0.9 * c["max_fo"])
    acq = ga.Acquisition(c["fc"], c["fs"], c["max_fo"], device=dev.index, mode=1, doppler_step=c["step"], noncoh_blocks=c["K"],
                         max_blocks=c["n_acq"])
    n_acq = min(c["n_acq"], acq.info["max_acq"])
    bits = ga.synth_capture_gpu(W * c["K"] * n_acq, c["fs"], c["fc"], sats, seed=3, device=dev.index)
    d_bits = torch.from_numpy(bits).to(dev)
    d_out = torch.zeros(n_acq * 32 * 32, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)          # kernels and the timing events on this stream
    torch.cuda.set_stream(stream)
    acq.set_stream(stream.cuda_stream)
    for _ in range(3):
        acq.acquire_device(d_bits.data_ptr(), n_acq, d_out.data_ptr())
    reps = 20 if c["K"] == 1 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        acq.acquire_device(d_bits.data_ptr(), n_acq, d_out.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    corr = n_acq * 32 * acq.n_doppler * c["K"]
    bpc = acq.info["bytes_per_corr"]
    native = acq.info["fft_len"] == W
    out = {"workload": f"GRID {name}: {c['what']}, {n_acq} acquisitions per launch",
           "value": corr / ms * 1e3, "unit": UNIT, "ms_per_launch": ms, "bytes_per_corr": bpc,
           "contract_gbs": corr * bpc / ms / 1e6, "frac_of_hbm_peak": corr * bpc / ms / 1e6 / peak_gbs,
           "stage_ms": {k: round(v, 4) for k, v in acq.stage_times().items()},
           "note": (f"native {W}-point prime-factor transform (DESIGN.md section 10)" if native else
                    f"zero-padded embedding of the {W}-point correlation (DESIGN.md section 10)")}
    acq.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-grid", action="store_true")
    ap.add_argument("--cpu-worker")
    ap.add_argument("--cpu-kind", default="reference")
    ap.add_argument("--cpu-runs", type=int, default=2)
    ap.add_argument("--cpu-index", type=int, default=0)
    args = ap.parse_args()
    if args.cpu_worker:
        return cpu_worker(args)
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

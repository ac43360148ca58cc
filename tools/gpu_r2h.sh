#!/bin/bash
# round 2, session h: 160-thread CTAs (5 warps x 3 CTAs/SM), GRID ticket scheduler A/B, full GPU test suite, bench
mkdir -p gpurun_out
GPSACQ_LIB=build/variants/t160.so timeout 60 python tools/tma_probe.py 64 2>&1 | tail -1
GPSACQ_LIB=build/variants/t160.so timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
for c in C1 C2 C3 C4; do
  timeout 120 python tools/bench_grid.py $c 2>&1 | python tools/grid_line.py | sed "s/^/ticket /"
  GPSACQ_STATIC_SCHED=1 timeout 120 python tools/bench_grid.py $c 2>&1 | python tools/grid_line.py | sed "s/^/static /"
done
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err

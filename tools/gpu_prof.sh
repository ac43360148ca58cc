#!/bin/bash
mkdir -p gpurun_out
echo "nproc=$(nproc) cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null) affinity=$(python -c 'import os;print(len(os.sched_getaffinity(0)))')"; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket'; free -g | head -2
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
ncu --set full --clock-control none --import-source on -k regex:cell_kernel -s 3 -c 1 -o gpurun_out/cell_norot -f python tools/quick_bench.py > gpurun_out/ncu_norot.log 2>&1
GPSACQ_LIB=build/variants/rot128.so ncu --set full --clock-control none --import-source on -k regex:cell_kernel -s 3 -c 1 -o gpurun_out/cell_rot128 -f python tools/quick_bench.py > gpurun_out/ncu_rot128.log 2>&1
ls -la gpurun_out

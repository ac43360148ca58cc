#!/bin/bash
mkdir -p gpurun_out
python tools/quick_bench.py 2>&1 | tail -1
for v in build/variants/*.so; do GPSACQ_LIB=$v python tools/quick_bench.py 2>&1 | tail -1; done
python tools/quick_bench.py 2.046e6 8.184e6 2>&1 | tail -1
GPSACQ_LIB=build/variants/noperm.so python tools/quick_bench.py 2.046e6 8.184e6 2>&1 | tail -1
python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
ncu --set full --clock-control none --import-source on -k regex:cell_kernel_tm -s 6 -c 1 -o gpurun_out/cell_prof_r02 -f python tools/quick_bench.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log

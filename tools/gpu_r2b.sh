#!/bin/bash
mkdir -p gpurun_out
build/warpmap > gpurun_out/warpmap.txt 2>&1; tail -12 gpurun_out/warpmap.txt
python tools/quick_bench.py 2>&1 | tail -1
for v in build/variants/*.so; do GPSACQ_LIB=$v python tools/quick_bench.py 2>&1 | tail -1; done
python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
grep "whole-capture" gpurun_out/pytest_gpu.log
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err

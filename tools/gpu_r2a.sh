#!/bin/bash
# round-2 session A: parity first (whole capture), then variant timing, then the bench
mkdir -p gpurun_out
ls -la oracle/_ref/data/ > gpurun_out/staged.txt 2>&1
python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
grep "whole-capture" gpurun_out/pytest_gpu.log
python tools/quick_bench.py 2>&1 | tail -1
for v in build/variants/*.so; do GPSACQ_LIB=$v python tools/quick_bench.py 2>&1 | tail -1; done
python tools/quick_bench.py 2.046e6 8.184e6 2>&1 | tail -1
python tools/quick_bench.py 0.62e6 2.8e6 2>&1 | tail -1
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err

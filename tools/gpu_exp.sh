#!/bin/bash
mkdir -p gpurun_out
python tools/quick_bench.py 2>&1 | tail -1
for v in build/variants/*.so; do GPSACQ_LIB=$v python tools/quick_bench.py 2>&1 | tail -1; done
python tools/quick_bench.py 2.046e6 8.184e6 2>&1 | tail -1
python tools/quick_bench.py 0.62e6 2.8e6 2>&1 | tail -1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log

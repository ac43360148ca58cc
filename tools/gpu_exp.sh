#!/bin/bash
mkdir -p gpurun_out
python tools/quick_bench.py 2>&1 | tail -1
for v in build/variants/*.so; do GPSACQ_LIB=$v python tools/quick_bench.py 2>&1 | tail -1; done
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log

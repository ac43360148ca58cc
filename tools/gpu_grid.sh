#!/bin/bash
# GRID mode on the GPU: parity tests, then throughput of the native transforms next to the embedding
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grid.py -x -q -m gpu > gpurun_out/pytest_grid.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_grid.log; tail -15 gpurun_out/pytest_grid.log
python tools/bench_grid.py 2>&1 | tee gpurun_out/bench_grid_native.jsonl | cut -c1-420
GPSACQ_GRID_EMBED=1 python tools/bench_grid.py C1 2>&1 | cut -c1-420

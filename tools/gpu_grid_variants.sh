#!/bin/bash
# GRID-mode tuning session: parity tests on the default build, then throughput of build/variants/*.so
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grid.py -x -q -m gpu > gpurun_out/pytest_grid.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_grid.log; tail -3 gpurun_out/pytest_grid.log
for v in build/variants/*.so; do echo "== $v"; GPSACQ_LIB=$v python tools/bench_grid.py 2>&1 | python tools/grid_line.py; done

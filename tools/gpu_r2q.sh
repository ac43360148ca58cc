#!/bin/bash
# round 2, session q: ticket scheduler in grid_cell_kernel (exact-length and embedding GRID paths): tests + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grid.py -x -q -m gpu > gpurun_out/pytest_grid.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_grid.log
timeout 120 python tools/bench_grid.py X10 2>&1 | python tools/grid_line.py | sed "s/^/ticket /"
GPSACQ_STATIC_SCHED=1 timeout 120 python tools/bench_grid.py X10 2>&1 | python tools/grid_line.py | sed "s/^/static /"
GPSACQ_GRID_EMBED=1 timeout 120 python tools/bench_grid.py C1 2>&1 | python tools/grid_line.py | sed "s/^/embed ticket /"
GPSACQ_GRID_EMBED=1 GPSACQ_STATIC_SCHED=1 timeout 120 python tools/bench_grid.py C1 2>&1 | python tools/grid_line.py | sed "s/^/embed static /"

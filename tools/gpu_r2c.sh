#!/bin/bash
mkdir -p gpurun_out
python tools/quick_bench.py 2>&1 | tail -1
for v in build/variants/twb.so build/variants/prea.so build/variants/twb_prea.so; do GPSACQ_LIB=$v python tools/quick_bench.py 2>&1 | tail -1; done
python tools/bench_grid.py C1 C2 2>&1 | python tools/grid_line.py
GPSACQ_LIB=build/variants/pfa96.so python tools/bench_grid.py C1 C2 2>&1 | python tools/grid_line.py
python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err

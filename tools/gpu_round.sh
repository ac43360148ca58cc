#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), launch list + full ncu capture of the cell kernels.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-grid > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_kernel_tm -s 4 -c 1 -o gpurun_out/cell_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-grid > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pfa_cell_kernel -s 3 -c 1 -o gpurun_out/pfa_cell_c1 -f python tools/bench_grid.py C1 > gpurun_out/ncu_pfa_c1.log 2>&1
fi
ls -la gpurun_out | head -30

#!/bin/bash
# round 2, session p: sub-partition balance A/B, then the full evidence run (tests, bench both arms, launch list, ncu captures)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-160 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-grid > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_kernel_tm -s 2 -c 1 -o gpurun_out/cell_prof_r02 -f python tools/launch_sweep.py 512 512 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:pfa_cell_kernel -s 3 -c 1 -o gpurun_out/pfa_cell_c1_r02 -f python tools/bench_grid.py C1 > gpurun_out/ncu_pfa_c1.log 2>&1; tail -1 gpurun_out/ncu_pfa_c1.log
ncu --set full --clock-control none --import-source on -k regex:pfa_cell_kernel -s 3 -c 1 -o gpurun_out/pfa_cell_c2_r02 -f python tools/bench_grid.py C2 > gpurun_out/ncu_pfa_c2.log 2>&1; tail -1 gpurun_out/ncu_pfa_c2.log
du -sh gpurun_out

#!/bin/bash
# round 2, session w: final check of the committed tree -- sanitizer, all GPU tests, smoke, bench (both arms), launch list
mkdir -p gpurun_out
bash tools/gpu_sanitize.sh | grep -E "^==|SUMMARY"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log | cut -c1-120
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-160 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-120 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-grid > gpurun_out/bench_under_ncu.log 2>&1

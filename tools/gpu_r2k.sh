#!/bin/bash
# round 2, session k: compute-sanitizer over every kernel family (incl. ticket scheduler, TMA kernel, table-driven forward
# kernel), stream converter throughput
mkdir -p gpurun_out
python tools/bench_frontend.py > gpurun_out/bench_frontend.jsonl 2> gpurun_out/bench_frontend.err; cat gpurun_out/bench_frontend.jsonl; tail -2 gpurun_out/bench_frontend.err
bash tools/gpu_sanitize.sh

#!/bin/bash
# round 2, session l: stream converters after the wide-load rewrite (tests + throughput), new parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_siggen.py -x -q -m gpu > gpurun_out/pytest_frontend.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_frontend.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pinned or ticket or tma" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
python tools/bench_frontend.py > gpurun_out/bench_frontend.jsonl 2> gpurun_out/bench_frontend.err; cat gpurun_out/bench_frontend.jsonl; tail -2 gpurun_out/bench_frontend.err

#!/bin/bash
# round 2, session x (the last GPU minutes of the round): second-generation stream converters -- parity (both kernel
# generations through every front-end test, v1 == v2 bit for bit, SHA-256 against the reference program), throughput of both,
# the packed-FP32 micro-benchmark, and -- only if there is time left -- one ncu capture of the threshold kernel.
mkdir -p gpurun_out
timeout 20 tools/ubench/fp32x2 > gpurun_out/fp32x2.txt 2>&1; echo "fp32x2 rc=$? t=$SECONDS"
timeout 150 python -m pytest tests/test_gpu_frontend.py -x -q > gpurun_out/pytest_frontend.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -3 gpurun_out/pytest_frontend.log
timeout 90 python tools/bench_frontend.py > gpurun_out/bench_frontend_v12.jsonl 2> gpurun_out/bench_frontend.err; echo "bench rc=$? t=$SECONDS"; cat gpurun_out/bench_frontend_v12.jsonl | cut -c1-170
if [ $SECONDS -lt 170 ]; then
  GPSACQ_NCU=1 timeout 80 ncu --set full --clock-control none --import-source on -k regex:iq8_to_bits_thr -c 1 -o gpurun_out/iq8_thr_r02 -f python tools/bench_frontend.py > gpurun_out/ncu_iq8_thr.log 2>&1; echo "ncu rc=$? t=$SECONDS"
fi

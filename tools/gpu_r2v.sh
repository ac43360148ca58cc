#!/bin/bash
# round 2, session v: forward kernel with the chunk staged by a TMA bulk copy, A/B; parity; other rates
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
GPSACQ_LIB=build/variants/nobulk.so timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
done
export SWEEP_FC=0.62e6 SWEEP_FS=2.8e6
timeout 120 python tools/launch_sweep.py 1024 1024 2>&1 | grep -E "sub=|rror"
GPSACQ_LIB=build/variants/nobulk.so timeout 120 python tools/launch_sweep.py 1024 1024 2>&1 | grep -E "sub=|rror"

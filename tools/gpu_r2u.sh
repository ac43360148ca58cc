#!/bin/bash
# round 2, session u: branch-free peak tracking in pass C, A/B twice; parity
for i in 1 2; do
timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
GPSACQ_LIB=build/variants/branchy.so timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2

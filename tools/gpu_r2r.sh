#!/bin/bash
# round 2, session r: barrier merge and unrolled B/C task loops, A/B twice each
for i in 1 2; do
timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
for v in nomerge unrollbc; do GPSACQ_LIB=build/variants/$v.so timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"; done
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not whole" 2>&1 | tail -2

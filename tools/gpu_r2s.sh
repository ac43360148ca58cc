#!/bin/bash
# round 2, session s: unrolled task loops (passes B/C default, pass A variant), other geometries, wide-window CTA shape
timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
for v in unrolla nounroll; do GPSACQ_LIB=build/variants/$v.so timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"; done
export SWEEP_FC=0.62e6 SWEEP_FS=2.8e6
timeout 120 python tools/launch_sweep.py 1024 1024 2>&1 | grep -E "sub=|rror"
GPSACQ_LIB=build/variants/nounroll.so timeout 120 python tools/launch_sweep.py 1024 1024 2>&1 | grep -E "sub=|rror"
export SWEEP_FC=1.75e6 SWEEP_FS=7e6
timeout 120 python tools/launch_sweep.py 2048 2048 2>&1 | grep -E "sub=|rror"
GPSACQ_LIB=build/variants/wide256.so timeout 120 python tools/launch_sweep.py 2048 2048 2>&1 | grep -E "sub=|rror"

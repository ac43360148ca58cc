#!/bin/bash
# round 2, session o: CTA shapes of the 10 x 4000 REF geometry (fs = 2.8 MHz) under the ticket scheduler; 8.184 MHz for the record
export SWEEP_FC=0.62e6 SWEEP_FS=2.8e6
timeout 120 python tools/launch_sweep.py 1024 1024 2>&1 | grep -E "sub=|rror"
for v in g4t160 g4t224; do GPSACQ_LIB=build/variants/$v.so timeout 120 python tools/launch_sweep.py 1024 1024 2>&1 | grep -E "sub=|rror"; done
export SWEEP_FC=2.046e6 SWEEP_FS=8.184e6
timeout 120 python tools/launch_sweep.py 2048 2048 2>&1 | grep -E "sub=|rror"

#!/bin/bash
mkdir -p gpurun_out
GPSACQ_LIB=build/variants/xldg.so timeout 60 python tools/tma_probe.py 64 2>&1 | tail -1
GPSACQ_CELL_TMA=1 GPSACQ_LIB=build/variants/xldg.so timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
GPSACQ_CELL_TMA=1 timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
timeout 120 python tools/launch_sweep.py 3584 3584 2>&1 | grep -E "sub=|rror"
ncu --set full --clock-control none --import-source on -k regex:cell_kernel_tm -s 2 -c 1 -o gpurun_out/cell_prof_r02 -f python tools/launch_sweep.py 512 512 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
GPSACQ_CELL_TMA=1 ncu --set full --clock-control none --import-source on -k regex:cell_kernel_tma -s 2 -c 1 -o gpurun_out/cell_tma_prof_r02 -f python tools/launch_sweep.py 512 512 > gpurun_out/ncu_full_tma.log 2>&1; tail -2 gpurun_out/ncu_full_tma.log

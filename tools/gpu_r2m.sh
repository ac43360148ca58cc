#!/bin/bash
# round 2, session m: CTA shapes of the K = 1 GRID kernels under the ticket scheduler
for c in C1 C2; do timeout 120 python tools/bench_grid.py $c 2>&1 | python tools/grid_line.py | sed "s/^/default /"; done
for v in c2t160 c2t192; do GPSACQ_LIB=build/variants/$v.so timeout 120 python tools/bench_grid.py C2 2>&1 | python tools/grid_line.py | sed "s/^/$v /"; done
for v in c1t160b3 c1t96b5; do GPSACQ_LIB=build/variants/$v.so timeout 120 python tools/bench_grid.py C1 2>&1 | python tools/grid_line.py | sed "s/^/$v /"; done

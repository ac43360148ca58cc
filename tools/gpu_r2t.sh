#!/bin/bash
# round 2, session t: unrolled task loops in pfa_cell_kernel (GRID), all four configs
for c in C1 C2 C3 C4; do
  timeout 120 python tools/bench_grid.py $c 2>&1 | python tools/grid_line.py | sed "s/^/default /"
  for v in pub puc pubc; do GPSACQ_LIB=build/variants/$v.so timeout 120 python tools/bench_grid.py $c 2>&1 | python tools/grid_line.py | sed "s/^/$v /"; done
done

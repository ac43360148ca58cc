#!/bin/bash
# round 2, session z: the same scalar-FADD question for the GRID prime-factor kernel (FMA pipe 71 % busy there):
# radix-16 = 4 x 4 with scalar radix-4 adds (GA_R4_SCALAR), prime radices with scalar input / output sums (GA_RP_SCALAR)
mkdir -p gpurun_out
for v in "" _r4s _rp3 _r4rp3 _r4rp1; do
  lib=gnss-gps-sdr_b200/csrc/libgpsacq$v.so
  [ -f $lib ] || continue
  GPSACQ_LIB=$PWD/$lib timeout 60 python tools/bench_grid.py C1 C2 C4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('$v', d['config'], '%.3f ms' % d['ms_per_batch'], '%.2f Mcorr/s' % (d['correlations_per_s'] / 1e6), 'frac %.3f' % d['frac_of_hbm_peak'], 'detected', d['detected'])
" | tee -a gpurun_out/scalar_adds_grid.txt
done
echo "t=$SECONDS"

#!/bin/bash
# round 2, session zz (last seconds of GPU budget): peak records of the GA_RP_SCALAR=3 build against the default build,
# hashed, on all four GRID configs (every prime radix: 3, 7, 11, 31) -- identical arithmetic, so identical bytes expected
mkdir -p gpurun_out
for v in "" _rp3; do
  GPSACQ_LIB=$PWD/gnss-gps-sdr_b200/csrc/libgpsacq$v.so timeout 50 python tools/bench_grid.py C1 C2 C3 C4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('$v', d['config'], '%.3f ms' % d['ms_per_batch'], '%.2f Mcorr/s' % (d['correlations_per_s'] / 1e6), 'frac %.3f' % d['frac_of_hbm_peak'], 'detected', d['detected'], 'sha', d['peaks_sha256'])
" | tee -a gpurun_out/scalar_adds_grid_hash.txt
done
echo "t=$SECONDS"

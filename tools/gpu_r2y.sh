#!/bin/bash
# round 2, session y: scalar FADDs (FMA-lite pipe) instead of FADD2 in the radix-4 / radix-5 butterflies of the REF cell
# kernel -- builds made with -DGA_R4_SCALAR / -DGA_R5_SCALAR, same batch, same checksum expected (identical arithmetic)
mkdir -p gpurun_out
for v in "" _r4s _r5a _r5b _r45; do
  lib=gnss-gps-sdr_b200/csrc/libgpsacq$v.so
  [ -f $lib ] || continue
  GPSACQ_LIB=$PWD/$lib timeout 60 python tools/launch_sweep.py 3584 3584 2>&1 | tail -1 | tee -a gpurun_out/scalar_adds.txt
done
echo "t=$SECONDS"

#!/bin/bash
# tools/build_variant.sh NAME "-DKNOB=..."  -> build/variants/NAME.so (a libgpsacq build with other -D knobs; load with GPSACQ_LIB)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas --expt-relaxed-constexpr $2 \
  -shared -o build/variants/$1.so gnss-gps-sdr_b200/csrc/gpsacq.cu -ldl

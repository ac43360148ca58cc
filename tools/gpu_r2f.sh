#!/bin/bash
# round 2, session f: ticket scheduler, launch-size sweep, CTA-shape variants
mkdir -p gpurun_out
python tools/launch_sweep.py 3584 128,3584 2>&1 | grep sub=
for v in build/variants/*.so; do GPSACQ_LIB=$v python tools/launch_sweep.py 3584 3584 2>&1 | grep sub=; done

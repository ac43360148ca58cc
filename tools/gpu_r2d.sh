#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
for sb in 128 64 256 512 7168; do
  GPSACQ_SUB_BLOCKS=$sb python bench.py --steps 10 --warmup 3 --no-grid --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('sub_blocks $sb value %.3f M e2e %.3f M frac %.3f whole %.3f launch_ms %.3f clocks %s'%(d['value']/1e6,d['e2e']['value']/1e6,r['frac'],r['whole_step_frac'],r['launch_ms'],d['clocks']))"
done
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err

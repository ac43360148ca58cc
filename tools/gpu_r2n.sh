#!/bin/bash
# round 2, session n (8 GPUs): bench under torchrun at N = 8 and N = 4, one-process group API at 8 GPUs
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  python - <<P
import json
d=json.load(open('gpurun_out/bench_n$n.json'))
print('N=$n value %.3f M e2e %.3f M ms/step %.3f frac %.3f clocks %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['clocks']), 'C4 sharded %.2f M' % (d.get('grid_mode_configs4_sharded',{}).get('value',0)/1e6))
P
done
timeout 300 python tools/bench_group.py 8 10 2>&1 | tail -1 | tee gpurun_out/bench_group_n8.json

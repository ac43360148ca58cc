#!/bin/bash
# round 2, session j (2 GPUs): group API / NCCL tests, bench under torchrun
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "group or shard or nccl or multi" > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-400 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('N=2 value %.3f M e2e %.3f M ms/step %.3f frac %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['roofline']['frac']), d.get('grid_mode_configs4_sharded',{}).get('value'))
P

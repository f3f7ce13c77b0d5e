#!/bin/bash
# Host threads of the drain (range pool) on the shape of a many-rank run: 2 ranks on 8 host cores (4 per rank, like 8 ranks on
# 32), weak leg 4096 channels per GPU + strong leg 2048 per GPU with the weak leg's decoder still alive.  Run with --gpus 2.
for ht in 1 2; do
  HBD_HOST_THREADS=$ht taskset -c 0-7 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + ht)) \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-wideband --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d['per_rank']; s=d['strong']
print('host_threads $ht: weak ms/step %.4f per-rank %s replay %s | strong ms/step %.4f per-rank %s drain/step %s' % (d['ms_per_step'], [round(x,3) for x in p['ms']], [round(x,3) for x in p['host_replay_ms_total']], s['ms_per_step'], [round(x,3) for x in s['per_rank_ms']], [round(x,4) for x in s['per_rank_host_drain_ms_per_step']]))"
done

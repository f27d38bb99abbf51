#!/bin/bash
# Round 2, GPU call W (eight B200s): the bench line at N = 8 with the final kernels (four-step passes on 2048- and 4096-wide
# slabs across ranks).
O=gpurun_out; T=r02w
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > $O/${T}_bench_n8.json 2> $O/${T}_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench_n8.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d.get('halo_wait'))
print(d['e2e']['value'], d['e2e']['ms_per_step'])
print(json.dumps(d['extra']['size_4096'])[:600])
PY
tail -3 $O/${T}_bench_n8.err

#!/bin/bash
# Round 2, GPU call U (two B200s): the final kernels across real ranks — NCCL / CUDA-IPC parity of run_slab_ranks.py (incl.
# four-step passes on 1024-, 2048- and 4096-wide slabs) and the bench line at N = 2 (2048^3 and the 4096^3 block).
O=gpurun_out; T=r02u
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_slab_ranks.py > $O/${T}_slab_ranks_n2.log 2>&1; echo "rc=$?" >> $O/${T}_slab_ranks_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err
tail -4 $O/${T}_slab_ranks_n2.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench_n2.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d.get('halo_wait'))
print(json.dumps(d['extra']['size_4096'])[:500])
PY

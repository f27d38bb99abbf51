#!/bin/bash
# Round 2, GPU call I (2 B200s): four-step passes across ranks (CUDA IPC, NVLink) — parity and a first timing.
O=gpurun_out; T=r02i
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29611 tests/run_slab_ranks.py > $O/${T}_slab_ranks_n2.log 2>&1; echo "rc=$?" >> $O/${T}_slab_ranks_n2.log
timeout 600 $TR --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 0 --big-steps 0 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n2.json
timeout 300 python tools/inproc_scale.py 2 2048 > $O/${T}_inproc_scale.txt 2>&1
ls -la $O | tail -5

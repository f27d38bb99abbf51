#!/bin/bash
# Round 2, GPU call L (8 B200s of one box): the final multi-GPU evidence — four-step passes across ranks, edge bands late
# (A/B against FS3D_S4_EDGE_EARLY=1 and against two-step passes).
# NCCL / IPC parity on 4 ranks, strong scaling at 2048^3 with the 4096^3 block and per-rank e2e, config 5 with the
# ray-march jumps, the in-process world.  ONE call (charged 8x).
O=gpurun_out; T=r02l
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 4 --master-port 29611 tests/run_slab_ranks.py > $O/${T}_slab_ranks_n4.log 2>&1; echo "rc=$?" >> $O/${T}_slab_ranks_n4.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1.json 2> $O/${T}_bench.err
for N in 2 4 8; do
  timeout 900 $TR --nproc-per-node $N --master-port 2950$N bench.py --gpus $N --steps 20 --warmup 5 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n$N.json
done
timeout 600 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 100 --warmup 3 --e2e-steps 0 --big-steps 0 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n8_100steps.json
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --impl reference 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_reference_n8.json
timeout 600 $TR --nproc-per-node 8 --master-port 29530 tools/run_configs.py 5 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_config5_n8.json
for N in 2 8; do timeout 300 python tools/inproc_scale.py $N 2048 >> $O/${T}_inproc_scale.txt 2>&1; done
timeout 300 python tools/inproc_scale.py 8 4096 >> $O/${T}_inproc_scale.txt 2>&1
FS3D_NO_FUSE4=1 timeout 600 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 5 --e2e-steps 0 --big-steps 0 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n8_two_steps_per_pass.json
for N in 4 8; do
FS3D_S4_EDGE_EARLY=1 timeout 600 $TR --nproc-per-node $N --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 0 --big-steps 0 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n${N}_edge_early.json
done
ls -la $O | tail -14

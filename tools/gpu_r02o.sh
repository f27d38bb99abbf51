#!/bin/bash
# Round 2, GPU call O (8 B200s): strong scaling of the tuned four-step kernel (P = 4, K = 4, 384 threads), device-timed lines only.
O=gpurun_out; T=r02o
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
X="--steps 20 --warmup 5 --e2e-steps 0 --big-steps 0 --m8-steps 0 --no-cpu"
python bench.py $X > $O/${T}_bench_n1.json 2> $O/${T}_bench.err
for N in 2 4 8; do
  timeout 600 $TR --nproc-per-node $N --master-port 2950$N bench.py --gpus $N $X 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n$N.json
done
timeout 300 python tools/inproc_scale.py 8 2048 > $O/${T}_inproc_scale.txt 2>&1

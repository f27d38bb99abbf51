#!/bin/bash
# usage: tools_scale.sh SIZE N [extra bench args]  -> one JSON line (rank 0)
SIZE=$1; N=$2; shift 2
if [ "$N" = "1" ]; then
  python bench.py --size $SIZE --gpus 1 --no-cpu --e2e-steps 0 "$@" 2>&1 | tail -1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29$((500+N)) bench.py --size $SIZE --gpus $N --no-cpu --e2e-steps 0 "$@" 2>&1 | grep '^{' | tail -1
fi

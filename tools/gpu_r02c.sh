#!/bin/bash
# Round 2, GPU call C (8 B200s of one box): NCCL / IPC parity, strong scaling at 2048^3 with the 4096^3 block and the
# per-rank end-to-end step, config 5, the in-process world.  Everything in ONE call (charged 8x).
O=gpurun_out; T=r02c
mkdir -p $O
nvidia-smi topo -m > $O/${T}_topo.txt 2>&1; nproc >> $O/${T}_topo.txt; numactl -H >> $O/${T}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 4 --master-port 29611 tests/run_slab_ranks.py > $O/${T}_slab_ranks_n4.log 2>&1; echo "rc=$?" >> $O/${T}_slab_ranks_n4.log
timeout 600 python -m pytest tests/test_multislab_gpu.py tests/test_skip_gpu.py -m gpu -x -q > $O/${T}_pytest_multi.log 2>&1; echo "rc=$?" >> $O/${T}_pytest_multi.log
python bench.py --steps 20 --warmup 5 --no-cpu > $O/${T}_bench_n1.json 2> $O/${T}_bench.err
for N in 2 4 8; do
  timeout 900 $TR --nproc-per-node $N --master-port 2950$N bench.py --gpus $N --steps 20 --warmup 5 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n$N.json
done
timeout 600 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 100 --warmup 3 --e2e-steps 0 --big-steps 0 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_bench_n8_100steps.json
timeout 600 $TR --nproc-per-node 8 --master-port 29530 tools/run_configs.py 5 2>> $O/${T}_bench.err | grep '^{' | tail -1 > $O/${T}_config5_n8.json
for N in 2 8; do timeout 300 python tools/inproc_scale.py $N 2048 >> $O/${T}_inproc_scale.txt 2>&1; done
timeout 300 python tools/inproc_scale.py 8 4096 >> $O/${T}_inproc_scale.txt 2>&1
# PUSH kernel with a real NVLink peer, one process (in-process world on 2 GPUs): one pass of ncu, a few counters
ncu --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_op_write.sum,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --replay-mode application -k regex:step_kernel -c 8 --csv --log-file $O/${T}_push_2gpu_ncu.csv python tools/inproc_scale.py 2 2048 > $O/${T}_push_2gpu_ncu.log 2>&1
ls -la $O | tail -20

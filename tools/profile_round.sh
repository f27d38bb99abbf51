#!/bin/bash
# Round evidence on ONE B200 (run under gpurun): bench line, launch list, ncu --set full of the step kernels.
# usage: bash tools/profile_round.sh TAG     -> gpurun_out/TAG_*
TAG=${1:-r01c}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_2048.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu --e2e-steps 1 > $O/${TAG}_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:step_kernel"
$NCU -s 2 -c 2 -o $O/prof_${TAG}_fused_2048 python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 0 --fused-only > $O/${TAG}_ncu_a.log 2>&1
$NCU -s 4 -c 2 -o $O/prof_${TAG}_fused_4096w python tools/passtime.py 4096 4096 512 0 > $O/${TAG}_ncu_b.log 2>&1
$NCU -s 8 -c 4 -o $O/prof_${TAG}_single_4096w python tools/passtime.py 4096 4096 512 2 > $O/${TAG}_ncu_c.log 2>&1
$NCU -s 8 -c 4 -o $O/prof_${TAG}_single_2048 python tools/passtime.py 2048 2048 2048 2 > $O/${TAG}_ncu_d.log 2>&1
for f in fused_2048 fused_4096w single_4096w single_2048; do
  ncu -i $O/prof_${TAG}_$f.ncu-rep --page raw --csv > $O/${TAG}_step_kernel_${f}_ncu_full_raw.csv 2>/dev/null
done
rm -f $O/prof_${TAG}_single_4096w.ncu-rep $O/prof_${TAG}_single_2048.ncu-rep $O/prof_${TAG}_fused_2048.ncu-rep
ls -la $O

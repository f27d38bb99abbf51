#!/bin/bash
# Round 2, GPU call A (one B200): parity suite, bench (both arms), config 3 skipping, ncu launch list + full captures
# of the fused step kernel, a SKIP kernel in its sparse phase, a PUSH kernel and the ray-march kernel.
O=gpurun_out; T=r02a
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/${T}_gpu.txt; nproc >> $O/${T}_gpu.txt; free -g >> $O/${T}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2>> $O/${T}_bench_n1.err
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2>> $O/${T}_bench_n1.err
python tools/skip_sparse.py 4000 > $O/${T}_skip_sparse.json 2> $O/${T}_skip_sparse.err
python tools/raymarch_time.py > $O/${T}_raymarch_time.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${T}_launches_2048.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu --e2e-steps 1 --big-steps 0 > $O/${T}_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:step_kernel -s 2 -c 2 -o $O/prof_${T}_fused_2048 python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 0 --fused-only --big-steps 0 > $O/${T}_ncu_a.log 2>&1
$NCU --profile-from-start off -c 4 -o $O/prof_${T}_skip_sparse python tools/skip_sparse.py 4000 --profile > $O/${T}_ncu_b.log 2>&1
$NCU --profile-from-start off -k regex:step_kernel -c 4 -o $O/prof_${T}_push_1gpu python tools/push_one_gpu.py > $O/${T}_ncu_c.log 2>&1
$NCU -k regex:raymarch_kernel -c 2 -o $O/prof_${T}_raymarch python tools/raymarch_time.py > $O/${T}_ncu_d.log 2>&1
for f in fused_2048 skip_sparse push_1gpu raymarch; do
  ncu -i $O/prof_${T}_$f.ncu-rep --page raw --csv > $O/${T}_${f}_ncu_full_raw.csv 2>/dev/null
done
rm -f $O/prof_${T}_fused_2048.ncu-rep $O/prof_${T}_push_1gpu.ncu-rep
ls -la $O

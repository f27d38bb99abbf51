#!/bin/bash
# Round 2, GPU call N (one B200): the four-step kernel with its tuned shape (P = 4, K = 4, 384 threads): whole suite, smoke, bench lines, ncu.
O=gpurun_out; T=r02n
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench.err
python bench.py > $O/${T}_bench_n1.json 2>> $O/${T}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
NCU="ncu --set full --clock-control none --import-source on -k regex:step4_kernel"
$NCU -s 1 -c 2 -o $O/prof_${T}_fused4_2048 python tools/passtime4.py 2048 > $O/${T}_ncu_a.log 2>&1
$NCU -s 1 -c 2 -o $O/prof_${T}_fused4_1024 python tools/passtime4.py 1024 > $O/${T}_ncu_b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${T}_launches_2048.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu --e2e-steps 1 --big-steps 0 --m8-steps 0 > $O/${T}_ncu_launches.log 2>&1
for f in fused4_2048 fused4_1024; do
  ncu -i $O/prof_${T}_$f.ncu-rep --page raw --csv > $O/${T}_${f}_ncu_full_raw.csv 2>/dev/null
  rm -f $O/prof_${T}_$f.ncu-rep
done
ls -la $O | tail -12

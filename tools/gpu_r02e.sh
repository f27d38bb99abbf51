#!/bin/bash
# Round 2, GPU call E (one B200): full parity suite after the ray-march skipping fix and fs3d_step_host_packed, ray-march
# timings + ncu page, bench line with the packed end-to-end leg, smoke.
O=gpurun_out; T=r02e
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python tools/raymarch_time.py 2048 2 > $O/${T}_raymarch_time_2048_mixed.txt 2>&1
python tools/raymarch_time.py 2048 3 > $O/${T}_raymarch_time_2048_random.txt 2>&1
python tools/raymarch_time.py 1024 1 > $O/${T}_raymarch_time_1024_sandblock.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raymarch_kernel|brick_build" -c 6 -o $O/prof_${T}_raymarch python tools/raymarch_time.py 2048 2 > $O/${T}_ncu_a.log 2>&1
ncu -i $O/prof_${T}_raymarch.ncu-rep --page raw --csv > $O/${T}_raymarch_ncu_full_raw.csv 2>/dev/null
rm -f $O/prof_${T}_raymarch.ncu-rep
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
ls -la $O | tail

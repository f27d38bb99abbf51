#!/bin/bash
# Round 2, GPU call B (one B200): full parity suite incl. schedule version 2, export / input replay tests, sparse
# skipping after the balanced plan, bench with the version-2 block, ncu of the version-2 kernels.
O=gpurun_out; T=r02b
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python tools/skip_sparse.py 4000 > $O/${T}_skip_sparse.json 2> $O/${T}_skip_sparse.err
FS3D_NO_PDL=1 python tools/skip_sparse.py 4000 > $O/${T}_skip_sparse_no_pdl.json 2>> $O/${T}_skip_sparse.err
python tools/raymarch_time.py 4096 > $O/${T}_raymarch_time_4096.txt 2>&1
python tools/raymarch_time.py 2048 2 > $O/${T}_raymarch_time_2048_mixed.txt 2>&1
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench_n1.err
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:step_kernel -c 2 -o $O/prof_${T}_m8_fused python tools/m8_profile.py 2048 > $O/${T}_ncu_a.log 2>&1
$NCU -k regex:step_kernel -c 4 -o $O/prof_${T}_m8_single python tools/m8_profile.py 2048 --single > $O/${T}_ncu_b.log 2>&1
$NCU -c 4 -o $O/prof_${T}_skip_sparse python tools/skip_sparse.py 4000 --profile > $O/${T}_ncu_c.log 2>&1
for f in m8_fused m8_single skip_sparse; do
  ncu -i $O/prof_${T}_$f.ncu-rep --page raw --csv > $O/${T}_${f}_ncu_full_raw.csv 2>/dev/null
  rm -f $O/prof_${T}_$f.ncu-rep
done
ls -la $O | tail -15

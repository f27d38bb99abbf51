#!/bin/bash
# Round 2, GPU call S (one B200): the group split with staggered units as the default for big single slabs (runtime choice),
# last CTA's span balanced: pass times, whole suite (with the forced-group parity tests), smoke, bench line, ncu capture.
O=gpurun_out; T=r02s
mkdir -p $O
{
  python tools/passtime4.py 2048; python tools/passtime4.py 1024
  echo "FS3D_S4_GROUP_SPAN=0 (per-unit split everywhere):"
  FS3D_S4_GROUP_SPAN=0 python tools/passtime4.py 2048 2>/dev/null | head -1
} > $O/${T}_passtime4.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench.err
NCU="ncu --set full --clock-control none --import-source on -k regex:step4_kernel"
$NCU -s 1 -c 2 -o $O/prof_${T}_fused4_2048 python tools/passtime4.py 2048 > $O/${T}_ncu_a.log 2>&1
ncu -i $O/prof_${T}_fused4_2048.ncu-rep --page raw --csv > $O/${T}_fused4_2048_ncu_full_raw.csv 2>/dev/null
rm -f $O/prof_${T}_fused4_2048.ncu-rep
cat $O/${T}_passtime4.txt; tail -3 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; cut -c1-330 $O/${T}_bench_n1_driverflags.json

#!/bin/bash
# Round 2, GPU call P (one B200): the ALU diet of the step kernels — block rule as an explicit 26-LOP3 network, edge bits
# exchanged byte-packed, constant right shifts issued as IMAD.HI.  A/B against the previous library (exp/libfs3d_base.so),
# the same code with plain shifts, and two band shapes; then the whole suite, smoke, the bench line and one ncu capture.
O=gpurun_out; T=r02p
mkdir -p $O
{
  echo "new code (default build: P4 K4 T384, IMAD.HI shifts):"
  python tools/passtime4.py 2048; python tools/passtime4.py 1024
  for v in base noimad P5K4T320 P6K4T256; do
    echo "$v:"
    FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 2048 2>&1 | head -1
    FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 1024 2>&1 | head -1
  done
} > $O/${T}_experiments_alu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench.err
NCU="ncu --set full --clock-control none --import-source on -k regex:step4_kernel"
$NCU -s 1 -c 2 -o $O/prof_${T}_fused4_2048 python tools/passtime4.py 2048 > $O/${T}_ncu_a.log 2>&1
ncu -i $O/prof_${T}_fused4_2048.ncu-rep --page raw --csv > $O/${T}_fused4_2048_ncu_full_raw.csv 2>/dev/null
rm -f $O/prof_${T}_fused4_2048.ncu-rep
cat $O/${T}_experiments_alu.txt; tail -3 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; cut -c1-400 $O/${T}_bench_n1_driverflags.json

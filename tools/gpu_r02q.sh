#!/bin/bash
# Round 2, GPU call Q (one B200): plain shifts as the default again (IMAD.HI lost the A/B of call P); A/B of the single
# load-issue site, the branchy STONE fix-up, the group work split with L2 eviction priorities, IMAD.HI in the hash only.
O=gpurun_out; T=r02q
mkdir -p $O
{
  echo "default build (P4 K4 T384, explicit LOP3 rule, byte-packed edges, plain shifts):"
  python tools/passtime4.py 2048; python tools/passtime4.py 1024
  for v in oi oisb grp grp0 hash; do
    echo "$v:"
    FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 2048 2>&1 | head -1
    FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 1024 2>&1 | head -1
  done
} > $O/${T}_experiments_alu2.txt 2>&1
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
for v in grp oisb; do
  FS3D_LIB=$PWD/exp/libfs3d_$v.so ncu $M --clock-control none -k regex:step4_kernel -s 1 -c 2 --csv --log-file $O/${T}_ncu_$v.csv python tools/passtime4.py 2048 > /dev/null 2>&1
done
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
FS3D_LIB=$PWD/exp/libfs3d_grp.so timeout 600 python -m pytest tests/test_step_gpu.py -m gpu -x -q > $O/${T}_pytest_grp.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_grp.log
cat $O/${T}_experiments_alu2.txt; tail -2 $O/${T}_pytest.log; tail -2 $O/${T}_pytest_grp.log; grep -v "^==" $O/${T}_ncu_grp.csv | cut -d, -f5,13- | tail -14

#!/bin/bash
# Round 2, GPU call R (one B200): the group work split of the four-step kernel with and without the y-stagger (does the
# pair two neighbouring bands both load hit L2 the second time?), against the default build.
O=gpurun_out; T=r02r
mkdir -p $O
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
{
  echo "default build:"
  python tools/passtime4.py 2048 | head -1; python tools/passtime4.py 1024 | head -1
  for v in grp1 grp2; do
    echo "$v:"
    FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 2048 2>&1 | head -1
    FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 1024 2>&1 | head -1
    FS3D_LIB=$PWD/exp/libfs3d_$v.so ncu $M --clock-control none -k regex:step4_kernel -s 1 -c 1 --csv --log-file $O/${T}_ncu_$v.csv python tools/passtime4.py 2048 > /dev/null 2>&1
    grep -v "^==" $O/${T}_ncu_$v.csv | cut -d, -f13- | tail -6
  done
} > $O/${T}_experiments_groups.txt 2>&1
FS3D_LIB=$PWD/exp/libfs3d_grp2.so timeout 600 python -m pytest tests/test_step_gpu.py -m gpu -x -q > $O/${T}_pytest_grp2.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_grp2.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
cat $O/${T}_experiments_groups.txt; tail -2 $O/${T}_pytest_grp2.log; tail -2 $O/${T}_pytest.log

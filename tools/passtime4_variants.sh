#!/bin/bash
# A/B of the four-step kernel's band size P, y-block K and CTA size (exp/libfs3d_*.so built with -DFS3D_S4_P=.. etc.)
O=gpurun_out
echo "default (P6 K8 T256):" > $O/r02m_experiments_step4.txt
python tools/passtime4.py 2048 | head -1 >> $O/r02m_experiments_step4.txt 2>&1
python tools/passtime4.py 1024 | head -1 >> $O/r02m_experiments_step4.txt 2>&1
for v in P4K4T384 P7K4T256 P6K4T256 P5K8T256 P4K8T320; do
  echo "$v:" >> $O/r02m_experiments_step4.txt
  FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 2048 2>&1 | head -1 >> $O/r02m_experiments_step4.txt
  FS3D_LIB=$PWD/exp/libfs3d_$v.so python tools/passtime4.py 1024 2>&1 | head -1 >> $O/r02m_experiments_step4.txt
done
cat $O/r02m_experiments_step4.txt

#!/bin/bash
# Round 2, GPU call F (one B200): parity suite + sparse skipping after the single-sweep plan kernel, ncu of that phase.
O=gpurun_out; T=r02f
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python tools/skip_sparse.py 4000 > $O/${T}_skip_sparse.json 2> $O/${T}_skip_sparse.err
python tools/skip_sparse.py 2000 >> $O/${T}_skip_sparse.json 2>> $O/${T}_skip_sparse.err
python tools/run_configs.py 3 > $O/${T}_config3.json 2>> $O/${T}_skip_sparse.err
ncu --set full --clock-control none --import-source on --profile-from-start off -c 4 -o $O/prof_${T}_skip_sparse python tools/skip_sparse.py 4000 --profile > $O/${T}_ncu_c.log 2>&1
ncu -i $O/prof_${T}_skip_sparse.ncu-rep --page raw --csv > $O/${T}_skip_sparse_ncu_full_raw.csv 2>/dev/null
rm -f $O/prof_${T}_skip_sparse.ncu-rep
ls -la $O | tail -8

#!/bin/bash
# Round 2, GPU call V (one B200): fs3d_step_host / fs3d_step_host_packed with four steps per call (chunks of whole bands of
# the four-step kernel): whole suite, smoke, bench line (e2e now four steps per call, the shorter forms beside it).
O=gpurun_out; T=r02v
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench.err
tail -3 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; tail -3 $O/${T}_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench_n1_driverflags.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'])
print(json.dumps(d['e2e'])[:1500])
PY

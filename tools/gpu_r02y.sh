#!/bin/bash
# Round 2, GPU call Y (one B200): schedule version 2 with the byte-packed edge exchange — whole suite and the bench line
# (extra.materials8 is the number that moves).
O=gpurun_out; T=r02y
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 120 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench.err
tail -2 $O/${T}_pytest.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench_n1_driverflags.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['config'].get('digest_check'))
print(json.dumps(d['extra']['materials8'])[:700])
PY

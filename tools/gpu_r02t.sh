#!/bin/bash
# Round 2, GPU call T (one B200): four-step passes on 4096-wide rows (four warps per band): whole suite with the new 4096
# cases, smoke, and the bench line, whose extra.size_4096 block is the 4096^3 grid on one GPU (digest after 24 steps must
# be 0xf1bbcbfe74983b25, what the two-step kernels gave at every N).
O=gpurun_out; T=r02t
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1_driverflags.json 2> $O/${T}_bench.err
tail -3 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02t_bench_n1_driverflags.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['digest_check'] if 'digest_check' in d else '')
print(json.dumps(d['extra']['size_4096'])[:700])
PY

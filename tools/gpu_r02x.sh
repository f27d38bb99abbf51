#!/bin/bash
# Round 2, GPU call X (one B200): last look at the committed tree — whole suite, smoke, compute-sanitizer memcheck over the
# four-step kernel's three row widths / two work splits / host streaming, default bench line.
O=gpurun_out; T=r02x
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_step4.py > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/${T}_memcheck.log
timeout 200 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench.err
tail -2 $O/${T}_pytest.log; tail -1 $O/${T}_smoke.log; tail -6 $O/${T}_memcheck.log; cut -c1-260 $O/${T}_bench_n1.json

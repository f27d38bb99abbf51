#!/bin/bash
# Round 2, GPU call G (one B200): the four-steps-per-pass kernel — parity subset, then timing.
O=gpurun_out; T=r02g
mkdir -p $O
timeout 1200 python -m pytest tests/test_step_gpu.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python tools/passtime4.py 2048 > $O/${T}_passtime4.txt 2>&1
python tools/passtime4.py 1024 >> $O/${T}_passtime4.txt 2>&1
ls -la $O | tail -5

#!/bin/bash
# Round 2, GPU call H (one B200): four-step passes across slabs (two ghost planes per side) on one device + whole suite.
O=gpurun_out; T=r02h
mkdir -p $O
timeout 1200 python -m pytest tests/test_push_one_gpu.py -m gpu -x -q > $O/${T}_pytest_push.log 2>&1; echo "rc=$?" >> $O/${T}_pytest_push.log
timeout 2400 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
ls -la $O | tail -4

#!/bin/bash
# One 1-GPU gpurun session for a kernel experiment: linear-train parity tests, then the A/B timing table.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q -k "train or sorted or additive or trajectory or subsample" 2>&1 | tail -15 > $O/b_pytest.log
ROWS=${ROWS:-67108864} TAG=${TAG:-exp} timeout 200 python tools/ab_train.py > $O/b_ab.log 2>&1
cat $O/b_pytest.log $O/b_ab.log

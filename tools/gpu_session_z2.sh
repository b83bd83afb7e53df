#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== training tests"; timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu > $O/z2_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/z2_pytest.log | cut -c1-250
python tools/train_bench.py 400 > $O/z2_train.txt 2>&1; echo "rc=$?"; tail -1 $O/z2_train.txt


#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/r_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/r_pytest.log | cut -c1-250
echo "== bench_hbm"; timeout 600 python tools/bench_hbm.py 1056 > $O/r_hbm.txt 2>&1; echo "rc=$?"; tail -3 $O/r_hbm.txt

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== parity tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q > $O/h_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/h_pytest.log | cut -c1-250
echo "== hbm"; python tools/bench_hbm.py 1056 2>&1 | tee $O/h_hbm.txt

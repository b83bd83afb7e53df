#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== metrics tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_replay_reference_script.py -q -k "metric or ssim or psnr or replay" > $O/j_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/j_pytest.log | cut -c1-250
echo "== hbm"; python tools/bench_hbm.py 1056 2>&1 | tee $O/j_hbm.txt

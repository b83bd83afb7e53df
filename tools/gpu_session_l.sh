#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== fft / kspace tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_replay_reference_script.py -q -k "fft or kspace or front_end or replay or metric" > $O/l_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/l_pytest.log | cut -c1-250
echo "== hbm"; python tools/bench_hbm.py 1056 2>&1 | tee $O/l_hbm.txt

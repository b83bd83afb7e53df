#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== graph test"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "graphed" > $O/g2_pytest.log 2>&1; echo "rc=$?"; tail -15 $O/g2_pytest.log | cut -c1-300
echo "== latency"; timeout 600 python tools/bench_latency.py 300 > $O/g2_latency.txt 2>&1; echo "rc=$?"; cat $O/g2_latency.txt | tail -8

#!/bin/bash
# dense-layer operand prefetch: parity (full suite), front-end timing, default bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/s_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/s_pytest.log | cut -c1-250
echo "== front end"; timeout 300 python tools/profile_frontend.py > $O/s_frontend.txt 2>&1; echo "rc=$?"; tail -3 $O/s_frontend.txt
echo "== bench default"; python bench.py --no-cpu-baseline > $O/s_bench.json 2> $O/s_bench.err; rc=$?; echo "rc=$rc"; python -c "
import json;d=json.loads([l for l in open('$O/s_bench.json') if l.startswith('{')][-1]);print(round(d['value'],1),'e2e',round(d['e2e']['value'],1),round(d['roofline']['achieved'],1),d['roofline']['kernel_share_of_step'],d['clocks'])"

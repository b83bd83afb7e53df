#!/bin/bash
# chained modulator kernel: parity first (with a hard timeout), then timing
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== modulator tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "modulator or stay_inside" > $O/ch_pytest1.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 $O/ch_pytest1.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/ch_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/ch_pytest.log | cut -c1-250
echo "== front end"; timeout 300 python tools/profile_frontend.py > $O/ch_frontend.txt 2>&1; echo "rc=$?"; tail -3 $O/ch_frontend.txt
echo "== bench"; python bench.py --no-cpu-baseline --no-burst > $O/ch_bench.json 2> $O/ch_bench.err; echo "rc=$?"; python -c "
import json;d=json.loads([l for l in open('$O/ch_bench.json') if l.startswith('{')][-1]);print(round(d['value'],1),'e2e',round(d['e2e']['value'],1),round(d['roofline']['achieved'],1),round(d['roofline']['kernel_share_of_step'],4),d['gpu_launches'])"

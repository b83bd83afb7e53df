#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pipeline tests"; timeout 900 python -m pytest tests -x -q -m gpu -k "pipeline or host or full_size" > $O/x_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/x_pytest.log | cut -c1-250
for i in 1 2; do
python bench.py --no-cpu-baseline --no-burst > $O/x_bench_$i.json 2> $O/x_bench_$i.err; python -c "
import json;d=json.loads([l for l in open('$O/x_bench_$i.json') if l.startswith('{')][-1]);print('chunk',d['config']['chunk_slices'],round(d['value'],1),'e2e',round(d['e2e']['value'],1),round(d['roofline']['achieved'],1),round(d['roofline']['kernel_share_of_step'],4),d['clocks']['sm_mhz'])"
done

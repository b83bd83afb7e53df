#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for c in 118 59 0; do
python bench.py --no-cpu-baseline --no-burst --chunk $c > $O/w_bench_$c.json 2> $O/w_bench_$c.err; python -c "
import json;d=json.loads([l for l in open('$O/w_bench_$c.json') if l.startswith('{')][-1]);print('chunk',d['config']['chunk_slices'],round(d['value'],1),'e2e',round(d['e2e']['value'],1),round(d['roofline']['achieved'],1),round(d['roofline']['kernel_share_of_step'],4),d['clocks']['sm_mhz'])"
done

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python bench.py --no-cpu-baseline > $O/b2_bench.json 2> $O/b2_bench.err; echo "rc=$?"; python -c "
import json;d=json.loads([l for l in open('$O/b2_bench.json') if l.startswith('{')][-1]);print(round(d['value'],1),'e2e',round(d['e2e']['value'],1),round(d['roofline']['achieved'],1),d['roofline']['dense_modulations_sustained'],d['roofline']['single_launch_after_idle_tflops'])"
tail -3 $O/b2_bench.err

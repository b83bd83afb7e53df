#!/bin/bash
# final 8-GPU / 4-GPU check (one box): N=8 once, N=4 once, driver-style arguments
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
for n in 8 4; do
  echo "== bench N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29880 + n)) bench.py --gpus $n --steps 3 --warmup 3 > $O/y_bench_n$n.json 2> $O/y_bench_n$n.err
  echo "rc=$?"; python -c "import json;d=json.loads([l for l in open('$O/y_bench_n$n.json') if l.startswith('{')][-1]);print(d['n_gpus'],round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['e2e']['result'],d['config']['chunk_slices'],round(d['roofline']['achieved'],1),d['roofline']['kernel_share_of_step'])"; grep -E "\[bench\]|Error|error" $O/y_bench_n$n.err | head -5
done

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
for rep in 1 2; do
  echo "== bench N=8 (rep $rep)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29880 + rep)) bench.py --gpus 8 --steps 10 --warmup 3 > $O/d_bench_n8_$rep.json 2> $O/d_bench_n8_$rep.err
  echo "rc=$?"; python -c "import json;d=json.loads([l for l in open('$O/d_bench_n8_$rep.json') if l.startswith('{')][-1]);print(d['n_gpus'],round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['e2e']['result'],d['config']['chunk_slices'],round(d['roofline']['achieved'],1),d['roofline']['kernel_share_of_step'])"; grep -E "\[bench\]|Error|error" $O/d_bench_n8_$rep.err | head -5
done
echo "== reference arm N=8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29890 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200

#!/bin/bash
# 2-GPU sanity after the reassembly rewrite (the band kernel stores into rank 0's buffer over peer memory)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/o_multi.log 2>&1; echo "rc=$?"; tail -3 $O/o_multi.log | cut -c1-250
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus 2 --steps 3 --warmup 3 > $O/o_bench_n2.json 2> $O/o_bench_n2.err
echo "rc=$?"; python -c "import json;d=json.loads([l for l in open('$O/o_bench_n2.json') if l.startswith('{')][-1]);print(d['n_gpus'],round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['config']['exchange'],round(d['roofline']['achieved'],1))"

#!/bin/bash
# multi-GPU session (gpurun --gpus 4 or 8): the regression tests of the sweep + full-size bench lines
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
echo "== multi-GPU tests"; timeout 1500 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_peer_gather_two_gpus" -m gpu -q -s > $O/d_multi.log 2>&1; echo "rc=$?"; tail -25 $O/d_multi.log | cut -c1-300
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    for rep in 1 2; do
      echo "== bench N=$n (rep $rep)"
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n * 10 + rep)) bench.py --gpus $n --steps 5 --warmup 3 > $O/d_bench_n${n}_$rep.json 2> $O/d_bench_n${n}_$rep.err
      echo "rc=$?"; python -c "import json;d=json.loads([l for l in open('$O/d_bench_n${n}_$rep.json') if l.startswith('{')][-1]);print(d['n_gpus'],round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['e2e']['result'],d['config']['chunk_slices'],d['config']['exchange'][:40])"; grep -E "\[bench\]|Error|error" $O/d_bench_n${n}_$rep.err | head -5
    done
  fi
done

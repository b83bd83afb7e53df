#!/bin/bash
# One gpurun call: correctness first, then the measurements of round 2.  Every step has its own timeout and log.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/a_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/a_pytest.log 2>&1; echo "rc=$?"; tail -15 $O/a_pytest.log
echo "== remaining tests after first failure (if any)"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > $O/a_pytest_all.log 2>&1; echo "rc=$?"; tail -30 $O/a_pytest_all.log | cut -c1-220
echo "== bench sine"; timeout 300 python bench.py --steps 3 --warmup 3 > $O/a_bench_sine.json 2> $O/a_bench_sine.err; echo "rc=$?"; tail -c 900 $O/a_bench_sine.json; tail -2 $O/a_bench_sine.err
echo "== bench morlet"; timeout 300 python bench.py --steps 3 --warmup 3 --activation morlet --no-cpu-baseline > $O/a_bench_morlet.json 2> $O/a_bench_morlet.err; echo "rc=$?"; tail -c 600 $O/a_bench_morlet.json; tail -2 $O/a_bench_morlet.err
echo "== bench cfg4 (L=9, Z=128)"; timeout 300 python bench.py --steps 3 --warmup 3 --num-layers 9 --latent-dim 128 --no-cpu-baseline > $O/a_bench_cfg4.json 2> $O/a_bench_cfg4.err; echo "rc=$?"; tail -c 600 $O/a_bench_cfg4.json; tail -2 $O/a_bench_cfg4.err
echo "== bench dense mods"; timeout 300 python bench.py --steps 3 --warmup 3 --mods dense --no-cpu-baseline > $O/a_bench_dense.json 2> $O/a_bench_dense.err; echo "rc=$?"; tail -c 600 $O/a_bench_dense.json; tail -2 $O/a_bench_dense.err
echo "== bench fp16x3"; timeout 300 python bench.py --steps 2 --warmup 3 --precision fp16x3 --mods dense --no-cpu-baseline --slices 2068 > $O/a_bench_x3.json 2> $O/a_bench_x3.err; echo "rc=$?"; tail -c 600 $O/a_bench_x3.json; tail -2 $O/a_bench_x3.err
echo "== precision table"; timeout 300 python tools/precision_table.py > $O/a_precision_table.txt 2>&1; echo "rc=$?"; cat $O/a_precision_table.txt
echo "== power split"
for v in default NO_MMA NO_SIN NO_BIAS NO_STS; do
  if [ $v = default ]; then timeout 120 python tools/power_split.py 60 >> $O/a_power.txt 2>&1; else MRINR_LIB=build/libmrinr_$v.so timeout 120 python tools/power_split.py 60 >> $O/a_power.txt 2>&1; fi
done
cat $O/a_power.txt

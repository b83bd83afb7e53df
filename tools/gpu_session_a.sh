#!/bin/bash
# One gpurun call: correctness first, then the measurements of round 2.  Every step has its own timeout and log.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/a_smi.txt 2>&1
# separate processes per group: a trapped kernel kills the CUDA context of its process only
echo "== g1 parity (round-1 scope + Morlet envelope)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not fp16x3 and not auto and not hard_cases" > $O/a_g1.log 2>&1; echo "rc=$?"; tail -12 $O/a_g1.log | cut -c1-250
echo "== g2 fp16x3 / auto / hard cases"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fp16x3 or auto or hard_cases" > $O/a_g2.log 2>&1; echo "rc=$?"; grep -E "max-abs|auto ->|passed|failed|Error|error" $O/a_g2.log | tail -40 | cut -c1-250
echo "== diag train"; timeout 300 python tools/diag_train.py > $O/a_diag_train.txt 2>&1; echo "rc=$?"; cat $O/a_diag_train.txt | cut -c1-200
echo "== diag train, SGEMM weight gradient (diagnostic build)"; MRINR_LIB=build/libmrinr_wgsgemm.so timeout 300 python tools/diag_train.py > $O/a_diag_train_sgemm.txt 2>&1; echo "rc=$?"; grep -E "===|forward|<<<<|Error|error" $O/a_diag_train_sgemm.txt | cut -c1-200
echo "== g3 training"; timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -s > $O/a_g3.log 2>&1; echo "rc=$?"; tail -40 $O/a_g3.log | cut -c1-250
echo "== g4 replay of test_mod_siren.py"; timeout 600 python -m pytest tests/test_replay_reference_script.py -m gpu -q > $O/a_g4.log 2>&1; echo "rc=$?"; tail -12 $O/a_g4.log | cut -c1-250
echo "== bench sine"; timeout 300 python bench.py --steps 3 --warmup 3 > $O/a_bench_sine.json 2> $O/a_bench_sine.err; echo "rc=$?"; tail -c 900 $O/a_bench_sine.json; tail -2 $O/a_bench_sine.err
for k in 70 66 62; do echo "== bench sine, front end overlapped, $k CTA pairs"; timeout 300 python bench.py --steps 3 --warmup 3 --overlap-clusters $k --no-cpu-baseline --no-burst > $O/a_bench_ov$k.json 2> $O/a_bench_ov$k.err; echo "rc=$?"; python -c "import json;d=json.load(open('$O/a_bench_ov$k.json'));print(d['value'],d['roofline']['achieved'],d['roofline']['kernel_share_of_step'],d['clocks'])"; tail -2 $O/a_bench_ov$k.err; done
echo "== bench morlet"; timeout 300 python bench.py --steps 3 --warmup 3 --activation morlet --no-cpu-baseline > $O/a_bench_morlet.json 2> $O/a_bench_morlet.err; echo "rc=$?"; tail -c 600 $O/a_bench_morlet.json; tail -2 $O/a_bench_morlet.err
echo "== bench cfg4 (L=9, Z=128)"; timeout 300 python bench.py --steps 3 --warmup 3 --num-layers 9 --latent-dim 128 --no-cpu-baseline > $O/a_bench_cfg4.json 2> $O/a_bench_cfg4.err; echo "rc=$?"; tail -c 600 $O/a_bench_cfg4.json; tail -2 $O/a_bench_cfg4.err
echo "== bench dense mods"; timeout 300 python bench.py --steps 3 --warmup 3 --mods dense --no-cpu-baseline > $O/a_bench_dense.json 2> $O/a_bench_dense.err; echo "rc=$?"; tail -c 600 $O/a_bench_dense.json; tail -2 $O/a_bench_dense.err
echo "== bench fp16x3"; timeout 300 python bench.py --steps 2 --warmup 3 --precision fp16x3 --mods dense --no-cpu-baseline --slices 2068 > $O/a_bench_x3.json 2> $O/a_bench_x3.err; echo "rc=$?"; tail -c 600 $O/a_bench_x3.json; tail -2 $O/a_bench_x3.err
echo "== precision table"; timeout 300 python tools/precision_table.py > $O/a_precision_table.txt 2>&1; echo "rc=$?"; cat $O/a_precision_table.txt
echo "== SM sweep"; timeout 200 python tools/sm_sweep.py > $O/a_sm_sweep.txt 2>&1; echo "rc=$?"; cat $O/a_sm_sweep.txt
echo "== power split"
for v in default NO_MMA NO_SIN NO_BIAS NO_STS; do
  if [ $v = default ]; then timeout 120 python tools/power_split.py 60 >> $O/a_power.txt 2>&1; else MRINR_LIB=build/libmrinr_$v.so timeout 120 python tools/power_split.py 60 >> $O/a_power.txt 2>&1; fi
done
cat $O/a_power.txt

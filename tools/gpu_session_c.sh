#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== g1"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not fp16x3 and not auto and not hard_cases" > $O/c_g1.log 2>&1; echo "rc=$?"; tail -5 $O/c_g1.log | cut -c1-250
echo "== g2 fp16x3 / auto / hard cases"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fp16x3 or auto or hard_cases" > $O/c_g2.log 2>&1; echo "rc=$?"; grep -E "max-abs|auto ->|passed|failed|Error|error" $O/c_g2.log | tail -40 | cut -c1-250
echo "== diag train (tensor-core weight gradient, scaled gradients)"; timeout 300 python tools/diag_train.py > $O/c_diag_train.txt 2>&1; echo "rc=$?"; cat $O/c_diag_train.txt | cut -c1-200
echo "== diag train, SGEMM weight gradient (diagnostic build)"; MRINR_LIB=build/libmrinr_wgsgemm.so timeout 300 python tools/diag_train.py > $O/c_diag_train_sgemm.txt 2>&1; echo "rc=$?"; grep -E "===|forward|<<<<|rror" $O/c_diag_train_sgemm.txt | cut -c1-200
echo "== g3 training"; timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -s > $O/c_g3.log 2>&1; echo "rc=$?"; tail -30 $O/c_g3.log | cut -c1-250
echo "== g4 replay"; timeout 600 python -m pytest tests/test_replay_reference_script.py -m gpu -q > $O/c_g4.log 2>&1; echo "rc=$?"; tail -3 $O/c_g4.log
echo "== morlet variants"
timeout 120 python tools/morlet_bench.py > $O/c_morlet.txt 2>&1
for m in 0x0 0x1 0x5 0x7; do MRINR_LIB=build/libmrinr_morlet_$m.so timeout 120 python tools/morlet_bench.py >> $O/c_morlet.txt 2>&1; done
timeout 120 python tools/morlet_bench.py >> $O/c_morlet.txt 2>&1
cat $O/c_morlet.txt
echo "== bench fp16x3"; timeout 300 python bench.py --steps 2 --warmup 3 --precision fp16x3 --mods dense --no-cpu-baseline --slices 2068 > $O/c_bench_x3.json 2> $O/c_bench_x3.err; echo "rc=$?"; tail -c 700 $O/c_bench_x3.json; tail -2 $O/c_bench_x3.err
echo "== bench auto on dense mods"; timeout 300 python bench.py --steps 2 --warmup 3 --precision auto --mods dense --no-cpu-baseline --slices 2068 > $O/c_bench_auto.json 2> $O/c_bench_auto.err; echo "rc=$?"; python -c "import json;d=json.load(open('$O/c_bench_auto.json'));print(d['value'],d['dtype'],d['config']['precision'])"; tail -2 $O/c_bench_auto.err
echo "== precision table"; timeout 300 python tools/precision_table.py > $O/c_precision_table.txt 2>&1; echo "rc=$?"; cat $O/c_precision_table.txt
echo "== training step timing"; timeout 200 python tools/train_bench.py > $O/c_train_bench.txt 2>&1; echo "rc=$?"; cat $O/c_train_bench.txt
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()"; echo "rc=$?"

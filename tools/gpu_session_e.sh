#!/bin/bash
# single-GPU session: full GPU suite, Morlet variants, training timing, remaining bench lines, then the ncu captures
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== full GPU suite (as the driver runs it)"; timeout 1200 python -m pytest tests -x -q -m gpu > $O/e_pytest.log 2>&1; echo "rc=$?"; tail -6 $O/e_pytest.log | cut -c1-250
echo "== morlet variants"
rm -f $O/e_morlet.txt
timeout 120 python tools/morlet_bench.py >> $O/e_morlet.txt 2>&1
for m in 0x0 0xA 0xF 0x5_g1 0x5_g2 0xF_g1 0xF_g2 0x5_w16 0xF_w16; do MRINR_LIB=build/libmrinr_morlet_$m.so timeout 120 python tools/morlet_bench.py >> $O/e_morlet.txt 2>&1; done
timeout 120 python tools/morlet_bench.py >> $O/e_morlet.txt 2>&1
cat $O/e_morlet.txt
echo "== training step timing"; timeout 200 python tools/train_bench.py > $O/e_train_bench.txt 2>&1; echo "rc=$?"; cat $O/e_train_bench.txt
echo "== bench morlet"; timeout 300 python bench.py --steps 3 --warmup 3 --activation morlet --no-cpu-baseline > $O/e_bench_morlet.json 2> $O/e_bench_morlet.err; echo "rc=$?"; python -c "import json;d=json.load(open('$O/e_bench_morlet.json'));r=d['roofline'];print(d['value'],r['achieved'],r['frac'],d['clocks'])"
echo "== bench fp16x3"; timeout 300 python bench.py --steps 2 --warmup 3 --precision fp16x3 --mods dense --no-cpu-baseline --slices 2068 > $O/e_bench_x3.json 2> $O/e_bench_x3.err; echo "rc=$?"; python -c "import json;d=json.load(open('$O/e_bench_x3.json'));r=d['roofline'];print(d['value'],r['achieved'],r['frac'],d['clocks'])"
echo "== bench sine (default line)"; timeout 300 python bench.py --steps 3 --warmup 3 > $O/e_bench_sine.json 2> $O/e_bench_sine.err; echo "rc=$?"; python -c "import json;d=json.load(open('$O/e_bench_sine.json'));r=d['roofline'];print(d['value'],d['e2e']['value'],r['achieved'],r['frac'],d['clocks'],d['cpu_baseline'])"
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/e_bench_ref.json 2> $O/e_bench_ref.err; echo "rc=$?"; tail -c 400 $O/e_bench_ref.json
bash tools/gpu_session_b.sh

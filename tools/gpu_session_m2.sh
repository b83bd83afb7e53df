#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python tools/morlet_bench.py 2>&1 | tail -1 | sed 's/default (mask 0xF: all FMA)/default (shipped: degree 4, mask 0x5)/' | tee $O/m2_morlet.txt
for v in 0x5 0x7 0xD; do MRINR_LIB=$PWD/build/libmrinr_d3_$v.so python tools/morlet_bench.py 2>&1 | tail -1 | tee -a $O/m2_morlet.txt; done
python tools/morlet_bench.py 2>&1 | tail -1 | sed 's/default (mask 0xF: all FMA)/default (shipped: degree 4, mask 0x5)/' | tee -a $O/m2_morlet.txt

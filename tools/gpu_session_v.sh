#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for c in 1 4; do echo "== dense cluster $c"; MRINR_LIB=$PWD/build/libmrinr_dc$c.so timeout 300 python tools/profile_frontend.py 2>&1 | tail -2; done
echo "== default (2)"; timeout 300 python tools/profile_frontend.py 2>&1 | tail -2

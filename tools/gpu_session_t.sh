#!/bin/bash
# ncu --set full of one modulator-layer launch (dense_split_kernel<256>, K = 512) and one 8x8-conv launch (<64>, K = 2048)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_split_kernel -s 11 -c 1 -f -o $O/t_dense256 python tools/profile_frontend.py 94000 2 > $O/t_ncu256.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_split_kernel -s 0 -c 1 -f -o $O/t_dense64 python tools/profile_frontend.py 94000 2 > $O/t_ncu64.log 2>&1; echo "ncu rc=$?"
ls -la $O/t_dense*.ncu-rep

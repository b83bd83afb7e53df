#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_conv_tc_kernel -s 3 -c 1 -f -o $O/c2_conv python tools/profile_frontend.py 94000 2 > $O/c2_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O/c2_conv.ncu-rep

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows320 -s 2 -c 1 -f -o $O/m_rows320 python tools/bench_hbm.py 1056 > $O/m_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O/m_rows320.ncu-rep

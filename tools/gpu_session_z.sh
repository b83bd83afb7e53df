#!/bin/bash
# training iteration: time + ncu launch list (current build)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python tools/train_bench.py 400 > $O/z_train.txt 2>&1; echo "rc=$?"; tail -1 $O/z_train.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file $O/z_train_launches.csv python tools/train_bench.py 400 > $O/z_ncu.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== full GPU suite"; timeout 1200 python -m pytest tests -x -q -m gpu > $O/f_pytest.log 2>&1; echo "rc=$?"; tail -4 $O/f_pytest.log | cut -c1-250
echo "== plain train bench"; python tools/train_bench.py 400 > $O/f_train_plain.txt 2>&1; rc=$?; echo "rc=$rc"; cat $O/f_train_plain.txt
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file $O/f_train_launches.csv python tools/train_bench.py 400 > $O/f_train_ncu.log 2>&1; echo "ncu rc=$?"
fi

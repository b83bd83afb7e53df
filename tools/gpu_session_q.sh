#!/bin/bash
# verification after the metric-kernel rewrite: full suite, smoke, default bench, refreshed ncu launch list of the bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/q_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/q_pytest.log | cut -c1-250
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" > $O/q_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/q_smoke.log
echo "== bench default"; python bench.py > $O/q_bench.json 2> $O/q_bench.err; rc=$?; echo "rc=$rc"; tail -1 $O/q_bench.json | cut -c1-400
if [ $rc -eq 0 ]; then
  echo "== ncu launch list (bench --steps 2 --warmup 1 on 2068 slices)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/q_launches.csv \
     python bench.py --steps 2 --warmup 1 --slices 2068 --no-cpu-baseline > $O/q_ncu.log 2>&1; echo "ncu rc=$?"
fi

#!/bin/bash
# final single-GPU verification: full suite, smoke, default bench (+ Morlet line)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/fin_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/fin_pytest.log | cut -c1-250
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" > $O/fin_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/fin_smoke.log
echo "== bench default"; python bench.py > $O/fin_bench.json 2> $O/fin_bench.err; echo "rc=$?"; tail -1 $O/fin_bench.json | cut -c1-1500
echo "== bench reference arm"; python bench.py --impl reference > $O/fin_bench_ref.json 2> $O/fin_bench_ref.err; echo "rc=$?"; tail -1 $O/fin_bench_ref.json | cut -c1-600

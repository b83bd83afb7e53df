#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== fft tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fft or kspace or full_size" > $O/f2_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/f2_pytest.log | cut -c1-250
echo "== bench_hbm"; timeout 600 python tools/bench_hbm.py 1056 > $O/f2_hbm.txt 2>&1; echo "rc=$?"; tail -3 $O/f2_hbm.txt
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none \
   -k regex:"rows320|cols320" --csv --log-file $O/f2_ncu.csv python tools/bench_hbm.py 1056 > $O/f2_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_metrics_median.py $O/f2_ncu.csv | tee $O/f2_ncu.txt

#!/bin/bash
# full GPU suite + smoke + ncu DRAM counters of the HBM kernels (current kernels)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/i_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/i_pytest.log | cut -c1-250
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" > $O/i_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/i_smoke.log
echo "== plain hbm bench"; python tools/bench_hbm.py 1056 > $O/i_hbm.txt 2>&1; rc=$?; echo "rc=$rc"; cat $O/i_hbm.txt
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
     -k regex:"image_to_patches|patches_to_image|minmax" --csv --log-file $O/i_hbm_ncu.csv python tools/bench_hbm.py 1056 > $O/i_ncu_hbm.log 2>&1; echo "ncu rc=$?"
  python tools/ncu_metrics_median.py $O/i_hbm_ncu.csv | tee $O/i_hbm_ncu.txt
fi

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none \
   -k regex:"reduce_kernel|ssim|rows|cols" --csv --log-file $O/k_ncu.csv python tools/bench_hbm.py 1056 > $O/k_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_metrics_median.py $O/k_ncu.csv | tee $O/k_ncu.txt

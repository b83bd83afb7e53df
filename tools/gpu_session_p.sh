#!/bin/bash
# SSIM / reduction-pass rewrite: parity tests, HBM bench, ncu counters of the two metric kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== metrics tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "metrics or ssim or psnr" > $O/p_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/p_pytest.log | cut -c1-250
echo "== bench_hbm"; timeout 600 python tools/bench_hbm.py 1056 > $O/p_hbm.txt 2>&1; echo "rc=$?"; tail -3 $O/p_hbm.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none \
   -k regex:"reduce_kernel|ssim" --csv --log-file $O/p_ncu.csv python tools/bench_hbm.py 1056 > $O/p_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_metrics_median.py $O/p_ncu.csv | tee $O/p_ncu.txt

#!/bin/bash
# ncu session (one gpurun call; every profiled command runs plain first, same arguments, and must exit 0).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --slices 940 --no-cpu-baseline --no-burst"
echo "== plain bench (short)"; $BENCH > $O/b_plain.json 2> $O/b_plain.err; rc=$?; echo "rc=$rc"; tail -c 300 $O/b_plain.json
if [ $rc -eq 0 ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/b_launches.csv $BENCH > $O/b_ncu_launch.log 2>&1; echo "rc=$?"
  echo "== ncu --set full, synthesis kernel (sine)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_tc5 -s 4 -c 1 -o $O/b_siren_sine $BENCH > $O/b_ncu_sine.log 2>&1; echo "rc=$?"
fi
MB="python bench.py --steps 2 --warmup 3 --slices 940 --no-cpu-baseline --no-burst --activation morlet"
echo "== plain bench morlet (short)"; $MB > $O/b_plain_m.json 2> $O/b_plain_m.err; rc=$?; echo "rc=$rc"
if [ $rc -eq 0 ]; then
  echo "== ncu --set full, synthesis kernel (morlet)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_tc5 -s 4 -c 1 -o $O/b_siren_morlet $MB > $O/b_ncu_morlet.log 2>&1; echo "rc=$?"
fi
echo "== plain hbm bench"; python tools/bench_hbm.py 1056 > $O/b_hbm.txt 2>&1; rc=$?; echo "rc=$rc"; cat $O/b_hbm.txt
if [ $rc -eq 0 ]; then
  echo "== ncu dram bytes of the HBM kernels"
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
     -k regex:"image_to_patches|patches_to_image|minmax|complex_abs|rows_kernel|cols_kernel|ssim|metrics" --csv --log-file $O/b_hbm_ncu.csv python tools/bench_hbm.py 1056 > $O/b_ncu_hbm.log 2>&1; echo "rc=$?"
fi
ls -la $O/b_* | head -30

#!/bin/bash
# refreshed ncu launch list of the bench (final build: chained modulator)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/q2_launches.csv \
   python bench.py --steps 2 --warmup 1 --slices 2068 --no-cpu-baseline --no-burst > $O/q2_ncu.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# 16 epilogue warps (make -C mri_inr_b200/csrc w16) against the shipped 8-warp kernel: Morlet and sine sustained rates
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/g_w16.txt
for lib in "" build/libmrinr_w16a.so build/libmrinr_w16b.so; do
  echo "== lib=${lib:-shipped}" | tee -a $O/g_w16.txt
  MRINR_LIB=$lib timeout 300 python tools/morlet_bench.py >> $O/g_w16.txt 2>&1; echo "morlet rc=$?"
  MRINR_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 3 --slices 2350 --e2e-steps 1 --no-cpu-baseline --no-burst > $O/g_bench_$(basename "${lib:-shipped}" .so).json 2>> $O/g_w16.txt; echo "bench rc=$?"
  python - <<PY >> $O/g_w16.txt
import json
try:
    l = json.loads(open("$O/g_bench_$(basename "${lib:-shipped}" .so).json").read().strip().splitlines()[-1])
    print("sine:", l["value"], l["unit"], "roofline", l["roofline"]["achieved"], l["roofline"]["frac"])
except Exception as e:
    print("bench line unreadable:", e)
PY
done
cat $O/g_w16.txt

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== dense-related tests"; timeout 900 python -m pytest tests -x -q -m gpu -k "modulator or encoder or train or forward or smoke" > $O/u_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/u_pytest.log | cut -c1-250
echo "== front end"; timeout 300 python tools/profile_frontend.py > $O/u_frontend.txt 2>&1; echo "rc=$?"; tail -3 $O/u_frontend.txt
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
   -k regex:"dense_split|encoder_conv" --csv --log-file $O/u_ncu.csv python tools/profile_frontend.py 94000 2 > $O/u_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_metrics_median.py $O/u_ncu.csv | tee $O/u_ncu.txt

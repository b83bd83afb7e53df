"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`:
python tools/ncu_summary.py raw.csv src.csv [top_n]"""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__sass_inst_executed_op_shared_ld.sum', 'sm__sass_inst_executed_op_shared_st.sum',
        'smsp__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_tensor.sum', 'smsp__cycles_active.avg']
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]:16s} {d[k][1]}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['# Samples']].isdigit()]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {s: sum(int(r[idx[s]]) for r in data if r[idx[s]].isdigit()) for s in stalls}
total = sum(int(r[idx['# Samples']]) for r in data)
print("total samples", total)
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {s:25s} {v:9d} {100 * v / total:5.1f}%")
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:topn]:
    st = {s: int(r[idx[s]]) for s in stalls if r[idx[s]].isdigit() and int(r[idx[s]]) > 0}
    main = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(r[idx['Address']][-5:], r[idx['# Samples']].rjust(8), f"{100*int(r[idx['# Samples']])/total:5.1f}%", r[idx['Source']][:58].ljust(58), main)

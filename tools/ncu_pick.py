"""stdin: `ncu --page raw --csv`; prints the metrics that matter for a tensor-core / HBM kernel, one column per launch."""
import csv
import sys

KEEP = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_uniform',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__pcsamp_warps_issue_stalled', 'smsp__average_warp_latency_issue_stalled', 'smsp__warp_issue_stalled',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_st.sum',
        'smsp__average_warps_issue_stalled']
rows = list(csv.reader(sys.stdin))
if not rows:
    sys.exit('no rows')
hdr = rows[0]
units = rows[1] if len(rows) > 1 else []
name_col = hdr.index('Kernel Name') if 'Kernel Name' in hdr else 4
for r in rows[2:]:
    print('==', r[name_col][:100])
for j, h in enumerate(hdr):
    if any(h.startswith(k) for k in KEEP):
        vals = [r[j] for r in rows[2:]]
        print('%-95s %-10s %s' % (h, units[j] if j < len(units) else '', '  '.join(vals)))

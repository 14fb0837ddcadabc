#!/usr/bin/env python
"""Print the key metrics of an `ncu --page raw --csv` dump, one column per profiled launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum',
 'smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
idx = {h: i for i, h in enumerate(hdr)}
print("kernels:", " | ".join(r[idx['Kernel Name']][20:75] for r in data))
for w in want:
    if w in idx:
        vals = [r[idx[w]] for r in data]
        try:
            if all(float(v.replace(',', '')) == 0 for v in vals): continue
        except ValueError: pass
        print("%-84s %-9s %s" % (w.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio','')[:84], units[idx[w]][:9], " | ".join("%12s" % v[:12] for v in vals)))

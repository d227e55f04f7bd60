"""Print a fixed set of metrics per launch from an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
want += [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    for i in idx:
        v = r[i]
        if 'stalled' in hdr[i]:
            try:
                if float(v) < 3: continue
            except ValueError: pass
        print(f"{hdr[i]:80s} {units[i]:14s} {v[:90]}")
    print()

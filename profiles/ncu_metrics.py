import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','launch__occupancy_limit_blocks','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_local_op_st.sum','sm__cycles_active.avg','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor']
for i,h in enumerate(hdr):
    if h in want: print(h, units[i], vals[i])
out=[]
for i,h in enumerate(hdr):
    if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try: out.append((float(vals[i]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except: pass
print(' | '.join('%s %.2f'%(h,v) for v,h in sorted(out,reverse=True)[:10]))

import csv,subprocess,sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','gpu__time_duration.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','launch__block_size','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__registers_per_thread','lts__t_sector_hit_rate.pct','lts__t_sectors.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.per_cycle_active','smsp__average_warp_latency_per_inst_issued.ratio','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__sass_average_branch_targets_threads_uniform.pct','smsp__thread_inst_executed_per_inst_executed.ratio']
print("| metric | unit | value |\n|---|---|---|")
for i,h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0)>0.3):
        print(f"| {h} | {units[i]} | {vals[i]} |")

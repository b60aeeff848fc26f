import csv,sys,subprocess,io
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
r=list(csv.reader(io.StringIO(out)))
h=r[0]; u=r[1]; v=r[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_active.avg','sm__cycles_active.max','sm__cycles_active.min','sm__cycles_elapsed.avg','sm__cycles_elapsed.avg.per_second','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for i,k in enumerate(h):
    if k in keys or ('issue_stalled' in k and 'per_issue_active' in k):
        print(f"{k:95s} {u[i]:12s} {v[i]}")

import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg']
want+= [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
idx=[(w,hdr.index(w)) for w in want if w in hdr]
for r in rows[2:]:
    for w,i in idx:
        v=r[i]
        if 'issue_stalled' in w:
            try:
                if float(v)<0.15: continue
            except: pass
            w=w.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio','')
        print(f"{w[:70]:70s} {v[:90]} {units[i]}")
    print('---')

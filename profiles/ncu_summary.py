import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr: print(w, r[hdr.index(w)][:90], rows[1][hdr.index(w)])
    st=[(float(r[i] or 0),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for i,h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('.ratio')]
    print('stalls/issue:', ', '.join(f"{n} {v:.2f}" for v,n in sorted(st,reverse=True)[:6]))

import csv,re,sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','sm__cycles_active.avg','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__t_sectors.avg.pct_of_peak_sustained_elapsed']
for w in want:
    idx=[i for i,c in enumerate(h) if c==w]
    if idx: print(w, rows[1][idx[0]], [r[idx[0]] for r in rows[2:]])
for i,c in enumerate(h):
    if re.search(r'smsp__average_warps_issue_stalled_.*_per_issue_active',c):
        v=float(rows[2][i])
        if v>0.05: print('%-90s %.3f'%(c, v))

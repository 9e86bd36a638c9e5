import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]; u=rows[1]; v=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','launch__grid_size','launch__waves_per_multiprocessor','smsp__thread_inst_executed_per_inst_executed.ratio']
for w in want:
    if w in h: i=h.index(w); print('  %-70s %s %s'%(w,v[i],u[i]))
out=[]
for i,name in enumerate(h):
    if 'issue_stalled' in name and name.endswith('per_issue_active.ratio'):
        try: out.append((float(v[i]),name.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except: pass
print('  stalls:', ', '.join('%s=%.2f'%(n,x) for x,n in sorted(out,reverse=True)[:6]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=rows[1]; si=h.index("Warp Stall Sampling (All Samples)"); s_=h.index("Source"); ie=h.index("Instructions Executed")
tot=0; out=[]; inst=0
for idx,r in enumerate(rows[2:]):
    try: x=float(r[si]); inst+=float(r[ie])
    except: continue
    tot+=x; out.append((x,idx,r[s_].strip()[:80]))
print('  total warp-inst', inst)
for x,idx,s in sorted(out,reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 14]: print('  %6.2f%% sass#%4d %s'%(100*x/tot,idx,s))

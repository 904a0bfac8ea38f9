import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct','sm__warps_active.avg.pct','launch__registers_per_thread','launch__block_size','launch__grid_size','smsp__average_warps_issue_stalled','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warp_latency_per_inst_issued']
for h,u,v in zip(hdr,units,vals):
    if any(h.startswith(w) for w in want) and 'pct_of_peak' not in h.replace('smsp__issue_active.avg.pct_of_peak','X').replace('sm__warps_active.avg.pct_of_peak','X') and 'per_second' not in h:
        try:
            if float(v)==0: continue
        except: pass
        print(h,u,v)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ia=hdr.index('Address'); isrc=hdr.index('Source'); ins=hdr.index('Instructions Executed'); smp=hdr.index('# Samples'); thr=hdr.index('Avg. Threads Executed')
tot=sum(int(r[ins]) for r in data); tots=sum(int(r[smp]) for r in data)
print('total inst',tot,'samples',tots)
base=int(data[0][ia],16)
acc=0; accs=0; start=0
step=int(sys.argv[2]) if len(sys.argv)>2 else 40
for i,r in enumerate(data):
    acc+=int(r[ins]); accs+=int(r[smp])
    if (i+1)%step==0 or i==len(data)-1:
        if acc*200>tot or accs*200>tots:
            print('%04x-%04x inst %5.1f%% samp %5.1f%%  thr %s  %s'%(int(data[start][ia],16)-base,int(r[ia],16)-base,100*acc/tot,100*accs/tots,r[thr],r[isrc].strip()[:40]))
        acc=0;accs=0;start=i+1

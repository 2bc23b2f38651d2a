import csv,sys,subprocess,collections
rep=sys.argv[1]; nframes=int(sys.argv[2]) if len(sys.argv)>2 else 32768
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__cycles_active.avg','l1tex__throughput.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','launch__grid_size','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    print('--- kernel', r[hdr.index('Kernel Name')][:50])
    for k in keys:
        if k in hdr: print(f'{k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}')
    ie=float(r[hdr.index('smsp__inst_executed.sum')]); print('warp-instr per frame', ie/nframes)
sass=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(sass.splitlines()))
hdr_idx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
KI=int(sys.argv[3]) if len(sys.argv)>3 else 0; start=hdr_idx[KI]; end=hdr_idx[KI+1]-1 if len(hdr_idx)>KI+1 else len(rows)
hdr=rows[start]; ci=hdr.index('Instructions Executed'); si=hdr.index('Source'); smp=hdr.index('# Samples')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body=[r for r in rows[start+1:end] if len(r)>smp and r[smp].isdigit()]
tot=sum(int(r[ci]) for r in body); ts=sum(int(r[smp]) for r in body)
byop=collections.Counter(); agg=collections.Counter()
for r in body:
    t=r[si].split(); op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    byop[op]+=int(r[ci])
    for c in stall_cols:
        if r[c].isdigit(): agg[hdr[c]]+=int(r[c])
print('sass lines',len(body))
print(' '.join(f'{op}:{n/nframes:.0f}' for op,n in byop.most_common(32)))
print(' '.join(f'{k[6:]}:{100*v/ts:.1f}%' for k,v in agg.most_common(10)))
top=sorted(body,key=lambda r:-int(r[smp]))[:14]
for r in top:
    st={hdr[c]:int(r[c]) for c in stall_cols if r[c].isdigit() and int(r[c])>0}
    main=sorted(st.items(), key=lambda kv:-kv[1])[:2]
    print(f'{100*int(r[smp])/ts:5.2f}% exec {int(r[ci])/nframes:6.1f}/fr  {r[si][:60]:60s} {main}')

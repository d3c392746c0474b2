import csv, collections, re, sys
raw, src = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(raw)))
hdr=rows[0]; units=rows[1]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i=hdr.index(k); print(k, '=', r[i], units[i])
    # stall reasons
    for i,h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
            try:
                v=float(r[i])
                if v>0.15: print('  stall', h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), round(v,2))
            except: pass
    print('---')
rows=list(csv.reader(open(src)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address']
start=hi[0]; end=hi[1]-1 if len(hi)>1 else len(rows)
hdr=rows[start]
ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
ops=collections.Counter(); samp=collections.Counter(); tot=0
for r in rows[start+1:end]:
    if len(r)<=ie: continue
    try: ex=int(r[ie])
    except: continue
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia])
    op=m.group(2).split('.')[0] if m else r[ia][:10]
    if op=='F2F' or op=='I2F' or op=='F2I': op=m.group(2)
    ops[op]+=ex; tot+=ex
    try: samp[op]+=int(r[isamp])
    except: pass
print('executed',tot)
for op,c in ops.most_common(28):
    print(f'{op:14s} {c:12d} {100*c/tot:5.1f}%  samples {samp[op]}')

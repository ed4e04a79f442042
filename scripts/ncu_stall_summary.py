"""Instruction mix and warp-stall samples per SASS opcode from `ncu -i X.ncu-rep --page source --csv` (runs without a GPU).

    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv && python scripts/ncu_stall_summary.py /tmp/src.csv
"""
import csv,sys,re,collections
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]
ix={k:i for i,k in enumerate(h)}
ops=collections.Counter(); samp=collections.Counter(); stall=collections.Counter()
stall_cols=[k for k in h if k.startswith('stall_')]
tot_inst=0; tot_s=0
for r in rows[2:]:
    if len(r)<len(h): continue
    src=r[ix['Source']].strip()
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src)
    op=m.group(2) if m else src[:20]
    op='.'.join(op.split('.')[:2]) if op.startswith(('MUFU','F2FP','LDS','STS','LDG','STG','LDTM','UTC','SYNCS')) else op.split('.')[0]
    n=int(r[ix['Instructions Executed']] or 0); s=int(r[ix['# Samples']] or 0)
    ops[op]+=n; samp[op]+=s; tot_inst+=n; tot_s+=s
    for k in stall_cols:
        v=r[ix[k]]
        if v: stall[k]+=int(v)
print('total inst',tot_inst,'samples',tot_s)
for op,n in ops.most_common(28): print(f'{op:14s} inst {n:10d} {100*n/tot_inst:5.1f}%  samples {samp[op]:7d} {100*samp[op]/max(tot_s,1):5.1f}%')
print({k:v for k,v in stall.most_common(8)})

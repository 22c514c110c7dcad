"""group the ncu source page into straight-line regions (equal execution count) and print their stall mix"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
cols=['stall_long_sb','stall_wait','stall_math','stall_short_sb','stall_selected','stall_not_selected','stall_barrier','stall_dispatch','stall_lg','stall_mio','stall_branch_resolving','stall_no_inst']
regions=[];cur=None
for r in data:
    e=r[ix['Instructions Executed']]
    if not e.isdigit(): continue
    e=int(e)
    if cur is None or cur['exec']!=e:
        cur={'exec':e,'n':0,'start':r[ix['Address']][-5:], **{c:0 for c in cols}}
        regions.append(cur)
    cur['n']+=1
    for c in cols:
        v=r[ix[c]]
        if v.isdigit(): cur[c]+=int(v)
tot=sum(sum(g[c] for c in cols) for g in regions)
print("total samples",tot)
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.012
for g in regions:
    s=sum(g[c] for c in cols)
    if s>thr*tot:
        print(g['start'], 'exec',g['exec'],'ninstr',g['n'],'samples',s, '%.1f%%'%(100*s/tot), {c[6:]:g[c] for c in cols if g[c]>0.08*s})

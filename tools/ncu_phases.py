"""Per-phase (between BAR.SYNC) instruction and stall-sample breakdown from an ncu source-page CSV."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
seen=set()
for ki,s in enumerate(starts):
    e=starts[ki+1] if ki+1<len(starts) else len(rows)
    name=rows[s][1][:60]
    if name in seen: continue
    seen.add(name)
    hdr=rows[s+1]; ix={h:i for i,h in enumerate(hdr)}
    stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    phase=0; ph=collections.defaultdict(lambda: collections.Counter())
    for r in rows[s+2:e]:
        if len(r)<6: continue
        ins=r[ix['Source']].strip()
        try: n=int(r[ix['Instructions Executed']])
        except: continue
        op=(ins.split()[0] if not ins.startswith('@') else ins.split()[1]).split('.')[0]
        ph[phase]['inst']+=n
        ph[phase]['samples']+=int(r[ix['Warp Stall Sampling (All Samples)']] or 0)
        for c in stall_cols:
            try: ph[phase][c]+=int(r[ix[c]])
            except: pass
        if op in ('LDS','STS'):
            try:
                ph[phase]['smem_wf']+=int(r[ix['L1 Wavefronts Shared']]); ph[phase]['smem_ideal']+=int(r[ix['L1 Wavefronts Shared Ideal']])
            except: pass
        if op=='BAR': phase+=1
    print('=====',name)
    tot=sum(v['samples'] for v in ph.values())
    for k in sorted(ph):
        v=ph[k]
        top=sorted(((v[c],c) for c in stall_cols),reverse=True)[:4]
        print(f" phase {k}: inst {v['inst']:9d} samples {v['samples']:6d} ({100*v['samples']/max(tot,1):4.1f}%) smem wf/ideal {v['smem_wf']}/{v['smem_ideal']}  top: "+', '.join(f"{c[6:]}={n}" for n,c in top))

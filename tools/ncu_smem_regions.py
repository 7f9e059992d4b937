import csv, collections, subprocess, sys
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[k for k,r in enumerate(rows[:5]) if 'Source' in r][0]
hdr=rows[hi]
i_src=hdr.index('Source'); i_ex=hdr.index('Instructions Executed'); i_wf=hdr.index('L1 Wavefronts Shared'); i_id=hdr.index('L1 Wavefronts Shared Ideal'); i_ss=hdr.index('Warp Stall Sampling (All Samples)')
agg=collections.defaultdict(lambda:[0,0,0,0]); bars=[]; tot=0
for k,r in enumerate(rows[hi+1:]):
    src=r[i_src].strip()
    try: ex=int(r[i_ex] or 0); wf=int(r[i_wf] or 0); idl=int(r[i_id] or 0); ss=int(r[i_ss] or 0)
    except Exception: continue
    t=src.split()
    if not t: continue
    op=t[1] if t[0].startswith('@') and len(t)>1 else t[0]
    if 'BAR' in src: bars.append(k)
    tot+=wf
    key=op.split('.')[0]+('.128' if '.128' in op else ('.64' if '.64' in op else ''))
    reg=sum(1 for b in bars if b<=k)
    a=agg[(reg,key)]; a[0]+=ex; a[1]+=wf; a[2]+=idl; a[3]+=ss
print('total shared wavefronts',tot,'barriers at',bars)
regs=collections.defaultdict(lambda:[0,0,0])
for (reg,key),a in sorted(agg.items()):
    regs[reg][0]+=a[0]; regs[reg][1]+=a[1]; regs[reg][2]+=a[3]
    if a[1]>0 or key in('DFMA','DMUL','DADD','RED','BAR','SYNCS'):
        print(reg,key,'exec',a[0],'wavefronts',a[1],'ideal',a[2],'samples',a[3])
for r,a in sorted(regs.items()): print('region',r,'exec',a[0],'wavefronts',a[1],'samples',a[2])

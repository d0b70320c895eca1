#!/usr/bin/env python
"""Group the per-call profile (tools/profile_calls.py output) of the conv entry points by kernel class."""
import re, collections, sys
rows=[]
for l in open(sys.argv[1]):
    m=re.match(r'\s*([\d.]+) ms\s+(\d+) x\s+([\d.]+) TF/s\s+([\d.]+) GB/s\s+(\S+)\s+(.*)',l)
    if m: rows.append((float(m.group(1)),int(m.group(2)),float(m.group(3)),m.group(5),m.group(6).strip()))
def cls(name,tag):
    if name=='conv2d_wgrad':
        m=re.match(r'M(\d+) taps(\d+) Ca(\d+) Cb(\d+) s(\d)',tag); M,t,ca,cb,s=map(int,m.groups())
        if t==9 and ca<=64 and s==1: return 'wgrad halo'
        if t==1: return 'wgrad 1x1 Ca%s'%('128' if ca==128 else 'other')
        return 'wgrad other (taps %d)'%t
    if name!='conv2d_fwd': return name
    m=re.match(r'M(\d+) K(\d+)x(\d+)x(\d+) N(\d+) s(\d)( pro)?( tc)?',tag); M,kh,kw,c,n,s=map(int,m.groups()[:6]); tc=m.group(8)
    if not tc: return 'fwd non-tc'
    if kh==3 and c%32==0 and n<=128 and s==1: return 'fwd halo N%d'%(32 if n<=32 else 64 if n<=64 else 128)
    if kh==1: return 'fwd 1x1 N%s'%('<=64' if n<=64 else '128' if n<=128 else '>128')
    return 'fwd other tc (k%d)'%kh
agg=collections.defaultdict(lambda:[0.0,0,0.0])
for ms,n,tf,name,tag in rows:
    a=agg[cls(name,tag)]; a[0]+=ms; a[1]+=n; a[2]+=tf*ms
for k,v in sorted(agg.items(),key=lambda kv:-kv[1][0]): print("%8.2f ms %4d calls %6.1f TF/s  %s"%(v[0],v[1],v[2]/v[0] if v[0] else 0,k))
print("total %.2f ms"%sum(r[0] for r in rows))

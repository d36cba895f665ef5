"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by enclosing function."""
import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
srcpath=sys.argv[2]
cur_file=None; agg={}; samples={}
hdr=None
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=='Function Name': continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<8: continue
    if r[0]!='':
        key=(cur_file,int(r[0]),r[1].strip()[:80])
        try: inst=int(r[7]); samp=int(r[6])
        except: continue
        agg[key]=agg.get(key,0)+inst; samples[key]=samples.get(key,0)+samp
tot=sum(agg.values()); tots=sum(samples.values())
src=open(srcpath).read().split('\n')
def func_of(line):
    for l in range(line,0,-1):
        t=src[l-1]
        m2=re.search(r'auto (\w+)=\[&\]',t)
        if m2: return 'lambda:'+m2.group(1)
        m=re.match(r'\s*(wb_[a-z0-9_]+)\(',t)
        if m and l>=2 and ('__global__' in src[l-2]): return m.group(1)
        m=re.search(r'__device__.*?(wb_[a-z0-9_]+)\(',t)
        if m: return m.group(1)
    return '?'
byf=collections.Counter(); bys=collections.Counter()
for k,v in agg.items():
    f = func_of(k[1]) if k[0]==srcpath.split('/')[-1] else k[0]
    byf[f]+=v; bys[f]+=samples[k]
print('total instructions %d, samples %d'%(tot,tots))
for f,v in byf.most_common(25):
    print('%5.1f%% inst  %5.1f%% samples  %s'%(100*v/tot,100*bys[f]/tots,f))

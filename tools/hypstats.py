import sys, numpy as np
sys.path.insert(0,'/root/repo')
from wolkenbase_b200 import api, synth
for n in (10_000_000, 100_000_000):
    cloud=synth.generate(2,n,seed=2)
    ctx=api.Context(0); ctx.set_params(); ctx.reserve(cloud.n)
    ctx.add_cloud(cloud); ctx.build(); ctx.scan()
    t0=ctx.tiles()
    ctx.postscan()
    t=ctx.tiles()
    g=ctx.geometry()
    h=t['hyperboloidSize']
    print(n, 'spacing',g.spacing,'tiles',len(t),'tree frac',t['treeFlags'].mean(),'npts mean',t['nPoints'].mean())
    print(' hyp before pct', np.percentile(t0['hyperboloidSize'],[1,10,50,90,99,100]).round(3))
    print(' hyp after  pct', np.percentile(h,[1,10,50,90,99,100]).round(3))
    print(' height pct', np.percentile(t['height'],[10,50,90,99]).round(3), 'density pct', np.percentile(t['density'],[10,50,90]).round(2))
    ctx.close()

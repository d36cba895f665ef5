import sys, numpy as np
sys.path.insert(0,'/root/repo')
from wolkenbase_b200 import api, synth
from oracle import wb_oracle as O
scene,n=int(sys.argv[1]),int(sys.argv[2])
cloud=synth.generate(scene,n,seed=scene)
res=O.run([O.file_from_cloud(cloud)],classify=False)
ctx=api.Context(0); ctx.set_params(); ctx.add_cloud(cloud); ctx.build(); ctx.scan()
t0=ctx.tiles(); ctx.postscan(); t=ctx.tiles()
rs=res.tiles_scan
for f in ("density","hyperboloidSize","height"):
    d=np.abs(t0[f].view(np.int64)-rs[f].view(np.int64))
    bad=np.nonzero(d>0)[0]
    print(f,'scan-stage diffs',len(bad),'max ulp',d.max() if len(d) else 0)
    for i in bad[:5]:
        print('   tile',t0['n'][i],'npts',t0['nPoints'][i],'gpu',repr(float(t0[f][i])),'ref',repr(float(rs[f][i])),'tree',t0['treeFlags'][i],rs['treeFlags'][i])
print('npoints max', t['nPoints'].max(), 'mean', t['nPoints'].mean())

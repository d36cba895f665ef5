"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): all kernels, ragged sizes."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
from wolkenbase_b200 import api, synth

for scene, n, take in [(2, 6000, 5003), (5, 4000, 3999), (3, 3000, 2977)]:
    cloud = synth.generate(scene, n, seed=scene)
    recs = np.ascontiguousarray(cloud.records[:take])
    take = recs.shape[0]
    ctx = api.Context(0)
    ctx.set_params()
    ctx.add_extent(cloud.min_corner, cloud.max_corner)
    ctx.add_las(recs, cloud.fmt, cloud.scale, cloud.offset)
    ctx.run()
    lab = ctx.labels(take)
    print(scene, take, np.bincount(lab, minlength=3)[:3], len(ctx.leaves()), len(ctx.tiles()), ctx.count_classes()[:3])
    ctx.close()

# round 2: the sharded pipeline (halo kernels, tile grid, Hilbert-ordered second store) with 3 ranks as threads on
# this GPU, identical locations across ranks and dropped records included; the census kernels
from wolkenbase_b200 import multigpu  # noqa: E402

d = synth.describe(2, 9000)
cuts = [d.grid_nx * k // 3 for k in range(4)]
strips, base = [], 0
for k in range(3):
    c = synth.generate(2, 9000, seed=5, region=(cuts[k], 0, cuts[k + 1] - cuts[k], d.grid_ny), gps_base=base)
    base += c.n
    strips.append(c)
strips[1].records[7, :12] = strips[0].records[11, :12]
strips[2].records[3::9, 14] &= 0xf8
labs, sst, st = multigpu.run_threads([[c] for c in strips], multigpu.PARAMS)
print("sharded", [int(np.bincount(l, minlength=3)[2]) for l in labs], [int(s["n_halo_classify"]) for s in sst])
cloud = synth.generate(2, 5000, seed=9)
ctx = api.Context(0)
ctx.keep_records()
ctx.set_params()
ctx.add_cloud(cloud)
ctx.run()
print("census", ctx.census(cap=8))
ctx.close()

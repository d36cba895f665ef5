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

"""Full parity at a few million points (the oracle needs ~15 s per scene on the GPU box's cores):
byte-identical dump, bit-identical hyperboloidSize, identical labels.  WB_LARGE=1 adds a 10 M run."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import wb_oracle as O  # noqa: E402
from wolkenbase_b200 import api, synth  # noqa: E402


SIZES = [(2, 3_000_000), (4, 2_000_000)] + ([(1, 10_000_000)] if os.environ.get("WB_LARGE") else [])


@pytest.mark.parametrize("scene,n", SIZES)
def test_full_parity_millions(scene, n):
    cloud = synth.generate(scene, n, seed=scene)
    ctx = api.Context(0)
    ctx.set_params()
    ctx.add_cloud(cloud)
    ctx.run()
    lab = ctx.labels(cloud.n)
    dump = ctx.dump()
    tiles = ctx.tiles()
    st = ctx.stats()
    ctx.close()
    res = O.run([O.file_from_cloud(cloud)])
    assert dump == res.dump
    assert len(tiles) == len(res.tiles)
    assert (tiles["nPoints"] == res.tiles["nPoints"]).all() and (tiles["treeFlags"] == res.tiles["treeFlags"]).all()
    ulp = np.abs(tiles["hyperboloidSize"].view(np.int64) - res.tiles["hyperboloidSize"].view(np.int64))
    mism = int((lab != res.labels).sum())
    print("scene %d: %d points, %d leaves, %d tiles (max %d points), hyperboloidSize ulp diffs %d (max %d), "
          "label mismatches %d, margin points gpu %d / oracle %d"
          % (scene, cloud.n, st["n_leaves"], len(tiles), tiles["nPoints"].max(), int((ulp > 0).sum()), int(ulp.max()),
             mism, st["n_margin"], res.margin_count))
    assert ulp.max() <= 4
    assert mism <= st["n_margin"] + res.margin_count

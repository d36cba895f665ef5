"""The sharded path (halo exchange, tile-table merge, two-stage build) reproduces the labels of the
single-GPU run exactly.  Ranks are simulated inside one process on one GPU (multigpu.run_local);
the same Rank code runs under torch.distributed in bench.py --gpus N."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from wolkenbase_b200 import api, multigpu, synth  # noqa: E402


def _single(clouds, params):
    ctx = api.Context(0)
    ctx.set_params(**params)
    for c in clouds:
        ctx.add_extent(c.min_corner, c.max_corner)
    for c in clouds:
        ctx.add_las(c.records, c.fmt, c.scale, c.offset)
    ctx.run()
    lab = ctx.labels(sum(c.n for c in clouds))
    tiles = ctx.tiles()
    ctx.close()
    return lab, tiles


@pytest.mark.parametrize("world,scene,n,dups", [(2, 2, 60000, 0), (3, 2, 90000, 0), (4, 5, 60000, 0),
                                                (2, 2, 40000, 2000)])
def test_sharded_equals_single(world, scene, n, dups):
    d = synth.describe(scene, n)
    cuts = [d.grid_nx * k // world for k in range(world + 1)]
    clouds, base = [], 0
    for k in range(world):
        c = synth.generate(scene, n, seed=31, region=(cuts[k], 0, cuts[k + 1] - cuts[k], d.grid_ny), gps_base=base)
        base += c.n
        if dups:                         # identical locations inside a strip (and so inside its halos)
            c = synth.with_duplicates(c, dups, k)
        clouds.append(c)
    params = dict(tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1)
    want, tiles = _single(clouds, params)
    # under the emulated library (tests/test_emulated_library.py) "device" memory is host memory
    dev = torch.device("cpu") if os.environ.get("WB_EMULATED") else torch.device("cuda", 0)
    ranks = [multigpu.Rank(k, world, api.Context(0), api.Context(0), params, dev) for k in range(world)]
    got = multigpu.run_local(ranks, clouds)
    # the merged, post-scanned tile table equals the single-GPU one
    t = ranks[0].a
    T = ranks[0].geom.snake_hi - ranks[0].geom.snake_lo + 1
    hyp = ranks[0].t_hyp.cpu().numpy().view(np.float64)
    npnt = ranks[0].t_np.cpu().numpy()
    idx = tiles["n"] - ranks[0].geom.snake_lo
    assert int((npnt != 0).sum()) == len(tiles)
    assert (npnt[idx] == tiles["nPoints"]).all()
    assert (hyp[idx] == tiles["hyperboloidSize"]).all()
    off = 0
    for k in range(world):
        assert (got[k] == want[off:off + clouds[k].n]).all(), "rank %d" % k
        off += clouds[k].n
    halo = sum(r.n_cls - r.n_own for r in ranks)
    assert halo > 0
    for r in ranks:
        r.a.close()
        r.b.close()

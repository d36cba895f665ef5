"""wb_shard_run over real NCCL — the transport of bench.py --gpus N and wolkencli --gpus N — against the ORACLE's
labels of the whole cloud.  Needs at least two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`, outcome
recorded in profiles/).  Two launch shapes: one process per rank (the ncclUniqueId handed over as bytes, no
torch.distributed anywhere), and one host thread per rank inside one process."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS = dict(tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1)


def _gpus():
    import torch
    return torch.cuda.device_count()


def _strips(scene, n, world, seed=31):
    from wolkenbase_b200 import synth
    d = synth.describe(scene, n)
    cuts = [d.grid_nx * k // world for k in range(world + 1)]
    clouds, base = [], 0
    for k in range(world):
        c = synth.generate(scene, n, seed=seed, region=(cuts[k], 0, cuts[k + 1] - cuts[k], d.grid_ny), gps_base=base)
        base += c.n
        clouds.append(c)
    return clouds


def _rank_process(rank, world, uid, scene, n, ret):
    sys.path.insert(0, ROOT)
    from wolkenbase_b200 import api, multigpu
    cloud = _strips(scene, n, world)[rank]
    ctx = api.Context(rank)
    comm = api.Comm.nccl(ctx, uid, rank, world)
    for _ in range(2):                                  # context and communicator reused from step to step
        m = multigpu.load_rank(ctx, [cloud], PARAMS)
        ctx.shard_run(comm)
    st = ctx.shard_stats()
    ret[rank] = {"labels": ctx.shard_labels(m).copy(), "halo": int(st["n_halo_classify"]),
                 "margin": int(ctx.stats()["n_margin"]), "por_max": float(st["por_max"])}
    comm.close()
    ctx.close()


def _check(labels, margins, clouds):
    from oracle import wb_oracle as O
    ref = O.run([O.file_from_cloud(c) for c in clouds], **PARAMS)
    got = np.concatenate(labels)
    mism = int((got != ref.labels).sum())
    assert mism <= sum(margins) + int(ref.margin_count), "%d labels differ from the oracle's" % mism
    return ref


@pytest.mark.parametrize("scene,n", [(2, 120000), (3, 100000)])
def test_one_process_per_rank_over_nccl_equals_oracle(scene, n):
    world = min(_gpus(), 4)
    if world < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from wolkenbase_b200 import api
    uid = api.Comm.unique_id()
    ret = mp.Manager().dict()
    mp.spawn(_rank_process, args=(world, uid, scene, n, ret), nprocs=world, join=True)
    ref = _check([ret[r]["labels"] for r in range(world)], [ret[r]["margin"] for r in range(world)],
                 _strips(scene, n, world))
    assert all(ret[r]["halo"] > 0 for r in range(world))
    assert all(ret[r]["por_max"] == float(ref.tiles["hyperboloidSize"].max()) for r in range(world))


def test_one_thread_per_rank_over_nccl_equals_oracle():
    world = min(_gpus(), 4)
    if world < 2:
        pytest.skip("needs two GPUs")
    from wolkenbase_b200 import multigpu
    clouds = _strips(5, 90000, world)
    labs, sst, st = multigpu.run_threads([[c] for c in clouds], PARAMS, devices=list(range(world)), transport="nccl")
    _check(labs, [int(s["n_margin"]) for s in st], clouds)
    assert all(s["n_halo_classify"] > 0 for s in sst)

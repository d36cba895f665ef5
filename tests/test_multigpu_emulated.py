"""The N>1 path as bench.py --gpus N runs it — one process per rank under torch.distributed, wb_shard_run inside the
library — with world_size 2 on the CPU: the ranks' contexts come from the emulated library (tests/simt: the library's
own source built for the host) and gloo carries the bytes through the CUSTOM transport (NCCL's place on the GPU box).
Labels must equal the oracle's for the whole cloud."""
import os
import socket
import subprocess
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")
EMULATED = os.path.join(SIMT, "libwolken_b200_emulated.so")
PARAMS = dict(tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1)
SCENE, N, SEED = 2, 30000, 31


def _clouds(world):
    from wolkenbase_b200 import synth
    d = synth.describe(SCENE, N)
    cuts = [d.grid_nx * k // world for k in range(world + 1)]
    clouds, base = [], 0
    for k in range(world):
        c = synth.generate(SCENE, N, seed=SEED, region=(cuts[k], 0, cuts[k + 1] - cuts[k], d.grid_ny), gps_base=base)
        base += c.n
        clouds.append(c)
    return clouds


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WB_LIB"] = EMULATED
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from wolkenbase_b200 import api, multigpu
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cloud = _clouds(world)[rank]
    ctx = api.Context(0)
    comm = multigpu.gloo_comm(ctx, dist, rank, world)
    for _ in range(2):                                  # twice: the context and the communicator are reused per step
        n = multigpu.load_rank(ctx, [cloud], PARAMS)
        ctx.shard_run(comm)
    labels = ctx.shard_labels(n)
    st = ctx.shard_stats()
    ret[rank] = {"labels": labels.copy(), "halo": int(st["n_halo_classify"]), "sent": int(st["bytes_sent"])}
    comm.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_run_world2_gloo_on_the_emulated_library():
    subprocess.check_call(["make", "-s", "-C", SIMT, "libwolken_b200_emulated.so"])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from oracle import wb_oracle as O
    clouds = _clouds(2)
    want = O.run([O.file_from_cloud(c) for c in clouds], **PARAMS).labels
    got = np.concatenate([ret[0]["labels"], ret[1]["labels"]])
    assert (got == want).all(), int((got != want).sum())
    assert ret[0]["halo"] > 0 and ret[1]["halo"] > 0
    assert ret[0]["sent"] == 16 * (ret[1]["halo"] + 0) or ret[0]["sent"] > 0

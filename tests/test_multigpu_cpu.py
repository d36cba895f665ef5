"""Host-side logic of the multi-GPU path on CPU: ownership intervals, halo reach, and the
all_to_all exchange under a world_size-2 gloo group."""
import math
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wolkenbase_b200 import multigpu as M


def test_ownership_bounds():
    b = M.ownership_bounds([(0.0, 9.9), (10.1, 19.9), (20.1, 30.0)])
    assert b[0] == (-math.inf, 10.0) and b[1] == (10.0, 20.0) and b[2] == (20.0, math.inf)
    assert M.ownership_bounds([(3.0, 4.0)]) == [(-math.inf, math.inf)]


def test_reach_radius_is_an_upper_bound():
    # brute force: for random P above Q, Hyperboloid::in (shape.cpp:127-135) implies dist <= reach
    rng = np.random.default_rng(0)
    zmax, t, s = 50.0, 0.05, 0.8
    for por_max in (0.3, 5.0, 200.0):
        zq = torch.tensor(rng.uniform(0, 50, 2000))
        r = M.reach_radius(zq, zmax, t, por_max, s).numpy()
        for _ in range(20):
            pz = rng.uniform(0, zmax, 2000)
            por = rng.uniform(0.01, por_max, 2000)
            d = rng.uniform(0, 400, 2000)
            zd = (pz - t + por) - zq.numpy()
            inside = (zd > 0) & (zd * zd - (d * s) ** 2 >= por * por)
            assert (d[inside] <= r[inside]).all()


def test_coords_two_roundings():
    xi = torch.tensor([0, 1, 123456789, -5], dtype=torch.int32)
    got = M.coords(xi, 0.001, 500000.0).numpy()
    want = np.array([500000.0 + 0.001 * float(v) for v in xi.tolist()])
    assert (got == want).all()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = M.TorchComm(dist)
    rng = np.random.default_rng(100 + rank)
    n = 5000 + 37 * rank
    # rank r owns x in [100 r, 100 r + 99.5]
    xi = torch.tensor(rng.integers(100000 * rank, 100000 * rank + 99500, n), dtype=torch.int32)
    yi = torch.tensor(rng.integers(0, 100000, n), dtype=torch.int32)
    zi = torch.tensor(rng.integers(0, 20000, n), dtype=torch.int32)
    cl = torch.tensor(rng.integers(0, 32, n), dtype=torch.uint8)
    ext = comm.all_gather_doubles([float(xi.min()) * 0.001, 0, 0, float(xi.max()) * 0.001, 100, 20], "cpu")
    strips = [(e[0], e[3]) for e in ext]
    x = M.coords(xi, 0.001, 0.0)
    z = M.coords(zi, 0.001, 0.0)
    radius = M.reach_radius(z, 20.0, 0.0, 2.0, 1.0)
    rows = M.pack((xi, yi, zi, cl))
    sends = [rows[:0] if k == rank else rows[M.select_for_strip(x, radius, strips[k])] for k in range(world)]
    recv = comm.all_to_all_rows(sends)
    tot = [torch.tensor([float(s.shape[0]) for s in sends])]
    comm.all_reduce_sum(tot)
    ret[rank] = {"sent": [s.numpy().copy() for s in sends], "recv": [r.numpy().copy() for r in recv],
                 "strips": strips, "mx": comm.all_reduce_max_scalar(float(rank), "cpu")}
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_exchange_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["mx"] == 1.0 and r1["mx"] == 1.0
    assert r0["strips"] == r1["strips"]
    # what rank 0 sent to rank 1 is what rank 1 received from rank 0, in order, and vice versa
    assert (r0["sent"][1] == r1["recv"][0]).all() and (r1["sent"][0] == r0["recv"][1]).all()
    assert r0["recv"][0].shape[0] == 0 and r1["recv"][1].shape[0] == 0
    # only points near the other strip travel: rank 0's strip ends at ~99.5, rank 1's starts at 100
    assert 0 < r0["sent"][1].shape[0] < 5000
    x_sent = r0["sent"][1][:, 0] * 0.001
    assert x_sent.min() > 100.0 - 0.5 - math.sqrt(20 * 20 + 2 * 2.0 * 20) - 1e-3

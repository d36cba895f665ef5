"""Host side of the multi-GPU path on the CPU: the gloo stand-ins for the three collectives wb_shard_run needs
(wb_comm_ops: all_gather, all_to_all_v, all_reduce_max_u8), world_size 2, on plain host buffers."""
import ctypes as C
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from wolkenbase_b200 import multigpu as M
    dist.init_process_group("gloo", rank=rank, world_size=world)
    all_gather, all_to_all_v, all_reduce_max_u8 = M.gloo_ops(dist, rank, world)
    u64 = C.POINTER(C.c_uint64)
    # all_gather: 5 bytes per rank
    send = np.arange(5, dtype=np.uint8) + 10 * rank
    recv = np.zeros(5 * world, dtype=np.uint8)
    assert all_gather(None, send.ctypes.data, recv.ctypes.data, 5) == 0
    # all_to_all_v: rank r sends (3 + r + k) bytes of value 16 r + k to rank k, from shuffled offsets
    cnt = np.array([3 + rank + k for k in range(world)], dtype=np.uint64)
    cnt[rank] = 0
    off = np.zeros(world, dtype=np.uint64)
    buf = np.zeros(64, dtype=np.uint8)
    pos = 40
    for k in range(world):
        pos -= int(cnt[k])
        off[k] = pos
        buf[pos:pos + int(cnt[k])] = 16 * rank + k
    rcnt = np.array([0 if k == rank else 3 + k + rank for k in range(world)], dtype=np.uint64)
    roff = np.array([7 * k for k in range(world)], dtype=np.uint64)
    rbuf = np.full(64, 255, dtype=np.uint8)
    assert all_to_all_v(None, buf.ctypes.data, off.ctypes.data_as(u64), cnt.ctypes.data_as(u64), rbuf.ctypes.data,
                        roff.ctypes.data_as(u64), rcnt.ctypes.data_as(u64)) == 0
    grid = np.zeros(1000, dtype=np.uint8)
    grid[rank::7] = 1 + rank
    assert all_reduce_max_u8(None, grid.ctypes.data, len(grid)) == 0
    ret[rank] = {"gather": recv.copy(), "rbuf": rbuf.copy(), "grid": grid.copy()}
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_collectives_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    for r in range(2):
        assert ret[r]["gather"].tolist() == [0, 1, 2, 3, 4, 10, 11, 12, 13, 14]
        other = 1 - r
        n = 3 + other + r
        want = np.full(64, 255, dtype=np.uint8)
        want[7 * other:7 * other + n] = 16 * other + r
        assert (ret[r]["rbuf"] == want).all()
        g = np.zeros(1000, dtype=np.uint8)
        g[0::7] = 1
        g[1::7] = 2
        assert (ret[r]["grid"] == g).all()

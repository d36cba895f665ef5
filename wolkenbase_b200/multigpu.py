"""One cloud over the GPUs of a box: launch harness around wb_shard_run.

All of the sharded pipeline — halo selection, exchange, tile grid, the two builds, classify — lives in the
library (csrc/wb_shard.cuh, NCCL called from C++).  This module only
  * hands the NCCL unique id of rank 0 to the other ranks of a torchrun job (`nccl_comm`),
  * runs several ranks as threads of one process over the LOCAL transport (`run_threads`: parity tests on one
    GPU, and on the CPU emulator of tests/simt),
  * offers gloo as a CUSTOM transport for the emulated library (`gloo_comm`: the world_size-2 CPU tests),
  * and holds the N>1 body of bench.py (`bench`).
The reference has no distributed mode (one std::thread pool, threads.cpp:91-113).
"""
import ctypes as C
import json
import os
import threading
import time

import numpy as np

from . import api

PARAMS = dict(tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1)


def load_rank(ctx, clouds, params=None, dev_records=None, window=None):
    """The rank's own files into its context: header corners + records (host arrays, or device pointers).
    window = (x_lo, x_hi): the rank is handed WHOLE files and keeps the records with x in that interval
    (wb_set_window); its header boxes are clipped to it."""
    ctx.clear()
    ctx.set_params(**dict(PARAMS, **(params or {})))
    if window is not None:
        ctx.set_window(*window)
    for c in clouds:
        mn, mx = list(c.min_corner), list(c.max_corner)
        if window is not None:
            mn[0], mx[0] = max(mn[0], window[0]), min(mx[0], window[1])
        ctx.add_extent(mn, mx)
    for i, c in enumerate(clouds):
        if dev_records is not None:
            ctx.add_las_device(dev_records[i], c.n, c.fmt, c.rec_len, c.scale, c.offset)
        else:
            ctx.add_las(c.records, c.fmt, c.scale, c.offset)
    return ctx.num_loaded() if window is not None else sum(c.n for c in clouds)


# ---------------------------------------------------------------------------- ranks as threads (LOCAL transport)

def run_threads(clouds_per_rank, params=None, devices=None, transport="local", windows=None):
    """clouds_per_rank[r] = list of synth.Cloud-like files of rank r (ascending x).  Every rank runs in its own host
    thread on devices[r] (default: all on device 0).  transport "local": plain copies + a barrier; "nccl": one NCCL
    communicator per thread, which needs a distinct device per rank (what wolkencli --gpus N does).
    windows[r] = (x_lo, x_hi): every rank is given the same whole files and keeps its x-interval.
    Returns (labels per rank, shard stats per rank, wb stats)."""
    W = len(clouds_per_rank)
    devices = devices or [0] * W
    group = api.LocalGroup(W) if transport == "local" else None
    uid = api.Comm.unique_id() if transport == "nccl" else None
    out = [None] * W
    err = [None] * W

    def work(r):
        try:
            ctx = api.Context(devices[r])
            comm = api.Comm.local(ctx, group, r) if group is not None else api.Comm.nccl(ctx, uid, r, W)
            n = load_rank(ctx, clouds_per_rank[r], params, window=windows[r] if windows else None)
            ctx.shard_run(comm)
            out[r] = (ctx.shard_labels(n), ctx.shard_stats(), ctx.stats())
            comm.close()
            ctx.close()
        except BaseException as e:      # a rank that dies would leave the others in the barrier: report and stop
            err[r] = e
            os._exit(97) if W > 1 and not isinstance(e, api.WolkenError) else None

    th = [threading.Thread(target=work, args=(r,)) for r in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if group is not None:
        group.close()
    for e in err:
        if e is not None:
            raise e
    return [o[0] for o in out], [o[1] for o in out], [o[2] for o in out]


# ---------------------------------------------------------------------------- transports under torch.distributed

def nccl_comm(ctx, dist, rank, world):
    """wb_comm over NCCL: rank 0's ncclUniqueId reaches the others through the job's process group."""
    box = [api.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return api.Comm.nccl(ctx, box[0], rank, world)


def _bytes_at(ptr, n):
    return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr)) if n else np.empty(0, dtype=np.uint8)


def gloo_ops(dist, rank, world):
    """The three collectives of wb_comm_ops over gloo, for buffers in HOST memory (the emulated library's 'device'
    pointers are host pointers).  Pointers arrive as integers; each returns 0 on success."""
    import torch

    def all_gather(user, d_send, d_recv, nbytes):
        try:
            send = torch.from_numpy(_bytes_at(d_send, nbytes).copy())
            outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, send)
            _bytes_at(d_recv, nbytes * world)[:] = torch.cat(outs).numpy()
            return 0
        except Exception:
            return 1

    def all_to_all_v(user, d_send, s_off, s_cnt, d_recv, r_off, r_cnt):
        try:
            sc = [int(s_cnt[k]) for k in range(world)]
            rc = [int(r_cnt[k]) for k in range(world)]
            parts = [_bytes_at(d_send + int(s_off[k]), sc[k]) for k in range(world)]
            send = torch.from_numpy(np.concatenate(parts) if sum(sc) else np.empty(0, dtype=np.uint8))
            recv = torch.empty(sum(rc), dtype=torch.uint8)
            dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc)
            got, pos = recv.numpy(), 0
            for k in range(world):
                _bytes_at(d_recv + int(r_off[k]), rc[k])[:] = got[pos:pos + rc[k]]
                pos += rc[k]
            return 0
        except Exception:
            return 1

    def all_reduce_max_u8(user, d_buf, n):
        try:
            t = torch.from_numpy(_bytes_at(d_buf, n))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return 0
        except Exception:
            return 1

    return all_gather, all_to_all_v, all_reduce_max_u8


def gloo_comm(ctx, dist, rank, world):
    """CUSTOM transport for the EMULATED library: gloo moves the bytes (world_size-2 CPU tests)."""
    return api.Comm.custom(ctx, *gloo_ops(dist, rank, world), rank, world)


# ---------------------------------------------------------------------------- parity on the ranks of a real job

PARITY_SCENE, PARITY_POINTS, PARITY_STRIPS, PARITY_SEED = 3, 400_000, 8, 17


def parity_strips():
    """The 8 files of the committed sharded fixture (tests/golden/make_sharded.py)."""
    from . import synth
    d = synth.describe(PARITY_SCENE, PARITY_POINTS)
    cuts = [d.grid_nx * k // PARITY_STRIPS for k in range(PARITY_STRIPS + 1)]
    clouds, base = [], 0
    for k in range(PARITY_STRIPS):
        c = synth.generate(PARITY_SCENE, PARITY_POINTS, seed=PARITY_SEED,
                           region=(cuts[k], 0, cuts[k + 1] - cuts[k], d.grid_ny), gps_base=base)
        base += c.n
        clouds.append(c)
    return clouds


def parity_check(ctx, comm, rank, world):
    """One small sharded run on the job's own ranks and transport, compared with the labels the oracle produced for
    the whole cloud (tests/golden/sharded/c3_8strips_400k.npz, committed; the oracle is not imported here).
    Returns {"points", "mismatches", "margin"} for this rank's records."""
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sharded",
                             "c3_8strips_400k.npz"))
    clouds = parity_strips()
    counts = [int(c) for c in g["counts"]]
    assert [c.n for c in clouds] == counts, "the generator no longer makes the fixture's cloud"
    f0, f1 = PARITY_STRIPS * rank // world, PARITY_STRIPS * (rank + 1) // world
    mine = clouds[f0:f1]
    n = load_rank(ctx, mine, PARAMS)
    ctx.shard_run(comm)
    lab = ctx.shard_labels(n)
    want = g["labels"][sum(counts[:f0]):sum(counts[:f1])]
    st = ctx.shard_stats()
    return {"points": int(n), "mismatches": int((lab != want).sum()),
            "margin": int(ctx.stats()["n_margin"]) + int(g["margin"]), "por_max_equal": st["por_max"] == float(g["hyp_max"])}


# ---------------------------------------------------------------------------- bench (N > 1)

def _spread(vals):
    return {"min": round(min(vals), 2), "mean": round(sum(vals) / len(vals), 2), "max": round(max(vals), 2)}


def bench(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from . import synth
    import bench as B
    dev = torch.device("cuda", local)
    # one process per GPU: run on the CPUs next to it, so that the pinned record buffer is local to its PCIe root
    numa = api.bind_to_gpu_numa(local) if os.environ.get("WB_NUMA_BIND", "1") != "0" else -1
    per_gpu = args.points or 125_000_000
    scene = args.scene or 3
    d = synth.describe(scene, per_gpu * world)
    c0 = d.grid_nx * rank // world
    c1 = d.grid_nx * (rank + 1) // world
    cloud = synth.generate(scene, per_gpu * world, seed=scene, region=(c0, 0, c1 - c0, d.grid_ny),
                           gps_base=d.grid_ny * c0)
    n = cloud.n
    pin = api.PinnedBuffer(n * cloud.rec_len)
    pin.array[:] = cloud.records.reshape(-1)
    host_recs = pin.array.reshape(n, cloud.rec_len)
    cloud.records = host_recs
    labels_pin = api.PinnedBuffer(n)
    dev_recs = torch.empty(n * cloud.rec_len, dtype=torch.uint8, device=dev)
    dev_recs.copy_(torch.from_numpy(host_recs.reshape(-1)))
    ctx = api.Context(local)
    ctx.reserve(int(n * 1.25) + 1024)
    comm = nccl_comm(ctx, dist, rank, world)

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step(resident):
        load_rank(ctx, [cloud], B.PARAMS, [dev_recs.data_ptr()] if resident else None)
        ctx.shard_run(comm)
        ctx.shard_labels(n, out=labels_pin.array)

    def timed(resident, steps, collect=None):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            step(resident)
            if collect is not None:
                collect.append((ctx.shard_stats(), ctx.stats(), (time.perf_counter() - t1) * 1e3))
        torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        dist.barrier()
        return allmax(time.perf_counter() - t0) / steps, mine / steps

    tot = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    n_total = int(tot.item())
    # labels of THIS job's ranks over THIS transport against the oracle's (a committed fixture), before any timing
    parity = None
    if world in (2, 4, 8):
        try:
            parity = parity_check(ctx, comm, rank, world)
        except Exception as e:
            parity = {"error": "%s: %s" % (type(e).__name__, e)}
        allp = [None] * world
        dist.all_gather_object(allp, parity)
        parity = allp
    for _ in range(args.warmup):
        step(True)
    l0 = ctx.stats()["kernel_launches"]
    sampler = B.ClockSampler(local)
    sampler.start()
    per_step = []
    dt, _ = timed(True, args.steps, per_step)
    clocks = sampler.stop()
    launches = (ctx.stats()["kernel_launches"] - l0) // max(1, args.steps)
    # e2e: pinned host records in, labels of the own records out, every step.  Serial (copy, then the sharded run,
    # then labels) and pipelined: a second context + communicator per rank, the next step's records are copied and
    # decoded into it by a loader thread while this step's collective run goes on (the collectives themselves stay
    # strictly one after another, in the same order on every rank).
    step(False)
    e2e_steps = max(1, min(args.steps, 3))
    dts, _ = timed(False, e2e_steps)
    h2d_ms = ctx.stats()["ms_h2d"]
    ctx2 = api.Context(local)
    ctx2.reserve(int(n * 1.25) + 1024)
    comm2 = nccl_comm(ctx2, dist, rank, world)
    pair = [(ctx, comm), (ctx2, comm2)]

    def load_e2e(i):
        load_rank(pair[i % 2][0], [cloud], B.PARAMS, None)

    def run_pipelined(steps):
        load_e2e(0)
        for i in range(steps):
            th = None
            if i + 1 < steps:
                th = threading.Thread(target=load_e2e, args=(i + 1,))
                th.start()
            c, cm = pair[i % 2]
            c.shard_run(cm)
            c.shard_labels(n, out=labels_pin.array)
            if th is not None:
                th.join()

    run_pipelined(2)
    pipe_steps = max(2, min(args.steps, 5))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    run_pipelined(pipe_steps)
    torch.cuda.synchronize()
    dist.barrier()
    dte = allmax(time.perf_counter() - t0) / pipe_steps
    comm2.close()
    ctx2.close()
    # what every rank saw (stage wall clocks of the library, averaged over the timed steps)
    keys = [k for k in per_step[0][0] if k.startswith("ms_")]
    mine = {k: sum(s[0][k] for s in per_step) / len(per_step) for k in keys}
    mine["step_wall"] = sum(s[2] for s in per_step) / len(per_step)
    mine["classify_kernel"] = sum(s[1]["ms_classify_kernel"] for s in per_step) / len(per_step)
    mine["h2d_decode_e2e"] = h2d_ms
    last = per_step[-1][0]
    mine_info = {"n_own": n, "n_halo_scan": last["n_halo_scan"], "n_halo_classify": last["n_halo_classify"],
                 "por_max": last["por_max"], "grid_cells": last["grid_cells"], "bytes_sent": last["bytes_sent"]}
    everyone = [None] * world
    dist.all_gather_object(everyone, (mine, mine_info))
    hist = torch.from_numpy(np.bincount(labels_pin.array, minlength=256).astype(np.int64)).to(dev)
    dist.all_reduce(hist)
    # like-for-like single-GPU figure: rank 0 alone on its own strip (the scene's geometry, no halo, no exchange)
    base = None
    # every rank lets go of its communicator at the same point (ncclCommDestroy waits for the peers in NCCL 2.28)
    comm.close()
    ctx.close()
    if not args.no_scaling_base:
        if rank == 0:
            solo = api.Context(local)
            solo.set_params(**B.PARAMS)
            try:
                base = B.scaling_base(solo, args, world, 0, steps=max(2, min(args.steps, 5)), cloud=cloud, dev=dev_recs)
            except Exception as e:
                base = {"error": "%s: %s" % (type(e).__name__, e)}
            solo.close()
        dist.barrier()
    if rank == 0:
        hbm, peak_src = B.peaks()
        ck_ms = mine["classify_kernel"]
        achieved = 13.0 * n / (ck_ms * 1e-3) / 1e9 if ck_ms > 0 else 0.0
        value = n_total / dt
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "timer": "wall clock between barriers, CUDA synchronised on both sides, max over ranks",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 multi-tile aerial cloud: %d points in %d x-strips (one LAS format %d file per "
                                   "GPU), NCCL halo exchange; tileSize 1 maxSlope 1 thickness 0 minHyperboloidSize 0.1"
                                   % (n_total, world, cloud.fmt),
                       "points": n_total, "points_per_gpu": n,
                       "l2": "inputs larger than L2 (%.1f GB of records per GPU per step)" % (n * cloud.rec_len / 1e9),
                       "parallelism": "%d spatial strips; halo rows over NCCL send/recv, tile grid all-reduce (1 B/cell)"
                                      % world},
            "stages_ms": {k[3:] if k.startswith("ms_") else k: _spread([e[0][k] for e in everyone]) for k in mine},
            "classify_kernel_ms_by_rank": [round(e[0]["classify_kernel"], 1) for e in everyone],
            "step_wall_ms_by_rank": [round(e[0]["step_wall"], 1) for e in everyone],
            "halo": {"scan_points": sum(e[1]["n_halo_scan"] for e in everyone),
                     "classify_points": sum(e[1]["n_halo_classify"] for e in everyone),
                     "classify_fraction": sum(e[1]["n_halo_classify"] for e in everyone) / n_total,
                     "por_max": everyone[0][1]["por_max"], "grid_cells": everyone[0][1]["grid_cells"],
                     "bytes_sent_per_step": sum(e[1]["bytes_sent"] for e in everyone)},
            "roofline": {"bound": "hbm", "kernel": "wb_classify_kernel (rank 0)", "achieved": achieved, "peak": hbm,
                         "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": peak_src},
            "e2e": {"value": n_total / dte, "unit": B.UNIT, "h2d_bytes_per_step": n_total * cloud.rec_len,
                    "d2h_bytes_per_step": n_total, "ms_per_step": dte * 1e3, "steps": pipe_steps,
                    "pipelined": "two contexts per rank: step i+1's H2D + decode overlap step i's sharded run and label D2H",
                    "serial": {"value": n_total / dts, "ms_per_step": dts * 1e3, "steps": e2e_steps},
                    "numa_node_rank0": numa,
                    "h2d_decode_ms": _spread([e[0]["h2d_decode_e2e"] for e in everyone]),
                    "h2d_gbs_per_gpu": _spread([n * cloud.rec_len / (e[0]["h2d_decode_e2e"] * 1e-3) / 1e9
                                                for e in everyone if e[0]["h2d_decode_e2e"] > 0] or [0.0])},
            "gpu_launches": int(launches), "clocks": clocks,
            "labels": {"ground": int(hist[2]), "nonground": int(hist[1])},
        }
        if parity is not None:
            ok = all("error" not in p and p["mismatches"] <= p["margin"] and p["por_max_equal"] for p in parity)
            line["parity"] = {"check": "C3 scene, 400 k points in 8 files over these %d ranks and NCCL, against the oracle's "
                                       "labels (tests/golden/sharded/c3_8strips_400k.npz)" % world,
                              "ok": ok, "points": sum(p.get("points", 0) for p in parity),
                              "mismatches": sum(p.get("mismatches", 0) for p in parity),
                              "errors": [p["error"] for p in parity if "error" in p]}
        if base is not None:
            line["weak_scaling_base"] = base
            if "value" in base:
                line["weak_efficiency"] = value / (world * base["value"])
        print(json.dumps(line))
    dist.destroy_process_group()

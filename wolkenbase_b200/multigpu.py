"""Spatial sharding of one cloud over the GPUs of a box (one process per GPU).

The reference has no distributed mode (one std::thread pool).  The path shards
because every label depends only on the points inside that point's downward hyperboloid and
every tile parameter only on the points inside the tile's cylinder:

  strips   each rank owns a strip in x (its own LAS files); the octree cube and the flowsnake
           come from the union of ALL files' header corners, so Morton keys, tile numbers and
           tile centres are identical everywhere;
  halo 1   points within 2.5 tile spacings of another rank's strip are sent there, so each rank
           can scan every tile whose centre lies in its ownership interval;
  merge    the dense tile table (nPoints, treeFlags, hyperboloidSize bits), zero outside the
           ownership interval, is summed over ranks (all-reduce) -> the global table; postscan
           then runs on every rank over the whole table;
  halo 2   a point Q can lie in the hyperboloid of P only if dist_xy <= sqrt(dz^2+2 por dz)/slope
           with dz <= zmax-thickness-Q.z and por <= max hyperboloidSize * slope^2 (shape.cpp:119-135),
           so Q goes to every rank whose strip is that close; each rank rebuilds its bucket
           hierarchy over own+halo points and classifies its OWN points only.
  order    local arrays are [halo from lower ranks | own | halo from higher ranks], each sender's
           points in their original order, so that the canonical order (Morton key, then input
           index) restricted to a rank equals the global one.

The exchange steps only move data (torch.distributed all_to_all / all_reduce over NCCL); all
arithmetic that decides a label runs in the library's kernels.  `LocalComm` runs the same code
for several simulated ranks inside one process (single-GPU parity tests, gloo CPU tests of the
selection logic).
"""
import math
import os
import time

import numpy as np
import torch

SCAN_HALO_SPACINGS = 2.5


# ---------------------------------------------------------------------------- pure helpers

def coords(ints, scale, offset):
    """(offset + scale*int) exactly as las.cpp:808 does it (two roundings), unit 1."""
    return ints.to(torch.float64) * scale + offset


def ownership_bounds(strips):
    """strips: list of (xlo, xhi) per rank, ascending.  Returns [(lo, hi)] half-open ownership
    intervals for tile centres: boundaries at the middle of the gap between adjacent strips."""
    n = len(strips)
    cuts = [-math.inf] + [0.5 * (strips[k][1] + strips[k + 1][0]) for k in range(n - 1)] + [math.inf]
    return [(cuts[k], cuts[k + 1]) for k in range(n)]


def reach_radius(z, zmax, thickness, por_max, slope):
    """Upper bound of the xy distance at which a point of height z can be inside the downward
    hyperboloid of ANY query point (vertex height <= zmax - thickness, polar radius <= por_max)."""
    dz = torch.clamp(zmax - thickness - z, min=0.0)
    return torch.sqrt(dz * dz + 2.0 * por_max * dz) / slope * (1 + 1e-9) + 1e-6


def select_for_strip(x, radius, strip):
    """Mask of points whose x lies within `radius` (scalar or per point) of the strip [lo,hi]."""
    lo, hi = strip
    return (x >= lo - radius) & (x <= hi + radius)


def pack(cols):
    """xi,yi,zi (int32) + cls (uint8) -> one (n,4) int32 tensor for the exchange."""
    xi, yi, zi, cl = cols
    return torch.stack([xi, yi, zi, cl.to(torch.int32)], dim=1).contiguous()


def unpack(t):
    return t[:, 0].contiguous(), t[:, 1].contiguous(), t[:, 2].contiguous(), t[:, 3].to(torch.uint8).contiguous()


# ---------------------------------------------------------------------------- communicators

class TorchComm:
    """torch.distributed (NCCL on the GPU box, gloo in the CPU tests)."""

    def __init__(self, dist):
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    def all_gather_doubles(self, vals, device):
        t = torch.tensor(vals, dtype=torch.float64, device=device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.cpu().tolist() for o in out]

    def all_to_all_rows(self, send_list):
        """send_list[k] = (m_k,4) int32 rows for rank k.  Returns the list received from each rank."""
        dev = send_list[0].device
        counts = torch.tensor([s.shape[0] for s in send_list], dtype=torch.int64, device=dev)
        rcounts = torch.empty_like(counts)
        self.dist.all_to_all_single(rcounts, counts)
        rc = rcounts.cpu().tolist()
        sc = counts.cpu().tolist()
        send = torch.cat(send_list, dim=0).contiguous()
        recv = torch.empty((sum(rc), 4), dtype=torch.int32, device=dev)
        self.dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc)
        return list(torch.split(recv, rc, dim=0))

    def all_reduce_sum(self, tensors):
        for t in tensors:
            self.dist.all_reduce(t)

    def all_reduce_max_scalar(self, v, device):
        t = torch.tensor([v], dtype=torch.float64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        self.dist.barrier()


# ---------------------------------------------------------------------------- one rank

class Rank:
    """The work of one GPU.  ctx_scan and ctx_cls are two wolkenbase_b200.api.Context objects on
    that GPU (the scan stage and the classify stage hold different halo sets)."""

    def __init__(self, rank, world, ctx_scan, ctx_cls, params, device):
        self.rank, self.world = rank, world
        self.a, self.b = ctx_scan, ctx_cls
        self.p = dict(tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1)
        self.p.update(params or {})
        self.device = device
        self.timing = {}

    # -- stage 0: decode own records, learn everybody's extents
    def load(self, cloud, dev_records=None):
        """cloud: synth.Cloud (or anything with records/fmt/scale/offset/min_corner/max_corner)."""
        self.cloud = cloud
        self.n_own = cloud.n
        self.scale, self.offset = cloud.scale, cloud.offset
        a = self.a
        a.clear()
        a.set_params(**self.p)
        a.add_extent(cloud.min_corner, cloud.max_corner)       # only so that decode has a context
        if dev_records is not None:
            a.add_las_device(dev_records.data_ptr(), cloud.n, cloud.fmt, cloud.rec_len, cloud.scale, cloud.offset)
        else:
            a.add_las(cloud.records, cloud.fmt, cloud.scale, cloud.offset)
        n = cloud.n
        self.own = (torch.empty(n, dtype=torch.int32, device=self.device),
                    torch.empty(n, dtype=torch.int32, device=self.device),
                    torch.empty(n, dtype=torch.int32, device=self.device),
                    torch.empty(n, dtype=torch.uint8, device=self.device))
        a.export_points_device(0, n, *[t.data_ptr() for t in self.own])
        self.x = coords(self.own[0], self.scale[0], self.offset[0])
        self.z = coords(self.own[2], self.scale[2], self.offset[2])
        return list(cloud.min_corner) + list(cloud.max_corner)

    def set_extents(self, extents):
        """extents[k] = [minx,miny,minz,maxx,maxy,maxz] of rank k's files (all ranks, same order)."""
        self.extents = extents
        self.strips = [(e[0], e[3]) for e in extents]
        self.owner_bounds = ownership_bounds(self.strips)
        self.zmax = max(e[5] for e in extents)
        self.zmin = min(e[2] for e in extents)

    def _setup(self, ctx):
        ctx.clear()
        ctx.set_params(**self.p)
        for e in self.extents:
            ctx.add_extent(e[0:3], e[3:6])

    def _sends(self, radius):
        rows = pack(self.own)
        out = []
        for k in range(self.world):
            if k == self.rank:
                out.append(rows[:0])
            else:
                out.append(rows[select_for_strip(self.x, radius, self.strips[k])])
        return out

    def _fill(self, ctx, received):
        """local array = [halo from lower ranks | own | halo from higher ranks]"""
        first = 0
        for k in range(self.world):
            if k == self.rank:
                own_first = first
                ctx.add_points_device(*[t.data_ptr() for t in self.own], self.n_own, self.scale, self.offset)
                first += self.n_own
            elif received[k].shape[0]:
                cols = unpack(received[k])
                ctx.add_points_device(*[t.data_ptr() for t in cols], cols[0].shape[0], self.scale, self.offset)
                first += cols[0].shape[0]
        ctx.set_own_range(own_first, own_first + self.n_own)
        self.own_first = own_first
        return first

    # -- stage 1: scan with the narrow halo
    def scan_sends(self):
        self._setup(self.a)
        g = self.a.geometry()
        self.geom = g
        return self._sends(SCAN_HALO_SPACINGS * g.spacing)

    def scan(self, received):
        a = self.a
        self.n_scan = self._fill(a, received)
        a.build()
        a.scan()
        T = self.geom.snake_hi - self.geom.snake_lo + 1
        self.t_np = torch.empty(T, dtype=torch.int32, device=self.device)
        self.t_tree = torch.empty(T, dtype=torch.int32, device=self.device)
        self.t_hyp = torch.empty(T, dtype=torch.int64, device=self.device)
        lo, hi = self.owner_bounds[self.rank]
        a.export_tiles_device(lo if lo > -math.inf else -1e300, hi if hi < math.inf else 1e300,
                              self.t_np.data_ptr(), self.t_tree.data_ptr(), self.t_hyp.data_ptr())
        return [self.t_np, self.t_tree, self.t_hyp]

    # -- stage 2: global table -> postscan -> how far classify can reach
    def postscan(self):
        a = self.a
        a.import_tiles_device(self.t_np.data_ptr(), self.t_tree.data_ptr(), self.t_hyp.data_ptr())
        a.postscan()
        a.export_tiles_device(-1e300, 1e300, self.t_np.data_ptr(), self.t_tree.data_ptr(), self.t_hyp.data_ptr())
        s = self.p["max_slope"]
        self.por_max = a.max_hyperboloid_size() * s * s
        return self.por_max

    def classify_sends(self):
        r = reach_radius(self.z, self.zmax, self.p["thickness"], self.por_max, self.p["max_slope"])
        return self._sends(r)

    def classify(self, received, labels_out=None):
        b = self.b
        self._setup(b)
        self.n_cls = self._fill(b, received)
        b.build()
        b.assign()
        b.import_tiles_device(self.t_np.data_ptr(), self.t_tree.data_ptr(), self.t_hyp.data_ptr(), postscanned=True)
        b.classify()
        lab = b.labels(self.n_cls)
        own = lab[self.own_first:self.own_first + self.n_own]
        if labels_out is not None:
            labels_out[:] = own
            return labels_out
        return own.copy()


# ---------------------------------------------------------------------------- drivers

def run_local(ranks, clouds):
    """All ranks in one process (tests): exchanges are list shuffles, the reduction a sum."""
    W = len(ranks)
    ext = [r.load(c) for r, c in zip(ranks, clouds)]
    for r in ranks:
        r.set_extents(ext)
    sends = [r.scan_sends() for r in ranks]
    tables = [r.scan([sends[src][r.rank] for src in range(W)]) for r in ranks]
    summed = [sum(t[i] for t in tables) for i in range(3)]
    for r in ranks:
        for i, t in enumerate([r.t_np, r.t_tree, r.t_hyp]):
            t.copy_(summed[i])
        r.postscan()
    sends = [r.classify_sends() for r in ranks]
    return [r.classify([sends[src][r.rank] for src in range(W)]) for r in ranks]


def run_distributed(rank_obj, cloud, comm, dev_records=None, labels_out=None, timing=None):
    """One rank under torch.distributed.  timing: optional dict that receives per-stage milliseconds."""
    r = rank_obj
    t = [time.perf_counter()]

    def lap(name):
        if timing is not None:
            torch.cuda.synchronize()
            now = time.perf_counter()
            timing[name] = timing.get(name, 0.0) + (now - t[0]) * 1e3
            t[0] = now

    ext = comm.all_gather_doubles(r.load(cloud, dev_records), r.device)
    r.set_extents(ext)
    lap("load_decode")
    sends = r.scan_sends()
    lap("halo1_select")
    recv = comm.all_to_all_rows(sends)
    lap("halo1_exchange")
    tabs = r.scan(recv)
    lap("build_scan_export")
    comm.all_reduce_sum(tabs)
    lap("tile_allreduce")
    r.postscan()
    lap("postscan")
    sends = r.classify_sends()
    lap("halo2_select")
    recv = comm.all_to_all_rows(sends)
    lap("halo2_exchange")
    out = r.classify(recv, labels_out)
    lap("build_classify_labels")
    return out


# ---------------------------------------------------------------------------- bench (N > 1)

def bench(args, rank, world, local):
    import json
    import torch.distributed as dist
    from . import api, synth
    import bench as B
    dev = torch.device("cuda", local)
    comm = TorchComm(dist)
    per_gpu = args.points or 125_000_000
    scene = args.scene or 3
    d = synth.describe(scene, per_gpu * world)
    c0 = d.grid_nx * rank // world
    c1 = d.grid_nx * (rank + 1) // world
    base = d.grid_ny * c0
    cloud = synth.generate(scene, per_gpu * world, seed=scene + rank * 0, region=(c0, 0, c1 - c0, d.grid_ny),
                           gps_base=base)
    n = cloud.n
    pin = api.PinnedBuffer(n * cloud.rec_len)
    pin.array[:] = cloud.records.reshape(-1)
    host_recs = pin.array.reshape(n, cloud.rec_len)
    cloud.records = host_recs
    labels_pin = api.PinnedBuffer(n)
    dev_recs = torch.empty(n * cloud.rec_len, dtype=torch.uint8, device=dev)
    dev_recs.copy_(torch.from_numpy(host_recs.reshape(-1)))
    ctx_a, ctx_b = api.Context(local), api.Context(local)
    cap = int(n * 1.8) + 1024
    ctx_a.reserve(int(n * 1.1) + 1024)
    ctx_b.reserve(cap)
    R = Rank(rank, world, ctx_a, ctx_b, B.PARAMS, dev)
    n_total = int(comm.all_reduce_max_scalar(0, dev))  # warm the communicator
    tot = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    n_total = int(tot.item())

    stage_ms = {}

    def step(resident, timing=None):
        return run_distributed(R, cloud, comm, dev_recs if resident else None, labels_pin.array, timing)

    for _ in range(args.warmup):
        step(True)
    l0 = ctx_a.stats()["kernel_launches"] + ctx_b.stats()["kernel_launches"]
    sampler = B.ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    comm.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True, stage_ms)
    torch.cuda.synchronize()
    comm.barrier()
    dt = comm.all_reduce_max_scalar(time.perf_counter() - t0, dev)
    clocks = sampler.stop()
    launches = (ctx_a.stats()["kernel_launches"] + ctx_b.stats()["kernel_launches"] - l0) // max(1, args.steps)
    sa, sb = ctx_a.stats(), ctx_b.stats()
    # e2e: pinned host records in, labels of the own points out
    step(False)
    torch.cuda.synchronize()
    comm.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        step(False)
    torch.cuda.synchronize()
    comm.barrier()
    dte = comm.all_reduce_max_scalar(time.perf_counter() - t0, dev) / e2e_steps
    halo = torch.tensor([R.n_scan - n, R.n_cls - n], dtype=torch.int64, device=dev)
    dist.all_reduce(halo)
    hist = torch.from_numpy(np.bincount(labels_pin.array, minlength=256).astype(np.int64)).to(dev)
    dist.all_reduce(hist)
    if rank == 0:
        hbm, peak_src = B.peaks()
        ck_ms = sb["ms_classify_kernel"]
        achieved = 13.0 * n / (ck_ms * 1e-3) / 1e9 if ck_ms > 0 else 0.0
        line = {
            "metric": B.METRIC, "value": n_total / (dt / args.steps), "unit": B.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 multi-tile aerial cloud: %d points in %d x-strips (one LAS format %d file per "
                                   "GPU), NCCL halo exchange; tileSize 1 maxSlope 1 thickness 0 minHyperboloidSize 0.1"
                                   % (n_total, world, cloud.fmt),
                       "points": n_total, "points_per_gpu": n,
                       "l2": "inputs larger than L2 (%.1f GB of records per GPU per step)" % (n * cloud.rec_len / 1e9),
                       "parallelism": "%d spatial strips, halo exchange + tile-table all-reduce" % world},
            "phases_ms_rank0": {"scan_stage_build": sa["ms_build"], "scan": sa["ms_scan"], "postscan": sa["ms_postscan"],
                                "classify_stage_build": sb["ms_build"], "classify": sb["ms_classify"]},
            "stages_ms_rank0": {k: round(v / args.steps, 2) for k, v in stage_ms.items()},
            "halo": {"scan_points": int(halo[0]), "classify_points": int(halo[1]),
                     "classify_fraction": float(halo[1]) / n_total, "por_max": R.por_max},
            "roofline": {"bound": "hbm", "kernel": "wb_classify_kernel (rank 0)", "achieved": achieved, "peak": hbm,
                         "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": peak_src},
            "e2e": {"value": n_total / dte, "unit": B.UNIT, "h2d_bytes_per_step": n_total * cloud.rec_len,
                    "d2h_bytes_per_step": n_total, "ms_per_step": dte * 1e3},
            "gpu_launches": int(launches), "clocks": clocks,
            "labels": {"ground": int(hist[2]), "nonground": int(hist[1])},
        }
        print(json.dumps(line))
    ctx_a.close()
    ctx_b.close()
    dist.destroy_process_group()

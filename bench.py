#!/usr/bin/env python
"""bench.py — LAS points classified per second through the ground-extraction hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                     (the reference's own CPU path, bounded sample)

A step = one full pass (decode -> Morton sort -> leaf split -> tile scan -> postscan -> classify)
over one synthetic cloud.  N=1 workload: BASELINE.json configs[1], the 100 M-point aerial tile.
`value` starts with the packed LAS records resident in HBM; `e2e` starts with them in pinned host
memory and ends with the class bytes back on the host.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "LAS points classified/sec"
UNIT = "points/s"
PARAMS = dict(tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def workload(args, world):
    """Scene, per-rank region and global grid for this run."""
    from wolkenbase_b200 import synth
    if args.points:
        per_gpu = args.points
    else:
        per_gpu = 100_000_000 if world == 1 else 125_000_000
    scene = args.scene or (2 if world == 1 else 3)
    d = synth.describe(scene, per_gpu * world)
    name = {1: "C1 street 10M", 2: "C2 aerial tile", 3: "C3 multi-tile aerial (format 6)",
            4: "C4 terrestrial", 5: "C5 steep urban"}[scene]
    return scene, d, per_gpu, name


def run_ours(args):
    import torch
    import torch.distributed as dist
    from wolkenbase_b200 import api, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torch.distributed.run with %d ranks" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from wolkenbase_b200 import multigpu
        return multigpu.bench(args, rank, world, local)

    scene, d, per_gpu, wname = workload(args, 1)
    t0 = time.time()
    cloud = synth.generate(scene, per_gpu, seed=scene)
    n, rec_len = cloud.n, cloud.rec_len
    gen_s = time.time() - t0
    # pinned host copy of the file body (what a reader would hand over)
    pin = api.PinnedBuffer(n * rec_len)
    pin.array[:] = cloud.records.reshape(-1)
    host_recs = pin.array.reshape(n, rec_len)
    labels_pin = api.PinnedBuffer(n)
    dev_recs = torch.empty(n * rec_len, dtype=torch.uint8, device="cuda")
    dev_recs.copy_(torch.from_numpy(host_recs.reshape(-1)), non_blocking=False)
    torch.cuda.synchronize()

    ctx = api.Context(local)
    ctx.set_params(**PARAMS)
    ctx.reserve(n)

    def step_resident():
        ctx.clear()
        ctx.add_extent(cloud.min_corner, cloud.max_corner)
        ctx.add_las_device(dev_recs.data_ptr(), n, cloud.fmt, rec_len, cloud.scale, cloud.offset)
        ctx.run()

    def step_e2e():
        ctx.clear()
        ctx.add_extent(cloud.min_corner, cloud.max_corner)
        ctx.add_las(host_recs, cloud.fmt, cloud.scale, cloud.offset)
        ctx.run()
        ctx.labels(n, out=labels_pin.array)

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    launches0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    sampler.start()
    phase = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.mark(0)                       # CUDA events on the library's own stream bracket the K steps
    for _ in range(args.steps):
        step_resident()
        st = ctx.stats()
        for k, v in st.items():
            if k.startswith("ms_"):
                phase[k] = phase.get(k, 0.0) + v
    ctx.mark(1)
    dt = ctx.mark_elapsed(0, 1) * 1e-3
    ctx.sync()
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st = ctx.stats()
    launches = (st["kernel_launches"] - launches0) // max(1, args.steps)
    ms_step = dt / args.steps * 1e3
    value = n / (dt / args.steps)
    for k in phase:
        phase[k] /= args.steps

    # e2e: pinned host records in, labels out, every step.  Serial first (copy, then compute, then labels), then the
    # way a job with more than one cloud runs: two contexts, the next step's records are copied and decoded into one
    # (loader thread) while the other classifies — every step's H2D and D2H still lie inside the timed region.
    step_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    ctx.mark(2)
    for _ in range(e2e_steps):
        step_e2e()
    ctx.mark(3)
    dte_serial = ctx.mark_elapsed(2, 3) * 1e-3 / e2e_steps
    torch.cuda.synchronize()
    dte_serial_wall = (time.perf_counter() - t0) / e2e_steps
    ste = ctx.stats()
    ctx2 = api.Context(local)
    ctx2.set_params(**PARAMS)
    ctx2.reserve(n)
    pair = [ctx, ctx2]

    def load_e2e(i):
        c = pair[i % 2]
        c.clear()
        c.add_extent(cloud.min_corner, cloud.max_corner)
        c.add_las(host_recs, cloud.fmt, cloud.scale, cloud.offset)

    def run_pipelined(steps):
        load_e2e(0)
        for i in range(steps):
            th = None
            if i + 1 < steps:
                th = threading.Thread(target=load_e2e, args=(i + 1,))
                th.start()
            pair[i % 2].run()
            pair[i % 2].labels(n, out=labels_pin.array)
            if th is not None:
                th.join()

    run_pipelined(2)                                   # warm-up of the second context
    pipe_steps = max(2, min(args.steps, 5))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_pipelined(pipe_steps)
    torch.cuda.synchronize()
    dte = (time.perf_counter() - t0) / pipe_steps
    dte_wall = dte
    ctx2.close()
    hist = ctx.count_classes()

    hbm, peak_src = peaks()
    # DRAM bytes of one launch of the dominant kernel from the latest committed `ncu --set full` capture of THIS kernel
    # version on THIS workload (profiles/r2_traffic.json names the capture); null when the point count differs
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["wb_classify_kernel"]
        if abs(tj["points"] - n) <= 0.01 * n and scene == tj.get("scene", 2):
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    ck_ms = phase.get("ms_classify_kernel", 0.0)
    alg_bytes = 13.0 * n                                     # 12 B read + 1 B written per point
    achieved = alg_bytes / (ck_ms * 1e-3) / 1e9 if ck_ms > 0 else 0.0
    sort_ms = phase.get("ms_sort", 0.0)
    sort_bytes = 32.0 * 8 * n                                # 8 passes x (8 read + 12 read + 12 written)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "ms_per_step_wall": dt_wall / args.steps * 1e3,
        "timer": "CUDA events on the library's compute stream around the K steps (wb_mark); wall clock alongside",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d points, LAS format %d (%d B records), scene %d seed %d; "
                               "tileSize 1 maxSlope 1 thickness 0 minHyperboloidSize 0.1" %
                               (wname, n, cloud.fmt, rec_len, scene, scene),
                   "points": n, "l2": "inputs larger than L2 (%.1f GB of records per step)" % (n * rec_len / 1e9),
                   "parallelism": "1 GPU"},
        "phases_ms": {k[3:]: round(v, 3) for k, v in sorted(phase.items())},
        "roofline": {"bound": "hbm", "kernel": "wb_classify_kernel", "achieved": achieved, "peak": hbm,
                     "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "note": "algorithmic 13 B/point; the kernel is bound by its FP64/shared-memory inner loop "
                             "(SURVEY 8d), DRAM traffic from ncu is in profiles/"},
        "roofline_sort": {"bound": "hbm", "kernel": "radix sort (8 passes)", "achieved": sort_bytes / (sort_ms * 1e-3) / 1e9
                          if sort_ms > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                          "frac": (sort_bytes / (sort_ms * 1e-3) / 1e9 / hbm) if sort_ms > 0 else 0.0},
        "roofline_decode": {"bound": "hbm", "kernel": "wb_decode_kernel", "achieved": (rec_len + 14.0) * n / (phase.get("ms_decode", 0.0) * 1e-3) / 1e9
                            if phase.get("ms_decode", 0.0) > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                            "frac": ((rec_len + 14.0) * n / (phase.get("ms_decode", 0.0) * 1e-3) / 1e9 / hbm)
                            if phase.get("ms_decode", 0.0) > 0 else 0.0,
                            "note": "L+14 B/point: record read, 3 x int32 + class + return number written"},
        "e2e": {"value": n / dte, "unit": UNIT, "h2d_bytes_per_step": n * rec_len, "d2h_bytes_per_step": n,
                "ms_per_step": dte * 1e3, "ms_per_step_wall": dte_wall * 1e3, "steps": pipe_steps,
                "timer": "wall clock around the K steps, device synchronised on both sides",
                "pipelined": "two contexts: step i+1's H2D + decode overlap step i's build/scan/classify and label D2H",
                "serial": {"value": n / dte_serial, "ms_per_step": dte_serial * 1e3, "ms_per_step_wall": dte_serial_wall * 1e3,
                           "steps": e2e_steps},
                "ms_h2d_decode": ste["ms_h2d"], "ms_d2h": ste["ms_d2h"]},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "labels": {"ground": int(hist[2]), "nonground": int(hist[1]), "margin_points": int(st["n_margin"]),
                   "untiled": int(st["n_untiled"])},
        "leaves": int(st["n_leaves"]), "tiles_nonempty": int(st["n_tiles_nonempty"]),
        "classify_work": {"nodes_per_point": st["cl_nodes"] * 32.0 / n, "chunks_per_point": st["cl_chunks"] * 32.0 / n,
                          "pair_tests_per_point": st["cl_pairs"] * 32.0 / n, "second_walk_points": int(st["n_second_walk"]),
                          "second_walk": {"warps_frac": st["cl_warps2"] * 32.0 / n, "nodes": st["cl_nodes2"] * 32.0 / n,
                                          "chunks": st["cl_chunks2"] * 32.0 / n, "pairs": st["cl_pairs2"] * 32.0 / n}},
        "gen_s": round(gen_s, 2),
    }
    if not args.no_scaling_base and not args.points and not args.scene:
        try:
            line["weak_scaling_base"] = scaling_base(ctx, args)
        except Exception as e:                      # an auxiliary figure must never cost the bench line
            line["weak_scaling_base"] = {"error": "%s: %s" % (type(e).__name__, e)}
            ctx.close()
            ctx = api.Context(local)
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(scene, args.cpu_points, bounded=True)
    ctx.close()
    print(json.dumps(line))


def scaling_base(ctx, args, world=8, rank=0, steps=2, cloud=None, dev=None):
    """The N>1 runs shard BASELINE configs[2] (C3, 125 M points per GPU, LAS format 6) while the N=1 line is
    configs[1] (C2), whose tile spacing and hyperboloid sizes make a point several times dearer.  So that
    N-GPU values can be set against a like-for-like single-GPU figure, the N=1 line also carries one GPU's
    throughput on ONE GPU'S SHARE of the 8-GPU scene: rank 0's x-strip of the 1 B-point cloud, with the whole
    scene's geometry (same root cube, same tile lattice), no halo and no exchange.  Records resident in HBM,
    CUDA events on the library's stream, 1 warm-up + `steps` timed passes.  (`--strip W:R` runs the same thing for
    rank R's strip of the W-GPU scene and prints it alone: the tool that showed where the 4-GPU run of round 1 went.)"""
    import torch
    from wolkenbase_b200 import synth
    per_gpu = args.points or 125_000_000
    scene = args.scene or 3
    d = synth.describe(scene, per_gpu * world)
    if cloud is None:
        c0, c1 = d.grid_nx * rank // world, d.grid_nx * (rank + 1) // world
        cloud = synth.generate(scene, per_gpu * world, seed=scene, region=(c0, 0, c1 - c0, d.grid_ny),
                               gps_base=d.grid_ny * c0)
    n = cloud.n
    lo = (d.offset[0], d.offset[1], cloud.min_corner[2])
    hi = (d.offset[0] + d.scale * d.extent_ticks, d.offset[1] + d.scale * d.extent_ticks, cloud.max_corner[2])
    if dev is None:
        dev = torch.from_numpy(cloud.records.reshape(-1)).cuda()
    torch.cuda.synchronize()

    def step():
        ctx.clear()
        ctx.add_extent(cloud.min_corner, cloud.max_corner)
        ctx.add_extent(lo, hi)                     # the other strips' corners: the scene's full xy extent
        ctx.add_las_device(dev.data_ptr(), n, cloud.fmt, cloud.rec_len, cloud.scale, cloud.offset)
        ctx.run()

    step()
    phase = {}
    ctx.mark(4)
    for _ in range(steps):
        step()
        for k, v in ctx.stats().items():
            if k.startswith("ms_"):
                phase[k[3:]] = round(phase.get(k[3:], 0.0) + v / steps, 3)
    ctx.mark(5)
    ms = ctx.mark_elapsed(4, 5) / steps
    st = ctx.stats()
    g = ctx.geometry()
    hist = ctx.count_classes()
    del dev
    return {"workload": "multi-tile aerial scene, rank %d's strip of the %d-GPU run: %d points, LAS format %d "
                        "(%d B records), the whole scene's geometry (tile spacing %.3f m), no halo"
                        % (rank, world, n, cloud.fmt, cloud.rec_len, g.spacing),
            "value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": 1,
            "classify_kernel_ms": st["ms_classify_kernel"], "phases_ms": phase,
            "max_hyperboloid_size": ctx.max_hyperboloid_size(),
            "labels": {"ground": int(hist[2]), "nonground": int(hist[1])},
            "classify_work": {"nodes": st["cl_nodes"] * 32.0 / n, "chunks": st["cl_chunks"] * 32.0 / n,
                              "pairs": st["cl_pairs"] * 32.0 / n}}


def cpu_baseline(scene, sample_points, bounded=True, threads=None):
    """The reference's CPU path on a bounded sample of the same scene: oracle/_ref (the compiled
    reference, all host threads) if it is there, else the oracle port."""
    from wolkenbase_b200 import synth
    from oracle import wb_oracle
    cores = os.cpu_count() or 1
    cloud = synth.generate(scene, sample_points, seed=scene)
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    sample = "%d-point crop of the same scene (same density and surface model)" % cloud.n
    if os.path.exists(ref):
        nthreads = threads or min(cores, 64)
        with tempfile.TemporaryDirectory() as td:
            las = os.path.join(td, "sample.las")
            cloud.write(las)
            t0 = time.time()
            try:
                out = subprocess.run([ref, "-t", str(nthreads), "-o", os.path.join(td, "ref"), las],
                                     capture_output=True, text=True, cwd=td, timeout=600)
                line = [l for l in out.stdout.splitlines() if "\"points\"" in l][-1]
                info = json.loads(line[line.index("{"):])
                total = info["read_build_s"] + info["scan_s"] + info["postscan_s"] + info["classify_s"]
                return {"value": cloud.n / total, "unit": UNIT, "cores": nthreads, "kind": "reference",
                        "sample": sample, "seconds": round(total, 2),
                        "phases_s": {k: info[k] for k in ("read_build_s", "scan_s", "postscan_s", "classify_s")}}
            except Exception as e:  # hung or crashed multithreaded reference: fall back to the port
                sample += " (compiled reference failed: %s)" % type(e).__name__
    t0 = time.time()
    wb_oracle.run([wb_oracle.file_from_cloud(cloud)], **PARAMS)
    dt = time.time() - t0
    return {"value": cloud.n / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "seconds": round(dt, 2)}


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref, all host threads) on a
    bounded sample of OUR arm's workload: same config, metric, unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from wolkenbase_b200 import synth
    world = max(1, args.gpus)
    scene, d, per_gpu, wname = workload(args, world)
    rec_len = synth.lib().wb_synth_record_length(d.fmt)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(scene, args.cpu_points)
        if i >= args.warmup:
            vals.append(last["value"])
        if last["seconds"] > 60:          # keep the whole run within minutes
            if not vals:
                vals.append(last["value"])
            break
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": last["seconds"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d points, LAS format %d (%d B records), scene %d seed %d; "
                                   "tileSize 1 maxSlope 1 thickness 0 minHyperboloidSize 0.1" %
                                   (wname, per_gpu * world, d.fmt, rec_len, scene, scene),
                       "points": per_gpu * world,
                       "sample": "each step = the reference's CPU path over a %s" % last["sample"]},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=0, help="points per GPU (default: the BASELINE config)")
    ap.add_argument("--scene", type=int, default=0)
    ap.add_argument("--cpu-points", type=int, default=400_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true",
                    help="skip the single-GPU run of the multi-GPU workload (weak_scaling_base)")
    ap.add_argument("--strip", default="", help="W:R — only time rank R's strip of the W-GPU C3 scene on one GPU")
    args = ap.parse_args()
    if args.strip:
        from wolkenbase_b200 import api
        w, r = (int(v) for v in args.strip.split(":"))
        ctx = api.Context(0)
        ctx.set_params(**PARAMS)
        print(json.dumps(scaling_base(ctx, args, w, r, steps=max(1, args.steps))))
        ctx.close()
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

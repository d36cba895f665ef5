"""The sharded pipeline at full size on ONE GPU: W ranks as host threads of this process (LOCAL transport), each with
its own context and its strip of the W-GPU C3 scene, all on device 0.  Prints the per-rank stage times (the ranks
share the GPU, so these say where the work is, not how fast N GPUs are) and, with --compare-single, checks the
gathered labels against ONE context classifying the whole cloud — the sharding's parity at bench scale.
    python tools/shard_check.py --world 4 --points 125000000 --compare-single"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wolkenbase_b200 import api, multigpu, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=4)
    ap.add_argument("--points", type=int, default=125_000_000, help="per rank")
    ap.add_argument("--scene", type=int, default=3)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--compare-single", action="store_true")
    args = ap.parse_args()
    import torch
    W = args.world
    d = synth.describe(args.scene, args.points * W)
    clouds, devs = [], []
    for r in range(W):
        c0, c1 = d.grid_nx * r // W, d.grid_nx * (r + 1) // W
        c = synth.generate(args.scene, args.points * W, seed=args.scene, region=(c0, 0, c1 - c0, d.grid_ny),
                           gps_base=d.grid_ny * c0)
        clouds.append(c)
        devs.append(torch.from_numpy(c.records.reshape(-1)).cuda())
    torch.cuda.synchronize()
    out = {"world": W, "points": [c.n for c in clouds]}
    want = None
    if args.compare_single:
        ctx = api.Context(0)
        ctx.set_params(**multigpu.PARAMS)
        for c in clouds:
            ctx.add_extent(c.min_corner, c.max_corner)
        for c, dv in zip(clouds, devs):
            ctx.add_las_device(dv.data_ptr(), c.n, c.fmt, c.rec_len, c.scale, c.offset)
        t = time.time()
        ctx.run()
        want = ctx.labels(sum(c.n for c in clouds))
        st = ctx.stats()
        out["single"] = {"seconds": round(time.time() - t, 2), "ms_classify_kernel": st["ms_classify_kernel"],
                         "n_duplicates": int(st["n_duplicates"]), "tiles": int(st["n_tiles_nonempty"]),
                         "ground": int((want == 2).sum()), "nonground": int((want == 1).sum())}
        ctx.close()
    group = api.LocalGroup(W)
    res = [None] * W

    def work(r):
        ctx = api.Context(0)
        comm = api.Comm.local(ctx, group, r)
        c = clouds[r]
        per = []
        for _ in range(args.steps):
            t = time.time()
            multigpu.load_rank(ctx, [c], None, [devs[r].data_ptr()])
            ctx.shard_run(comm)
            lab = ctx.shard_labels(c.n)
            per.append((time.time() - t) * 1e3)
        s, w = ctx.shard_stats(), ctx.stats()
        res[r] = (lab, s, w, per)
        comm.close()
        ctx.close()

    th = [threading.Thread(target=work, args=(r,)) for r in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    out["ranks"] = []
    for r in range(W):
        lab, s, w, per = res[r]
        out["ranks"].append({"step_ms": [round(p, 1) for p in per],
                             "stages_ms": {k[3:]: round(v, 1) for k, v in s.items() if k.startswith("ms_")},
                             "classify_kernel_ms": round(w["ms_classify_kernel"], 1), "halo_scan": int(s["n_halo_scan"]),
                             "halo_classify": int(s["n_halo_classify"]), "por_max": s["por_max"],
                             "grid_cells": int(s["grid_cells"]), "n_duplicates": int(w["n_duplicates"])})
    if want is not None:
        got = np.concatenate([res[r][0] for r in range(W)])
        out["mismatches_vs_single_context"] = int((got != want).sum())
    print(json.dumps(out))


if __name__ == "__main__":
    main()

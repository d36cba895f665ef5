"""Work model of the classify kernel on the real bench workload, without a GPU: the oracle builds the canonical
order and the tile table of the C2 scene at full size, then the SIMT emulator (tests/simt) runs the kernel source
on a sample of warps and reports the per-warp work counters.  Usage:
    python tools/model_bench_scene.py POINTS WARPS [VARIANT_FLAGS...]      e.g. 100000000 300 "-DWB_CL_REFILTER=1"
The prepared scene is cached in /tmp (npz) so that variants can be compared on identical inputs.
WB_MODEL_RUN=64 samples contiguous runs of 64 chunks instead of single ones (needed to model WB_CL_COMPACT2, which
gathers the pending queries of neighbouring chunks); WB_MODEL_SITES=N prints the N busiest intrinsic call sites."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
from oracle import wb_oracle as O  # noqa: E402
from wolkenbase_b200 import synth  # noqa: E402
import emul  # noqa: E402


def prepare(scene, n_points, seed):
    cache = "/tmp/wb_model_scene%d_%d.npz" % (scene, n_points)
    if os.path.exists(cache):
        z = np.load(cache)
        return z["pts"], z["hyp"]
    L = O.lib()
    t = time.time()
    cloud = synth.generate(scene, n_points, seed=seed)
    n = cloud.n
    xyz = np.ascontiguousarray(cloud.ints())
    pts = np.empty((n, 3), dtype=np.float64)
    L.wbo_coords(xyz.ctypes.data, n, O._d3(cloud.scale), O._d3(cloud.offset), 1.0, pts.ctypes.data)
    del xyz
    corners = np.ascontiguousarray(np.array([cloud.min_corner, cloud.max_corner], dtype=np.float64))
    center, side, cube = (C.c_double * 3)(), C.c_double(), (C.c_double * 4)()
    L.wbo_size_fit(corners.ctypes.data, 2, center, C.byref(side))
    L.wbo_bbox_cube(corners.ctypes.data, 2, cube)
    keys = np.empty(n, dtype=np.uint64)
    order = np.empty(n, dtype=np.uint32)
    L.wbo_sort(pts.ctypes.data, n, center, side.value, keys.ctypes.data, order.ctypes.data)
    del keys
    pts = np.ascontiguousarray(pts[order])
    del order
    print("sorted %d points in %.0f s" % (n, time.time() - t), flush=True)
    cap = n // 4 + 1024
    tiles = np.zeros(cap, dtype=O.TILE_DTYPE)
    nt = L.wbo_scan(pts.ctypes.data, n, cube, 1.0, 0.1, tiles.ctypes.data, cap)
    assert 0 <= nt <= cap, nt
    tiles = tiles[:nt].copy()
    spacing, lo, hi = C.c_double(), C.c_int(), C.c_int()
    L.wbo_snake_set_size(cube[3], 1.0, C.byref(spacing), C.byref(lo), C.byref(hi))
    L.wbo_postscan(tiles.ctypes.data, nt, spacing.value)
    print("%d tiles, spacing %.3f, %.0f s" % (nt, spacing.value, time.time() - t), flush=True)
    hyp = np.empty(n, dtype=np.float64)
    L.wbo_point_hyperboloid_sizes(pts.ctypes.data, n, cube, 1.0, tiles.ctypes.data, nt, hyp.ctypes.data)
    print("hyperboloid sizes: median %.2f max %.1f, %.0f s" % (np.nanmedian(hyp), np.nanmax(hyp), time.time() - t), flush=True)
    np.savez(cache, pts=pts, hyp=hyp)
    return pts, hyp


def _spread(v):
    v = v & np.uint64(0xfffff)
    for sh, mask in ((16, 0x0000ffff0000ffff), (8, 0x00ff00ff00ff00ff), (4, 0x0f0f0f0f0f0f0f0f),
                     (2, 0x3333333333333333), (1, 0x5555555555555555)):
        v = (v | (v << np.uint64(sh))) & np.uint64(mask)
    return v


def _hilbert(ix, iy, bits):
    """Hilbert index of integer cells (vectorised x,y -> d)."""
    x, y = ix.astype(np.int64), iy.astype(np.int64)
    d = np.zeros(len(x), dtype=np.int64)
    s = 1 << (bits - 1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        x &= s - 1
        y &= s - 1
        s >>= 1
    return d


def reorder(pts, hyp, how):
    """WB_MODEL_ORDER=xy|hilbert: the same points chunked along a 2-D curve over xy instead of the canonical
    (3-D Morton) order — what a classify-only re-sort would give."""
    if not how:
        return pts, hyp
    x, y = pts[:, 0], pts[:, 1]
    bits = 20
    side = max(x.max() - x.min(), y.max() - y.min()) + 1e-9
    ix = ((x - x.min()) / side * (1 << bits)).astype(np.uint64)
    iy = ((y - y.min()) / side * (1 << bits)).astype(np.uint64)
    key = _hilbert(ix, iy, bits) if how == "hilbert" else (_spread(ix) | (_spread(iy) << np.uint64(1)))
    o = np.argsort(key, kind="stable")
    return np.ascontiguousarray(pts[o]), np.ascontiguousarray(hyp[o])


def main():
    n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    n_warps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    variants = sys.argv[3:] or [""]
    scene = int(os.environ.get("WB_MODEL_SCENE", "2"))
    pts, hyp = prepare(scene, n_points, scene)
    pts, hyp = reorder(pts, hyp, os.environ.get("WB_MODEL_ORDER", ""))
    n = len(pts)
    n_chunks = (n + 31) // 32
    rng = np.random.default_rng(1)
    run = int(os.environ.get("WB_MODEL_RUN", "1"))          # warps per sample: >1 takes contiguous runs of chunks
    sample = np.sort(rng.choice(max(1, n_chunks - run), size=min(max(1, n_warps // run), n_chunks), replace=False))
    for i, v in enumerate(variants):
        tot = {}
        t = time.time()
        for c in sample:
            _, w = emul.classify(pts, hyp, chunks=(int(c), int(c) + run), variant=v, out="libwb_simt_m%d.so" % i)
            for k, x in w.items():
                tot[k] = tot.get(k, 0) + x
        m = len(sample) * run
        sites = emul.site_counts(variant=v, out="libwb_simt_m%d.so" % i)
        trips = emul.emu_counts(variant=v, out="libwb_simt_m%d.so" % i)
        print("  reach tests at expansion per warp: chunks %.1f, internal nodes %.1f" % (trips[0] / m, trips[1] / m))
        print("variant %r: per warp over %d warps: nodes %.1f chunks %.1f pairs %.1f | pass 2: nodes %.1f chunks %.1f "
              "pairs %.1f | intrinsics %.0f  (%.0f s)" %
              (v, m, tot["nodes"] / m, tot["chunks"] / m, tot["pairs"] / m, tot["nodes2"] / m, tot["chunks2"] / m,
               tot["pairs2"] / m, tot["collectives"] / m, time.time() - t), flush=True)
        src = open(os.path.join(ROOT, "wolkenbase_b200", "csrc", "wb_kernels.cuh")).read().split("\n")
        print("  warp-wide intrinsics per warp by source line (both passes):")
        for line, cnt in sorted(sites.items(), key=lambda kv: -kv[1])[:int(os.environ.get("WB_MODEL_SITES", "16"))]:
            print("    %7.1f  %4d: %s" % (cnt / m, line, src[line - 1].strip()[:100]))


if __name__ == "__main__":
    main()

"""Golden vectors at CONFIG scale (BASELINE.json configs[0], [1]): the oracle runs decode, canonical order, octree
leaves, tile scan and postscan over the WHOLE synthetic cloud the bench uses, then classifies a SAMPLE of its points —
each one against the whole cloud (wbo_classify_sel) — because classifying all 1e8 points takes the CPU hours.
    python tests/golden/make_config_scale.py SCENE POINTS [SAMPLE [STRIPS]]      e.g. 2 100000000 400000   (about 25 min)
STRIPS > 1: the cloud is the concatenation of that many x-strip files, generated exactly as bench.py --gpus STRIPS
generates them (BASELINE configs[2], the multi-tile scene); the vectors are then for the SHARDED run, input order =
strip after strip (file name ..._stripsW.npz).
Writes tests/golden/config/config_scale_s<SCENE>_<POINTS>.npz:
    sample      input indices of the sampled points (uint32): random singles plus whole runs of 2048 neighbours in
                canonical order, so that both the typical and the locally worst case are in it
    labels      the oracle's class bytes for them
    margin      how many of them had an in/out test within 1e-12 of the hyperboloid's surface
    dump_sha256, tiles_sha256, order_sha256   digests of the octree dump text, of the tile table (n, nPoints,
                treeFlags, bits of hyperboloidSize, ascending n) and of the canonical order (uint32 input indices)
    n_leaves, n_tiles, n_duplicates, hyp_median, hyp_max
tests/test_gpu_large.py compares the CUDA path's output on the same generated cloud with these."""
import ctypes as C
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wb_oracle as O  # noqa: E402
from wolkenbase_b200 import synth  # noqa: E402

PARAMS = dict(tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1)


def tiles_digest(tiles):
    h = hashlib.sha256()
    for f, dt in (("n", np.int32), ("nPoints", np.int32), ("treeFlags", np.int32)):
        h.update(np.ascontiguousarray(tiles[f].astype(dt)).tobytes())
    h.update(np.ascontiguousarray(tiles["hyperboloidSize"].astype(np.float64)).view(np.uint64).tobytes())
    return h.hexdigest()


def main():
    scene, n_points = int(sys.argv[1]), int(sys.argv[2])
    n_sample = int(sys.argv[3]) if len(sys.argv) > 3 else 400000
    strips = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    t0 = time.time()
    L = O.lib()
    if strips > 1:
        d = synth.describe(scene, n_points)
        clouds = []
        for r in range(strips):
            c0, c1 = d.grid_nx * r // strips, d.grid_nx * (r + 1) // strips
            clouds.append(synth.generate(scene, n_points, seed=scene, region=(c0, 0, c1 - c0, d.grid_ny),
                                         gps_base=d.grid_ny * c0))
    else:
        clouds = [synth.generate(scene, n_points, seed=scene)]
    cloud = clouds[0]
    assert all(c.scale == cloud.scale and c.offset == cloud.offset for c in clouds)
    n = sum(c.n for c in clouds)
    xyz = np.ascontiguousarray(np.concatenate([c.ints() for c in clouds]))
    for c in clouds:
        c.records = None
    pts = np.empty((n, 3), dtype=np.float64)
    L.wbo_coords(xyz.ctypes.data, n, O._d3(cloud.scale), O._d3(cloud.offset), 1.0, pts.ctypes.data)
    # identical locations: the first in input order keeps its place (octree.cpp:620-662)
    key = np.ascontiguousarray(xyz).view([("x", np.int32), ("y", np.int32), ("z", np.int32)]).reshape(-1)
    _, first = np.unique(key, return_index=True)
    n_dup = n - len(first)
    del key, xyz
    assert n_dup == 0, "scene with identical locations: extend this script with the representative map"
    corners = np.ascontiguousarray(np.array([k for c in clouds for k in (c.min_corner, c.max_corner)], dtype=np.float64))
    center, side, cube = (C.c_double * 3)(), C.c_double(), (C.c_double * 4)()
    L.wbo_size_fit(corners.ctypes.data, len(corners), center, C.byref(side))
    L.wbo_bbox_cube(corners.ctypes.data, len(corners), cube)
    keys = np.empty(n, dtype=np.uint64)
    order = np.empty(n, dtype=np.uint32)
    L.wbo_sort(pts.ctypes.data, n, center, side.value, keys.ctypes.data, order.ctypes.data)
    print("sorted %d points, %.0f s" % (n, time.time() - t0), flush=True)
    cap = max(16, n // 32 + 16)
    leaves = np.zeros(cap, dtype=O.LEAF_DTYPE)
    nl = L.wbo_leaves(keys.ctypes.data, n, center, side.value, leaves.ctypes.data, cap)
    assert nl <= cap
    buf = C.create_string_buffer(int(nl) * 120 + 64)
    ln = L.wbo_dump(leaves.ctypes.data, nl, buf, len(buf))
    dump_sha = hashlib.sha256(buf.raw[:ln]).hexdigest()
    del keys, buf, leaves
    order_sha = hashlib.sha256(order.tobytes()).hexdigest()
    pts = np.ascontiguousarray(pts[order])
    print("%d leaves, dump %d bytes, %.0f s" % (nl, ln, time.time() - t0), flush=True)
    capt = n // 4 + 1024
    tiles = np.zeros(capt, dtype=O.TILE_DTYPE)
    nt = L.wbo_scan(pts.ctypes.data, n, cube, PARAMS["tile_size"], PARAMS["min_hyperboloid_size"], tiles.ctypes.data, capt)
    assert 0 <= nt <= capt, nt
    tiles = tiles[:nt].copy()
    spacing, lo, hi = C.c_double(), C.c_int(), C.c_int()
    L.wbo_snake_set_size(cube[3], PARAMS["tile_size"], C.byref(spacing), C.byref(lo), C.byref(hi))
    L.wbo_postscan(tiles.ctypes.data, nt, spacing.value)
    tiles = tiles[np.argsort(tiles["n"], kind="stable")]
    print("%d tiles, spacing %.3f, %.0f s" % (nt, spacing.value, time.time() - t0), flush=True)
    rng = np.random.default_rng(scene * 1000 + 7)
    runs = max(1, n_sample // 4 // 2048)
    starts = rng.integers(0, n - 2048, runs)
    pos = np.concatenate([rng.choice(n, n_sample - runs * 2048, replace=False)] +
                         [np.arange(s, s + 2048) for s in starts]).astype(np.uint64)
    pos = np.unique(pos)
    lab = np.zeros(len(pos), dtype=np.uint8)
    margins = C.c_uint64()
    L.wbo_classify_sel(pts.ctypes.data, n, cube, PARAMS["tile_size"], PARAMS["max_slope"], PARAMS["thickness"],
                       tiles.ctypes.data, nt, pos.ctypes.data, len(pos), lab.ctypes.data, C.byref(margins))
    print("classified %d sampled points, %.0f s" % (len(pos), time.time() - t0), flush=True)
    hyp = tiles["hyperboloidSize"]
    out = os.path.join(ROOT, "tests", "golden", "config", "config_scale_s%d_%d%s.npz"
                       % (scene, n_points, "_strips%d" % strips if strips > 1 else ""))
    np.savez_compressed(out, scene=scene, n_points=n_points, n=n, strips=strips, counts=np.array([c.n for c in clouds]), sample=order[pos.astype(np.int64)], labels=lab,
                        margin=int(margins.value), dump_sha256=dump_sha, tiles_sha256=tiles_digest(tiles),
                        order_sha256=order_sha, n_leaves=int(nl), n_tiles=int(nt), n_duplicates=int(n_dup),
                        hyp_median=float(np.median(hyp)), hyp_max=float(hyp.max()), spacing=spacing.value)
    print(out, os.path.getsize(out), "bytes; labels", np.bincount(lab, minlength=3).tolist(), "margin", margins.value)


if __name__ == "__main__":
    main()

"""Pin the CPU restatement (oracle/wb_oracle.c) against the compiled, unmodified reference
(oracle/_ref/ref_driver, 1 thread, canonical order — SURVEY.md §8c).

Runs only where oracle/_ref exists (it is built from /root/reference by `make -C oracle ref`).
  python oracle/validate_against_ref.py [--golden]   # --golden also (re)writes tests/golden/*
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wb_oracle  # noqa: E402
from wolkenbase_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
REF_TILE = np.dtype([("n", "<i4"), ("ex", "<i4"), ("ey", "<i4"), ("nPoints", "<i4"), ("treeFlags", "<i4"),
                     ("nGround", "<i4"), ("density", "<f8"), ("hyperboloidSize", "<f8"), ("height", "<f8")])

CASES = [  # name, scene, n, seed, params
    ("street_20k", 1, 20000, 1, {}),
    ("aerial_60k", 2, 60000, 2, {}),
    ("urban_30k", 5, 30000, 5, {}),
    ("terrestrial_40k", 4, 40000, 4, {}),
    ("aerial_30k_thick", 2, 30000, 7, {"thickness": 0.05, "max_slope": 0.7, "tile_size": 2.0,
                                       "min_hyperboloid_size": 0.2}),
    # 400 records share their XYZ with another record: the reference keeps one point per location
    # (octree.cpp:620-662); the lost records read 255 in ref_labels
    ("aerial_20k_dups", 2, 20000, 77, {"dups": 400}),
    # LAS 1.4 format 6 (30 B records): the reference's other readPoint branch (las.cpp:759-781)
    ("multitile_fmt6_40k", 3, 40000, 3, {}),
    ("urban_40k_slope2", 5, 40000, 9, {"max_slope": 2.0, "thickness": 0.1}),
    ("street_30k_tile3", 1, 30000, 4, {"tile_size": 3.0}),
    ("terrestrial_30k_slope05", 4, 30000, 6, {"max_slope": 0.5, "min_hyperboloid_size": 1.0}),
    # round 2: a steep slope over small tiles, a thick layer under a large minimum hyperboloid, small tiles on format 6
    ("aerial_25k_steep_tile05", 2, 25000, 13, {"max_slope": 3.0, "tile_size": 0.5}),
    ("urban_20k_thick_bigmin", 5, 20000, 15, {"thickness": 0.3, "min_hyperboloid_size": 2.0}),
    ("multitile_fmt6_30k_tile07", 3, 30000, 21, {"tile_size": 0.7, "max_slope": 1.5}),
]


# Several input files at once (wolkencanvas.cpp:502-527 reads a list): each part is
#   {"scene", "n", "seed", "region": None | "left" | "right", "shift": [dx_ticks, dy_ticks], "gps_after": index of the
#    part whose point count its gpsTime continues from}
# "shift" re-expresses the same points with another header offset (las.cpp:808 uses each file's own).
MULTI_CASES = [
    ("multi_halves_shifted", [
        {"scene": 2, "n": 40000, "seed": 29, "region": "left", "shift": [0, 0]},
        {"scene": 2, "n": 40000, "seed": 29, "region": "right", "shift": [20000, -7000], "gps_after": 0}], {}),
    ("multi_fmt1_fmt6", [
        {"scene": 2, "n": 20000, "seed": 31, "region": None, "shift": [0, 0]},
        {"scene": 3, "n": 15000, "seed": 32, "region": None, "shift": [-3000, 5000], "gps_after": 0}], {}),
]


def clouds_from_parts(parts):
    """The synthetic files a MULTI_CASES entry names (shared with tests/test_oracle_golden.py)."""
    import ctypes as C
    clouds = []
    for part in parts:
        d = synth.describe(part["scene"], part["n"])
        region = None
        if part.get("region") == "left":
            region = (0, 0, d.grid_nx // 2, d.grid_ny)
        elif part.get("region") == "right":
            region = (d.grid_nx // 2, 0, d.grid_nx - d.grid_nx // 2, d.grid_ny)
        base = clouds[part["gps_after"]].n if "gps_after" in part else 0
        c = synth.generate(part["scene"], part["n"], seed=part["seed"], region=region, gps_base=base)
        dx, dy = part.get("shift", [0, 0])
        if dx or dy:
            recs = c.records.copy()
            ints = np.ascontiguousarray(recs[:, :12]).view(np.int32).reshape(-1, 3).copy()
            ints[:, 0] -= dx
            ints[:, 1] -= dy
            recs[:, :12] = ints.view(np.uint8).reshape(-1, 12)
            nd = synth.SynthDesc()
            for f, _ in synth.SynthDesc._fields_:
                setattr(nd, f, getattr(c.desc, f))
            nd.offset[0] = c.desc.offset[0] + c.desc.scale * dx
            nd.offset[1] = c.desc.offset[1] + c.desc.scale * dy
            bbox = c.bbox.copy()
            bbox[0] -= dx
            bbox[3] -= dx
            bbox[1] -= dy
            bbox[4] -= dy
            hdr = np.zeros(375, dtype=np.uint8)
            size = synth.lib().wb_synth_header(C.byref(nd), c.n, bbox.ctypes.data, hdr.ctypes.data)
            c = synth.Cloud(nd, hdr[:size].copy(), recs, bbox)
        clouds.append(c)
    return clouds


def compare_multi(name, parts, p, golden_dir=None):
    clouds = clouds_from_parts(parts)
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for i, c in enumerate(clouds):
            paths.append(os.path.join(td, "%s_%d.las" % (name, i)))
            c.write(paths[-1])
        info, rdump, rlabels, rtiles = run_ref(paths, os.path.join(td, "ref"), p)
    res = wb_oracle.run([wb_oracle.file_from_cloud(c) for c in clouds], **p)
    ot = res.tiles
    rep = {"case": name, "points": int(sum(c.n for c in clouds)),
           "root": list(res.root_center) == info["root_center"] and res.root_side == info["root_side"],
           "snake": res.snake_index == info["snake_index"] and res.spacing == info["spacing"],
           "dump_equal": res.dump == rdump, "tiles": [len(ot), len(rtiles)]}
    ok = rep["root"] and rep["snake"] and rep["dump_equal"] and len(ot) == len(rtiles)
    if ok:
        for fld in ("n", "ex", "ey", "nPoints", "treeFlags"):
            ok = ok and bool((ot[fld] == rtiles[fld]).all())
        for fld in ("density", "hyperboloidSize", "height"):
            ok = ok and bool((ot[fld].view(np.uint64) == rtiles[fld].view(np.uint64)).all())
    stored = rlabels != 255
    rep["duplicates"] = [int(res.n_duplicates), int(info["duplicates"]), int((~stored).sum())]
    ok = ok and len(set(rep["duplicates"])) == 1
    rep["label_mismatch"] = int((res.labels[stored] != rlabels[stored]).sum()) if len(rlabels) == len(res.labels) else -1
    ok = ok and rep["label_mismatch"] == 0
    rep["labels_hist"] = np.bincount(rlabels, minlength=3)[:3].tolist()
    rep["ok"] = bool(ok)
    if golden_dir and ok:
        np.savez_compressed(os.path.join(golden_dir, name + ".npz"), parts=json.dumps(parts), params=json.dumps(p),
                            ref_dump=np.frombuffer(rdump.encode("utf-8"), dtype=np.uint8),
                            ref_labels=rlabels, ref_tiles=rtiles,
                            ref_root=np.array(info["root_center"] + [info["root_side"]]),
                            ref_spacing=info["spacing"], ref_snake_index=info["snake_index"])
    return rep


def run_ref(las_paths, out_prefix, p):
    cmd = [REF, "-t", "1", "-c", "-o", out_prefix,
           "-T", repr(p.get("tile_size", 1.0)), "-S", repr(p.get("max_slope", 1.0)),
           "-K", repr(p.get("thickness", 0.0)), "-M", repr(p.get("min_hyperboloid_size", 0.1))] + las_paths
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.dirname(out_prefix))
    line = [l for l in out.stdout.splitlines() if l.lstrip().startswith("{") or "{\"points\"" in l][-1]
    info = json.loads(line[line.index("{"):])
    dump = open(out_prefix + ".dump", encoding="utf-8").read()
    labels = np.fromfile(out_prefix + ".labels", dtype=np.uint8)
    raw = open(out_prefix + ".tiles", "rb").read()
    tiles = np.frombuffer(raw[32:], dtype=REF_TILE)
    return info, dump, labels, tiles


def compare(name, scene, n, seed, p, golden_dir=None):
    cloud = synth.generate(scene, n, seed=seed)
    p = dict(p)
    dups = p.pop("dups", 0)
    if dups:
        cloud = synth.with_duplicates(cloud, dups, seed)
    with tempfile.TemporaryDirectory() as td:
        las = os.path.join(td, name + ".las")
        cloud.write(las)
        info, rdump, rlabels, rtiles = run_ref([las], os.path.join(td, "ref"), p)
    res = wb_oracle.run([wb_oracle.file_from_cloud(cloud)], **p)
    ok = True
    rep = {"case": name, "points": int(cloud.n)}
    rep["root"] = (list(res.root_center) == info["root_center"] and res.root_side == info["root_side"])
    rep["snake"] = (res.snake_index == info["snake_index"] and res.spacing == info["spacing"])
    rep["dump_equal"] = (res.dump == rdump)
    rep["leaves"] = len(res.leaves)
    ot = res.tiles
    rep["tiles"] = [len(ot), len(rtiles)]
    if len(ot) == len(rtiles):
        same_addr = bool((ot["n"] == rtiles["n"]).all() and (ot["ex"] == rtiles["ex"]).all() and (ot["ey"] == rtiles["ey"]).all())
        rep["tile_addr_equal"] = same_addr
        rep["tile_npoints_equal"] = bool((ot["nPoints"] == rtiles["nPoints"]).all())
        rep["tile_treeflags_equal"] = bool((ot["treeFlags"] == rtiles["treeFlags"]).all())
        for fld in ("density", "hyperboloidSize", "height"):
            a, b = ot[fld], rtiles[fld]
            rep["tile_%s_bitexact" % fld] = int((a.view(np.uint64) != b.view(np.uint64)).sum())
    else:
        ok = False
    stored = rlabels != 255                      # lost duplicates were never stored
    rep["duplicates"] = [int(res.n_duplicates), int(info["duplicates"]), int((~stored).sum())]
    ok = ok and len(set(rep["duplicates"])) == 1
    mism = int((res.labels[stored] != rlabels[stored]).sum())
    rep["label_mismatch"] = mism
    rep["labels_hist"] = np.bincount(rlabels, minlength=3)[:3].tolist()
    rep["margin_points"] = int(res.margin_count)
    for k in ("root", "snake", "dump_equal", "tile_addr_equal", "tile_npoints_equal", "tile_treeflags_equal"):
        ok = ok and bool(rep.get(k))
    for fld in ("density", "hyperboloidSize", "height"):
        ok = ok and rep.get("tile_%s_bitexact" % fld) == 0
    ok = ok and mism == 0
    rep["ok"] = ok
    if golden_dir and ok:
        os.makedirs(golden_dir, exist_ok=True)
        np.savez_compressed(os.path.join(golden_dir, name + ".npz"),
                            scene=scene, n=n, seed=seed, params=json.dumps(p), dups=dups,
                            ref_dump=np.frombuffer(rdump.encode("utf-8"), dtype=np.uint8),
                            ref_labels=rlabels, ref_tiles=rtiles,
                            ref_root=np.array(info["root_center"] + [info["root_side"]]),
                            ref_spacing=info["spacing"], ref_snake_index=info["snake_index"])
    return rep


if __name__ == "__main__":
    golden = os.path.join(ROOT, "tests", "golden") if "--golden" in sys.argv else None
    allok = True
    for c in CASES:
        r = compare(*c, golden_dir=golden)
        print(json.dumps(r))
        allok = allok and r["ok"]
    if golden:
        os.makedirs(os.path.join(golden, "multi"), exist_ok=True)
    for name, parts, p in MULTI_CASES:
        r = compare_multi(name, parts, p, golden_dir=os.path.join(golden, "multi") if golden else None)
        print(json.dumps(r))
        allok = allok and r["ok"]
    sys.exit(0 if allok else 1)

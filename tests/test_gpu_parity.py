"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed
compiled-reference fixtures.  Bit-exact for integer/byte/index work; labels may differ only on
points whose in/out decision lies within relative 1e-12 of the hyperboloid surface (counted)."""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import wb_oracle as O  # noqa: E402
from wolkenbase_b200 import api, synth  # noqa: E402

GOLDEN = sorted(os.path.basename(p)[:-4] for p in
                glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def _run_gpu(ctx, clouds, params):
    ctx.clear()
    ctx.set_params(**params)
    for c in clouds:
        ctx.add_extent(c.min_corner, c.max_corner)
    for c in clouds:
        ctx.add_las(c.records, c.fmt, c.scale, c.offset)
    ctx.run()
    n = sum(c.n for c in clouds)
    return n


def _check_against_oracle(ctx, clouds, params, res=None):
    n = _run_gpu(ctx, clouds, params)
    if res is None:
        res = O.run([O.file_from_cloud(c) for c in clouds], **params)
    g = ctx.geometry()
    assert list(g.root_center) == list(res.root_center) and g.root_side == res.root_side
    assert list(g.cube) == list(res.cube) and g.spacing == res.spacing
    assert (g.snake_lo, g.snake_hi) == (res.lo, res.hi)
    # K1 decode: bit-exact integers and class bytes
    x, y, z, cls = ctx.decoded(n)
    ints = np.concatenate([c.ints() for c in clouds])
    assert (x == ints[:, 0]).all() and (y == ints[:, 1]).all() and (z == ints[:, 2]).all()
    # K2/K3 order: bit-exact keys and permutation
    nv = ctx.stats()["n_points"]              # stored points: records minus dropped and lost ones
    assert nv == len(res.keys)
    order, keys = (a[:nv] for a in ctx.order(n))
    assert (keys == res.keys).all()
    assert (order == res.order_input).all()
    # K4 leaves: byte-identical dump
    assert ctx.dump() == res.dump
    lv = ctx.leaves()
    assert (lv["first"] == res.leaves["first"]).all() and (lv["count"] == res.leaves["count"]).all()
    # leaf z range = OctBuffer low/high
    zs = res.points_sorted[:, 2]
    for i in range(0, len(lv), max(1, len(lv) // 50)):
        a, b = int(lv["first"][i]), int(lv["first"][i]) + int(lv["count"][i])
        assert lv["low"][i] == zs[a:b].min() and lv["high"][i] == zs[a:b].max()
    # K7/K8 tiles
    t = ctx.tiles()
    assert len(t) == len(res.tiles)
    for f in ("n", "ex", "ey", "nPoints", "treeFlags"):
        assert (t[f] == res.tiles[f]).all(), f
    ulp = {}
    for f in ("density", "hyperboloidSize", "height"):
        d = np.abs(t[f].view(np.int64) - res.tiles[f].view(np.int64))
        ulp[f] = int((d > 0).sum())
        assert d.max() == 0, (f, int(d.max()))     # bit-identical so far on every scene: no tolerance to hide behind
    # K9 labels
    lab = ctx.labels(n)
    st = ctx.stats()
    mism = int((lab != res.labels).sum())
    assert mism <= st["n_margin"] + res.margin_count, (mism, st["n_margin"], res.margin_count)
    hist = ctx.count_classes()
    assert (hist[:3] == np.bincount(lab, minlength=3)[:3]).all()
    return {"mismatch": mism, "tile_ulp_diffs": ulp, "stats": st}


@pytest.mark.parametrize("scene,n,seed", [(2, 5000, 3), (1, 20000, 1), (2, 60000, 2), (5, 30000, 5), (4, 40000, 4)])
def test_pipeline_matches_oracle(ctx, scene, n, seed):
    cloud = synth.generate(scene, n, seed=seed)
    rep = _check_against_oracle(ctx, [cloud], {})
    assert rep["mismatch"] == 0


def test_identical_locations(ctx):
    """OctBuffer::put (octree.cpp:620-662) keeps one point per XYZ: the dump, the order, the tile
    counts and the labels follow the reference on clouds with pairs, longer chains and a file read twice."""
    base = synth.generate(2, 20000, seed=31)
    few = synth.with_duplicates(base, 300, 1)
    rep = _check_against_oracle(ctx, [few], {})
    assert rep["mismatch"] == 0 and 250 < rep["stats"]["n_duplicates"] <= 300
    many = synth.with_duplicates(synth.with_duplicates(base, 15000, 2), 15000, 3)      # chains of 3 and more
    rep = _check_against_oracle(ctx, [many], {})
    assert rep["mismatch"] == 0 and rep["stats"]["n_duplicates"] > 9000         # multiplicities up to 10
    assert rep["stats"]["n_points"] + rep["stats"]["n_duplicates"] == base.n
    # the same file twice: every point of the second copy is lost, and gets the first copy's class
    rep = _check_against_oracle(ctx, [base, base], {})
    assert rep["mismatch"] == 0 and rep["stats"]["n_duplicates"] == base.n
    lab = ctx.labels(2 * base.n)
    assert (lab[:base.n] == lab[base.n:]).all()
    # and the context is clean afterwards
    rep = _check_against_oracle(ctx, [base], {})
    assert rep["mismatch"] == 0 and rep["stats"]["n_duplicates"] == 0


def test_nondefault_params(ctx):
    cloud = synth.generate(2, 30000, seed=7)
    p = {"thickness": 0.05, "max_slope": 0.7, "tile_size": 2.0, "min_hyperboloid_size": 0.2}
    rep = _check_against_oracle(ctx, [cloud], p)
    assert rep["mismatch"] == 0


@pytest.mark.parametrize("case", GOLDEN)
def test_matches_compiled_reference_fixture(ctx, case, golden_dir):
    """Against outputs of the UNMODIFIED reference (oracle/_ref, generated in the build container)."""
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    p = json.loads(str(g["params"]))
    cloud = synth.generate(int(g["scene"]), int(g["n"]), seed=int(g["seed"]))
    dups = int(g["dups"]) if "dups" in g else 0
    if dups:
        cloud = synth.with_duplicates(cloud, dups, int(g["seed"]))
    n = _run_gpu(ctx, [cloud], p)
    assert ctx.dump() == bytes(g["ref_dump"]).decode("utf-8")
    t, rt = ctx.tiles(), g["ref_tiles"]
    assert len(t) == len(rt)
    for f in ("n", "ex", "ey", "nPoints", "treeFlags"):
        assert (t[f] == rt[f]).all(), f
    assert np.abs(t["hyperboloidSize"].view(np.int64) - rt["hyperboloidSize"].view(np.int64)).max() <= 4
    lab = ctx.labels(n)
    stored = g["ref_labels"] != 255          # records the reference lost to an identical location
    assert int((~stored).sum()) == ctx.stats()["n_duplicates"]
    mism = int((lab[stored] != g["ref_labels"][stored]).sum())
    assert mism <= ctx.stats()["n_margin"], mism


def test_formats_and_multifile(ctx):
    """Formats 6 (30 B) and 3 (34 B) decode; two files with different offsets merge into one cloud."""
    a = synth.generate(3, 12000, seed=21)                      # format 6
    b = synth.generate(5, 9000, seed=22)                       # format 3
    assert a.fmt == 6 and b.fmt == 3
    rep = _check_against_oracle(ctx, [a], {})
    assert rep["mismatch"] == 0
    d = synth.describe(2, 40000)
    half = d.grid_nx // 2
    left = synth.generate(2, 40000, seed=9, region=(0, 0, half, d.grid_ny))
    right = synth.generate(2, 40000, seed=9, region=(half, 0, d.grid_nx - half, d.grid_ny), gps_base=left.n)
    rep = _check_against_oracle(ctx, [left, right], {})
    assert rep["mismatch"] == 0


def test_ragged_and_tiny(ctx):
    """Point counts around the bucket capacity 537 and the chunk size 32 (testsplitfile's primes,
    wolkentest.cpp:948-981), and a 1-point cloud."""
    for n in (1, 31, 33, 523, 541, 4297, 4327):
        d = synth.describe(2, 10000)
        cloud = synth.generate(2, 10000, seed=n)
        recs = np.ascontiguousarray(cloud.records[:n])
        sub = synth.Cloud(cloud.desc, cloud.header, recs, cloud.bbox)     # header extents of the full scene
        rep = _check_against_oracle(ctx, [sub], {})
        assert rep["mismatch"] == 0


def test_return_number_zero_rule(ctx):
    """threads.cpp:485-530: if record 0 has a return number, records with return number 0 are
    dropped; otherwise they are kept."""
    cloud = synth.generate(2, 8000, seed=13)
    recs = cloud.records.copy()
    recs[5::7, 14] &= 0xf8                                       # return number 0 on every 7th record
    sub = synth.Cloud(cloud.desc, cloud.header, recs, cloud.bbox)
    res = O.run([O.file_from_cloud(sub)])
    n = _run_gpu(ctx, [sub], {})
    st = ctx.stats()
    keep = sub.records[:, 14] & 7 != 0
    assert st["n_dropped"] == int((~keep).sum()) and st["n_points"] == int(keep.sum())
    lab = ctx.labels(n)
    assert (lab[keep] == res.labels).all()
    assert (lab[~keep] == 0).all()                               # untouched class byte
    assert ctx.dump() == res.dump
    recs2 = cloud.records.copy()
    recs2[:, 14] &= 0xf8                                         # all zero: kept, treated as return 1
    sub2 = synth.Cloud(cloud.desc, cloud.header, recs2, cloud.bbox)
    n = _run_gpu(ctx, [sub2], {})
    assert ctx.stats()["n_dropped"] == 0
    # a caller that embuffers every point itself (wolkencli.cpp:104-108) keeps them all
    ctx.set_return_zero_rule(True)
    try:
        n = _run_gpu(ctx, [sub], {})
        st = ctx.stats()
        assert st["n_dropped"] == 0 and st["n_points"] == sub.n
        assert (ctx.labels(n) == O.run([O.file_from_cloud(cloud)]).labels).all()   # same XYZ as the untouched cloud
    finally:
        ctx.set_return_zero_rule(False)


def test_injected_tile_table(ctx):
    """Classification parity can be judged independently of scan parity: inject the oracle's tiles."""
    cloud = synth.generate(2, 20000, seed=17)
    res = O.run([O.file_from_cloud(cloud)])
    ctx.clear()
    ctx.set_params()
    ctx.add_cloud(cloud)
    ctx.build()
    ctx.scan()
    t = np.zeros(len(res.tiles), dtype=api.TILE_DTYPE)
    for f in ("n", "ex", "ey", "nPoints", "treeFlags", "density", "hyperboloidSize", "height"):
        t[f] = res.tiles[f]
    ctx.set_tiles(t)
    ctx.classify()
    assert (ctx.labels(cloud.n) == res.labels).all()


def test_patch_records(ctx):
    cloud = synth.generate(2, 6000, seed=19)
    _run_gpu(ctx, [cloud], {})
    lab = ctx.labels(cloud.n)
    recs = cloud.records.copy()
    ctx.patch_records(recs, cloud.fmt)
    assert ((recs[:, 15] & 31) == lab).all()
    other = np.delete(np.arange(recs.shape[1]), 15)
    assert (recs[:, other] == cloud.records[:, other]).all()


def test_size_independent_properties(ctx):
    """At a size the CPU oracle would not finish quickly: order is a permutation sorted by key,
    leaves tile the array, bucket sizes respect the capacity, class counts add up, and the run
    is reproducible."""
    cloud = synth.generate(2, 2_000_000, seed=23)
    n = _run_gpu(ctx, [cloud], {})
    order, keys = ctx.order(n)
    assert (np.diff(keys.astype(np.int64)) >= 0).all()
    assert (np.sort(order) == np.arange(n, dtype=np.uint32)).all()
    eq = keys[1:] == keys[:-1]
    assert (order[1:][eq] > order[:-1][eq]).all()                # ties broken by input index
    lv = ctx.leaves()
    assert lv["first"][0] == 0 and (lv["first"][1:] == lv["first"][:-1] + lv["count"][:-1]).all()
    assert int(lv["count"].sum()) == n and lv["count"].max() <= 537
    # a leaf's parent cube holds more than 537 points: merge siblings and count
    lab1 = ctx.labels(n).copy()
    hist = ctx.count_classes()
    assert int(hist.sum()) == n and hist[1] + hist[2] == n
    t = ctx.tiles()
    assert int(t["nPoints"].sum()) == ctx.stats()["n_memberships"]
    n = _run_gpu(ctx, [cloud], {})
    assert (ctx.labels(n) == lab1).all()
    # some, not all, of the scene is non-ground (the share depends on the tile spacing the snake picks)
    frac = hist[1] / n
    assert 0.05 < frac < 0.6


@pytest.mark.parametrize("scene,n", [(1, 300000), (4, 300000), (5, 300000), (2, 400000)])
def test_baseline_scenes_at_scale(ctx, scene, n):
    """The four BASELINE scene models at a few hundred thousand points (deeper octrees, skewed
    leaf occupancy for the terrestrial scene, stacked walls for the urban one): full parity."""
    cloud = synth.generate(scene, n, seed=scene)
    rep = _check_against_oracle(ctx, [cloud], {})
    assert rep["mismatch"] == 0


def test_device_math_known_answers(ctx):
    """atan2i, hypot and the sector binning on the device against the oracle / libm."""
    import ctypes
    rng = np.random.default_rng(7)
    n = 400000
    x = rng.normal(size=n) * 10.0 ** rng.integers(-3, 3, size=n)
    y = rng.normal(size=n) * 10.0 ** rng.integers(-3, 3, size=n)
    # exact axes, diagonals, sector edges and the origin
    special = np.array([[1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1], [0, 0], [3, 4],
                        [1e-300, 1], [1, 1e-300], [-2.5, 1e-17]], dtype=np.float64)
    x[:len(special)] = special[:, 0]
    y[:len(special)] = special[:, 1]
    a, h, s = ctx.test_math(y, x)
    L = O.lib()
    L.wbo_fill_tan_tables()
    want_a = np.array([L.wbo_atan2i(float(yy), float(xx)) for yy, xx in zip(y[:60000], x[:60000])], dtype=np.int32)
    assert (a[:60000] == want_a).all(), int((a[:60000] != want_a).sum())
    libm = ctypes.CDLL("libm.so.6")
    libm.hypot.restype = ctypes.c_double
    libm.hypot.argtypes = [ctypes.c_double, ctypes.c_double]
    want_h = np.array([libm.hypot(float(xx), float(yy)) for xx, yy in zip(x[:60000], y[:60000])])
    assert (h[:60000].view(np.uint64) == want_h.view(np.uint64)).all()
    # testintegertrig (wolkentest.cpp:122-132): atan2i(cossin(i)) folds back to i
    ang = rng.integers(-2**31 + 100000, 2**31 - 100000, size=20000)
    pi = np.arctan(np.longdouble(1)) * 4
    th = ang.astype(np.longdouble) * pi / np.longdouble(1073741824.)
    a2, _, _ = ctx.test_math(np.sin(th).astype(np.float64), np.cos(th).astype(np.float64))

    def fold(v):
        v = v.astype(np.int64) & 0xffffffff
        return np.where(((v >> 30) % 3) != 0, v ^ 0x80000000, v)

    assert (fold(a2) == fold(ang)).all()
    # sectors: where the fast binning answers, it agrees with atan2i's own sector
    u = a.astype(np.int64) & 0x7fffffff
    ok = s >= 0
    assert ok.mean() > 0.99
    assert ((u[ok] >> 25) == s[ok]).all()


def test_error_behaviour(ctx):
    """Every entry point reports instead of crashing: wrong phase order, unsupported formats,
    missing extents (the reference asserts or prints; SURVEY.md §8b 'Error conventions')."""
    cloud = synth.generate(2, 3000, seed=3)
    ctx.clear()
    ctx.set_params()
    with pytest.raises(api.WolkenError, match="no points"):
        ctx.build()
    with pytest.raises(api.WolkenError, match="waveform|not supported"):
        ctx.add_las(np.zeros((10, 57), dtype=np.uint8), 4, cloud.scale, cloud.offset)
    with pytest.raises(api.WolkenError, match="shorter"):
        ctx.add_las(np.zeros((10, 20), dtype=np.uint8), 1, cloud.scale, cloud.offset)
    with pytest.raises(api.WolkenError, match="unknown"):
        ctx.add_las(np.zeros((10, 28), dtype=np.uint8), 11, cloud.scale, cloud.offset)
    ctx.add_las(cloud.records, cloud.fmt, cloud.scale, cloud.offset)
    with pytest.raises(api.WolkenError, match="extent"):
        ctx.build()                                               # no header corners given
    ctx.add_extent(cloud.min_corner, cloud.max_corner)
    with pytest.raises(api.WolkenError, match="not built"):
        ctx.scan()
    ctx.build()
    with pytest.raises(api.WolkenError, match="postscan"):
        ctx.classify()
    with pytest.raises(api.WolkenError, match="not classified"):
        ctx.labels(cloud.n)
    with pytest.raises(api.WolkenError, match="already built"):
        ctx.add_las(cloud.records, cloud.fmt, cloud.scale, cloud.offset)
    ctx.scan()
    ctx.postscan()
    with pytest.raises(api.WolkenError, match="once"):
        ctx.postscan()
    ctx.classify()
    assert ctx.labels(cloud.n).max() <= 2
    with pytest.raises(api.WolkenError):
        ctx.set_params(tile_size=-1.0)


def _shifted(cloud, dx_ticks, dy_ticks, new_scale=None):
    """The same points expressed with another LAS offset (and optionally a coarser-looking scale
    factor 1: ints divided is not possible in general, so only the offset moves)."""
    recs = cloud.records.copy()
    ints = np.ascontiguousarray(recs[:, :12]).view(np.int32).reshape(-1, 3).copy()
    ints[:, 0] -= dx_ticks
    ints[:, 1] -= dy_ticks
    recs[:, :12] = ints.view(np.uint8).reshape(-1, 12)
    d = synth.SynthDesc()
    for f, _ in synth.SynthDesc._fields_:
        setattr(d, f, getattr(cloud.desc, f))
    d.offset[0] = cloud.desc.offset[0] + cloud.desc.scale * dx_ticks
    d.offset[1] = cloud.desc.offset[1] + cloud.desc.scale * dy_ticks
    bbox = cloud.bbox.copy()
    bbox[0] -= dx_ticks; bbox[3] -= dx_ticks
    bbox[1] -= dy_ticks; bbox[4] -= dy_ticks
    return synth.Cloud(d, cloud.header, recs, bbox)


def test_files_with_different_offsets_and_units(ctx):
    """Per-file scale/offset (las.cpp:808 uses each header's own) and a length unit other than 1
    (LasHeader::setUnit; mainwindow.cpp:398-411 'lengthUnit')."""
    d = synth.describe(2, 50000)
    half = d.grid_nx // 2
    left = synth.generate(2, 50000, seed=29, region=(0, 0, half, d.grid_ny))
    right = synth.generate(2, 50000, seed=29, region=(half, 0, d.grid_nx - half, d.grid_ny), gps_base=left.n)
    right2 = _shifted(right, 20000, -7000)                          # same ground, different header offset
    rep = _check_against_oracle(ctx, [left, right2], {})
    assert rep["mismatch"] == 0
    # unit: international foot
    cloud = synth.generate(2, 20000, seed=30)
    unit = 0.3048
    ctx.clear()
    ctx.set_params()
    ctx.add_extent([c * unit for c in cloud.min_corner], [c * unit for c in cloud.max_corner])
    ctx.add_las(cloud.records, cloud.fmt, cloud.scale, cloud.offset, unit)
    ctx.run()
    res = O.run([O.file_from_cloud(cloud)], unit=unit)
    assert ctx.dump() == res.dump
    assert (ctx.labels(cloud.n) == res.labels).all()


def test_encode_same_layout_returns_the_input_records(ctx, tmp_path):
    """wb_encode (LasHeader::readPoint -> writePoint, las.cpp:735-904) with the input's own format,
    scale and offset must give back the input records in canonical order with only the class
    replaced; the header figures are the totals, per-return counts and integer extremes."""
    cloud = synth.generate(5, 30000, seed=71)                      # format 3: gps + RGB
    rng = np.random.default_rng(5)
    recs = cloud.records.copy()
    recs[:, 12:14] = rng.integers(0, 256, (cloud.n, 2), dtype=np.uint8)           # intensity
    recs[:, 14] = rng.integers(1, 8, cloud.n, dtype=np.uint8) | (rng.integers(0, 32, cloud.n, dtype=np.uint8) << 3)
    recs[:, 16:20] = rng.integers(0, 256, (cloud.n, 4), dtype=np.uint8)           # angle, user, source
    recs[:, 16] = rng.integers(-90, 91, cloud.n).astype(np.int8).view(np.uint8)   # angles that survive deg->bin->deg
    ctx.clear()
    ctx.keep_records(True)
    ctx.set_params()
    ctx.add_extent(cloud.min_corner, cloud.max_corner)
    ctx.add_las(recs, cloud.fmt, cloud.scale, cloud.offset)
    ctx.run()
    n = cloud.n
    lab = ctx.labels(n)
    order, _ = ctx.order(n)
    cnt = ctx.leaf_class_counts()                                   # one slot: every class
    lv = ctx.leaves()
    assert (cnt[:, 0] == lv["count"]).all()
    L = cloud.rec_len
    dest = lv["first"].astype(np.uint64) * L
    out, st = ctx.encode(cloud.fmt, L, cloud.scale, cloud.offset, dest, np.zeros(len(lv), np.uint32), 1, n * L)
    want = recs[order].copy()
    want[:, 15] = (want[:, 15] & 0xe0) | (lab[order] & 31)
    got = out.reshape(n, L)
    assert (got == want).all(), np.nonzero((got != want).any(axis=0))[0]
    # the same records streamed from device memory into a file, at an offset, in two spans
    _, st2 = ctx.encode(cloud.fmt, L, cloud.scale, cloud.offset, dest, np.zeros(len(lv), np.uint32), 1, n * L,
                        fetch=False)
    assert st2 == st
    with open(str(tmp_path / "spans.bin"), "wb") as f:
        f.write(b"H" * 375)
        f.flush()
        half = (n // 2) * L
        ctx.write_encoded(f.fileno(), 375 + half, half, n * L - half)
        ctx.write_encoded(f.fileno(), 375, 0, half)
    raw = np.fromfile(str(tmp_path / "spans.bin"), dtype=np.uint8)
    assert len(raw) == 375 + n * L and (raw[:375] == ord("H")).all() and (raw[375:] == out).all()
    with pytest.raises(api.WolkenError):
        ctx.write_encoded(0, 0, n * L - 10, 100)
    ints = cloud.ints()
    assert st[0]["n_points"][0] == n and st[0]["imin"] == ints.min(axis=0).tolist() and st[0]["imax"] == ints.max(axis=0).tolist()
    by_ret = np.bincount(recs[:, 14] & 7, minlength=16)
    assert st[0]["n_points"][1:8] == by_ret[1:8].tolist()
    # per class: one slot and one file per class present
    classes = np.unique(lab).tolist()
    K = len(classes)
    cnt = ctx.leaf_class_counts(classes)
    assert (cnt.sum(axis=1) == lv["count"]).all()
    assert cnt.sum(axis=0).tolist() == [int((lab == c).sum()) for c in classes]
    start = np.zeros_like(cnt, dtype=np.uint64)
    start[1:] = np.cumsum(cnt, axis=0)[:-1]
    tot = cnt.sum(axis=0).astype(np.uint64)
    base = (np.concatenate([[0], np.cumsum(tot)[:-1]]) * L).astype(np.uint64)
    dest = base[None, :] + start * L
    file_of = np.tile(np.arange(K, dtype=np.uint32), (len(lv), 1))
    out, st = ctx.encode(cloud.fmt, L, cloud.scale, cloud.offset, dest, file_of, K, n * L, classes=classes)
    got = out.reshape(n, L)
    lo = lab[order]
    for k, c in enumerate(classes):
        a = int(base[k]) // L
        assert (got[a:a + int(tot[k])] == want[lo == c]).all(), c
    assert [s["n_points"][0] for s in st] == tot.tolist()
    ctx.clear()
    ctx.keep_records(False)
    # without the records the call must refuse, not invent them
    ctx.add_extent(cloud.min_corner, cloud.max_corner)
    ctx.add_las(recs, cloud.fmt, cloud.scale, cloud.offset)
    ctx.run()
    with pytest.raises(api.WolkenError):
        ctx.encode(cloud.fmt, L, cloud.scale, cloud.offset, lv["first"].astype(np.uint64) * L,
                   np.zeros(len(lv), np.uint32), 1, n * L)


def test_add_las_file_equals_add_las(ctx, tmp_path):
    """The file reader pipeline (pread threads -> pinned ring -> H2D -> decode) delivers exactly what
    wb_add_las gets from memory: several chunks, a ragged last one, kept records, and a file that
    is shorter than its header says is an error, not a partial cloud."""
    cloud = synth.generate(2, 2_600_000, seed=81)                  # 3 chunks of 1 Mi records
    n = cloud.n
    path = str(tmp_path / "big.las")
    cloud.write(path)
    hdr = len(cloud.header.tobytes())
    _run_gpu(ctx, [cloud], {})
    want_dec = [a.copy() for a in ctx.decoded(n)]
    want_lab = ctx.labels(n).copy()
    for keep in (False, True):
        ctx.clear()
        ctx.keep_records(keep)
        ctx.set_params()
        ctx.add_extent(cloud.min_corner, cloud.max_corner)
        ctx.add_las_file(path, hdr, n, cloud.fmt, cloud.rec_len, cloud.scale, cloud.offset)
        got = ctx.decoded(n)
        for a, b in zip(got, want_dec):
            assert (a == b).all()
        ctx.run()
        assert (ctx.labels(n) == want_lab).all()
    ctx.clear()
    ctx.keep_records(False)
    ctx.add_extent(cloud.min_corner, cloud.max_corner)
    with pytest.raises(api.WolkenError):
        ctx.add_las_file(path, hdr, n + 1000, cloud.fmt, cloud.rec_len, cloud.scale, cloud.offset)
    with pytest.raises(api.WolkenError):
        ctx.add_las_file(str(tmp_path / "missing.las"), hdr, n, cloud.fmt, cloud.rec_len, cloud.scale, cloud.offset)
    ctx.clear()
    # a small file after a big one reuses the ring
    small = synth.generate(1, 5000, seed=82)
    p2 = str(tmp_path / "small.las")
    small.write(p2)
    ctx.set_params()
    ctx.add_extent(small.min_corner, small.max_corner)
    ctx.add_las_file(p2, len(small.header.tobytes()), small.n, small.fmt, small.rec_len, small.scale, small.offset)
    ctx.run()
    res = O.run([O.file_from_cloud(small)])
    assert (ctx.labels(small.n) == res.labels).all()


def test_store_queries_on_device(ctx):
    """OctStore::countPointsIn / hiLoPointsIn / pointsIn (octree.cpp:1214-1293) on the device for
    every shape of shape.cpp against the oracle's Shape::in over the canonical-order points:
    same counts, same z range, same points in the same order — including shapes that miss, that
    swallow the cloud, that open upward, and degenerate ones."""
    cloud = synth.generate(5, 40000, seed=91)
    n = _run_gpu(ctx, [cloud], {})
    x, y, z = ctx.points_sorted(n)
    pts = np.stack([x, y, z], axis=1)
    order, _ = ctx.order(n)
    rng = np.random.default_rng(7)
    c = pts[rng.integers(0, n, 64)]                                # centres on the cloud
    ext = pts.max(axis=0) - pts.min(axis=0)
    cases = []
    for i in range(12):
        p = c[i]
        cases.append((api.SPHERE, [p[0], p[1], p[2], [0.0, 0.3, 2.0, 15.0, 1e4][i % 5]]))
        cases.append((api.PARABOLOID, [p[0], p[1], p[2] + [0.0, 1.0, 5.0][i % 3], [0.5, 13.0, -2.0, 0.0][i % 4]]))
        cases.append((api.HYPERBOLOID, [p[0], p[1], p[2] + [0.0, 2.0][i % 2], [0.1, 0.5, 20.0][i % 3], [1.0, 0.3, 2.0, -1.0][i % 4]]))
        cases.append((api.CYLINDER, [p[0], p[1], [0.0, 0.58, 7.0, 1e4][i % 4]]))
        cases.append((api.COLUMN, [p[0], p[1], [0.01, 1.0, 30.0][i % 3]]))
    cases.append((api.SPHERE, [pts[:, 0].min() - 500, pts[:, 1].min() - 500, 0, 10]))       # misses everything
    cases.append((api.CYLINDER, [pts[:, 0].mean(), pts[:, 1].mean(), float(ext[:2].max())]))  # swallows everything
    shp = np.concatenate([api.shapes(k, p) for k, p in cases])
    count, lo, hi = ctx.query_batch(shp)
    hits = 0
    for i, (k, p) in enumerate(cases):
        m = O.shape_filter(k, p, pts)
        assert int(count[i]) == int(m.sum()), (i, k, p)
        if m.any():
            assert lo[i] == z[m].min() and hi[i] == z[m].max(), (i, k, p)
            hits += 1
        else:
            assert lo[i] == np.inf and hi[i] == -np.inf
        total, pos, idx, got = ctx.query_points(shp[i:i + 1])
        assert total == int(m.sum()) and (pos == np.nonzero(m)[0]).all() and (idx == order[m]).all()
        assert (got == pts[m]).all()
    assert hits > 40 and int(count[-1]) == n and int(count[-2]) == 0
    # a raster of columns (WolkenCanvas::pixelColorRead for every pixel) in one call
    gx, gy = np.meshgrid(np.linspace(pts[:, 0].min(), pts[:, 0].max(), 64), np.linspace(pts[:, 1].min(), pts[:, 1].max(), 48))
    side = 2.5
    cols = api.shapes(api.COLUMN, np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, side)], axis=1))
    count, lo, hi = ctx.query_batch(cols)
    for i in rng.integers(0, len(cols), 60):
        m = O.shape_filter(api.COLUMN, [gx.ravel()[i], gy.ravel()[i], side], pts)
        assert int(count[i]) == int(m.sum())
        if m.any():
            assert lo[i] == z[m].min() and hi[i] == z[m].max()
    assert int(count.sum()) > n // 2
    # capped output reports the full count
    total, pos, idx, got = ctx.query_points(shp[-1:], cap=100)
    assert total == n and len(pos) == 100 and (pos == np.arange(100)).all()
    with pytest.raises(api.WolkenError):
        ctx.query_batch(api.shapes(9, [0, 0, 0]))


def _census(cloud, records=None):
    ctx = api.Context(0)
    ctx.keep_records()
    ctx.set_params()
    ctx.add_extent(cloud.min_corner, cloud.max_corner)
    ctx.add_las(cloud.records if records is None else np.ascontiguousarray(records), cloud.fmt, cloud.scale, cloud.offset)
    ctx.build()
    out = ctx.census(cap=16)
    ctx.close()
    return out


def test_census_of_the_store():
    """censusPoints (testpattern.cpp:56-123): test data carries the point number as GPS time; every stored point sets
    one bit; numbers seen twice and numbers missing below the highest one are reported."""
    cloud = synth.generate(2, 20000, seed=7)                     # format 1: GPS time at byte 20
    n = cloud.n
    c = _census(cloud)
    assert (c["status"], c["n_stored"], c["max_point"], c["n_missing"]) == (0, n, n, 0)
    # records removed from the file are missing from the store
    gone = np.array([5, 77, 1000, n - 2])
    c = _census(cloud, np.delete(cloud.records, gone, axis=0))
    assert (c["status"], c["max_point"], c["n_missing"], c["missing"]) == (0, n, 4, gone.tolist())
    # a record at the XYZ of an earlier one replaces it in the store: the earlier number is lost
    recs = cloud.records.copy()
    recs[900, :12] = recs[17, :12]
    recs[901, :12] = recs[17, :12]
    c = _census(cloud, recs)
    assert (c["status"], c["n_stored"], c["n_missing"], c["missing"]) == (0, n - 2, 2, [17, 900])
    # the same number twice
    recs = cloud.records.copy()
    recs[40, 20:28] = recs[41, 20:28]
    c = _census(cloud, recs)
    assert (c["status"], c["n_duplicate"], c["missing"]) == (1, 1, [40])
    # not test data
    recs = cloud.records.copy()
    recs[3, 20:28] = np.frombuffer(np.float64(12.5).tobytes(), dtype=np.uint8)
    assert _census(cloud, recs)["status"] == -1


def test_census_needs_the_records():
    cloud = synth.generate(2, 5000, seed=8)
    ctx = api.Context(0)
    ctx.set_params()
    ctx.add_cloud(cloud)
    ctx.build()
    with pytest.raises(api.WolkenError, match="not kept"):
        ctx.census()
    ctx.close()

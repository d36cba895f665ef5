"""The reference-shaped C++ surface (wolkenbase_b200/host) end to end on the GPU: wolkencli's
dump and classified LAS output against the oracle, and the OctStore query API against brute force."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import wb_oracle as O  # noqa: E402
from wolkenbase_b200 import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "wolkenbase_b200", "host", "wolkencli")
QUERY = os.path.join(ROOT, "wolkenbase_b200", "host", "wolkenquery")


def _read_las(path):
    raw = np.fromfile(path, dtype=np.uint8)
    hs = int(raw[94:96].view("<u2")[0])
    off = int(raw[96:100].view("<u4")[0])
    fmt = int(raw[104])
    ln = int(raw[105:107].view("<u2")[0])
    n = int(raw[107:111].view("<u4")[0]) if fmt < 6 else int(raw[247:255].view("<u8")[0])
    recs = raw[off:off + n * ln].reshape(n, ln)
    return fmt, recs, raw[:hs]


def test_cli_reference_mode_dump(tmp_path):
    """Without -o the CLI does what the reference's does: header lines, cube, octree dump."""
    cloud = synth.generate(2, 20000, seed=41)
    las = str(tmp_path / "in.las")
    cloud.write(las)
    out = subprocess.run([CLI, "--dump", str(tmp_path / "dumpfile"), las], capture_output=True, text=True, check=True)
    res = O.run([O.file_from_cloud(cloud)], classify=False)
    assert "Version 1.2 %d points, format 1" % cloud.n in out.stdout
    assert "All points in octree" in out.stdout and "Dumping octree" in out.stdout
    cx, cy, cz = (O.ldecimal(v) for v in res.root_center)
    assert "(%s,%s,%s)±%g" % (cx, cy, cz, res.root_side) in out.stdout
    assert open(tmp_path / "dumpfile", encoding="utf-8").read() == res.dump


def test_cli_embuffer_flow_equals_act_read(tmp_path):
    """The reference's wolkencli feeds the pool point by point (readPoint + embufferPoint, wolkencli.cpp:104-108);
    the shim turns those calls into runs of records for the device.  Two files, one with zero return numbers
    (which this flow keeps: nothing filters the points the caller embuffers)."""
    a = synth.generate(2, 12000, seed=43)
    b = synth.generate(2, 9000, seed=44)
    rb = b.records.copy()
    rb[3::5, 14] &= 0xf8
    b = synth.Cloud(b.desc, b.header, rb, b.bbox)
    la, lb = str(tmp_path / "a.las"), str(tmp_path / "b.las")
    a.write(la)
    b.write(lb)
    out = subprocess.run([CLI, "--embuffer", "-o", str(tmp_path / "o"), "--lossless", "--separate-classes", "0",
                          "--dump", str(tmp_path / "d"), la, lb], capture_output=True, text=True, check=True)
    assert "%d points, %d points in buffer" % (a.n, a.n) in out.stdout
    assert "%d points, %d points in buffer" % (b.n, a.n + b.n) in out.stdout
    keep_all = synth.Cloud(b.desc, b.header, synth.generate(2, 9000, seed=44).records, b.bbox)
    res = O.run([O.file_from_cloud(a), O.file_from_cloud(keep_all)])
    assert open(tmp_path / "d", encoding="utf-8").read() == res.dump
    fmt, recs, _ = _read_las(str(tmp_path / "o.las"))
    assert recs.shape[0] == a.n + b.n
    assert ((recs[:, 15] & 31) == res.labels).all()


def test_cli_classify_and_write(tmp_path):
    cloud = synth.generate(5, 30000, seed=5)                     # format 3 (34 B), config 5's scene
    las = str(tmp_path / "urban.las")
    cloud.write(las)
    res = O.run([O.file_from_cloud(cloud)])
    # one file, class byte in place (config 1 style)
    subprocess.run([CLI, "-o", str(tmp_path / "one"), "--lossless", "--separate-classes", "0", "--dump",
                    str(tmp_path / "d1"), las], capture_output=True, text=True, check=True)
    fmt, recs, hdr = _read_las(str(tmp_path / "one.las"))
    assert fmt == 3 and recs.shape == cloud.records.shape
    assert ((recs[:, 15] & 31) == res.labels).all()
    other = np.delete(np.arange(recs.shape[1]), 15)
    assert (recs[:, other] == cloud.records[:, other]).all()
    assert bytes(hdr[26:38]) == b"MODIFICATION"
    # split ground / nonground with a points-per-file limit (config 5 style)
    r = subprocess.run([CLI, "-o", str(tmp_path / "split.las"), "--lossless", "--points-per-file", "10000", las],
                       capture_output=True, text=True, check=True, cwd=str(tmp_path))
    n_ground, n_non = int((res.labels == 2).sum()), int((res.labels == 1).sum())
    assert "2 %d" % n_ground in r.stdout and "1 %d" % n_non in r.stdout
    got = {1: [], 2: []}
    for name in sorted(os.listdir(tmp_path)):
        if name.startswith("split-"):
            cls = 2 if "-ground-" in name else 1
            f, rr, _ = _read_las(str(tmp_path / name))
            assert rr.shape[0] <= 10000 and ((rr[:, 15] & 31) == cls).all()
            got[cls].append(rr)
    for cls, want_n in ((1, n_non), (2, n_ground)):
        allr = np.concatenate(got[cls])
        assert allr.shape[0] == want_n
        # multiset of gpsTime equals the reference's selection
        gps = np.sort(np.ascontiguousarray(allr[:, 20:28]).view("<f8").ravel())
        want = np.sort(np.nonzero(res.labels == cls)[0].astype(np.float64))
        assert (gps == want).all()
    assert len(got[2]) == -(-n_ground // 10000)


@pytest.mark.parametrize("mode", ["device", "host"])
def test_octstore_queries(tmp_path, mode):
    """findBlocks / pointsIn / countPointsIn / hiLoPointsIn vs brute force (cf. testflat, wolkentest.cpp:143-167),
    with the queries on the GPU (wb_query_*) and on the CPU mirror of the octree walk (--host)."""
    host = ["--host"] if mode == "host" else []
    cloud = synth.generate(2, 40000, seed=43)
    las = str(tmp_path / "q.las")
    cloud.write(las)
    res = O.run([O.file_from_cloud(cloud)], classify=False)
    pts = res.points_sorted
    c = [float(v) for v in pts.mean(axis=0)]
    # cylinder
    out = json.loads(subprocess.run([QUERY, las, "cyl", repr(c[0]), repr(c[1]), "3.5"] + host, capture_output=True, text=True,
                                    check=True).stdout.strip().splitlines()[-1])
    inside = np.hypot(pts[:, 0] - c[0], pts[:, 1] - c[1]) <= 3.5
    assert out["count"] == out["points"] == int(inside.sum()) and out["sorted"] == 1 and out["consistent"] == 1
    assert out["lo"] == pts[inside, 2].min() and out["hi"] == pts[inside, 2].max()
    assert out["total_points"] == cloud.n and out["total_blocks"] == len(res.leaves)
    assert 0 < out["blocks"] < len(res.leaves)
    # sphere
    out = json.loads(subprocess.run([QUERY, las, "sph", repr(c[0]), repr(c[1]), repr(c[2]), "2.0"] + host, capture_output=True,
                                    text=True, check=True).stdout.strip().splitlines()[-1])
    d = np.hypot(np.hypot(pts[:, 0] - c[0], pts[:, 1] - c[1]), pts[:, 2] - c[2])
    assert out["count"] == int((d <= 2.0).sum())
    # downward hyperboloid from 3 m above the centroid
    v = (c[0], c[1], c[2] + 3.0)
    out = json.loads(subprocess.run([QUERY, las, "hyp"] + [repr(x) for x in v] + ["0.5", "1"] + host, capture_output=True,
                                    text=True, check=True).stdout.strip().splitlines()[-1])
    por = 0.5
    zd = (v[2] + por) - pts[:, 2]
    dd = np.hypot(v[0] - pts[:, 0], v[1] - pts[:, 1])
    inside = (zd > 0) & (zd * zd - dd * dd >= por * por)
    assert out["count"] == int(inside.sum()) and out["count"] > 10
    # paraboloid and column against the oracle's Shape::in
    for kind, code, prm in (("par", 1, [c[0], c[1], c[2] + 2.0, 6.0]), ("col", 4, [c[0], c[1], 3.0])):
        out = json.loads(subprocess.run([QUERY, las, kind] + [repr(x) for x in prm] + host, capture_output=True,
                                        text=True, check=True).stdout.strip().splitlines()[-1])
        m = O.shape_filter(code, prm, pts)
        assert out["count"] == out["points"] == int(m.sum()) and out["count"] > 5, kind
        assert out["lo"] == pts[m, 2].min() and out["hi"] == pts[m, 2].max()
        assert out["sorted"] == 1 and out["consistent"] == 1


def _las14(path):
    raw = np.fromfile(path, dtype=np.uint8)
    h = raw[:375]
    info = {"version": (int(h[24]), int(h[25])), "header_size": int(h[94:96].view("<u2")[0]),
            "offset_to_points": int(h[96:100].view("<u4")[0]), "fmt": int(h[104]), "len": int(h[105:107].view("<u2")[0]),
            "scale": h[131:155].view("<f8").tolist(), "offset": h[155:179].view("<f8").tolist(),
            "maxmin": h[179:227].view("<f8").tolist(), "n": int(h[247:255].view("<u8")[0]),
            "by_return": h[255:375].view("<u8").tolist(), "legacy_n": int(h[107:111].view("<u4")[0]),
            "system_id": bytes(h[26:58]).rstrip(b"\0").decode()}
    recs = raw[info["offset_to_points"]:info["offset_to_points"] + info["n"] * info["len"]].reshape(info["n"], info["len"])
    return info, recs


@pytest.mark.parametrize("writer", ["device", "host"])
@pytest.mark.parametrize("scene,n,separate,ppf", [(5, 30000, 1, 8000), (2, 20000, 0, 0), (3, 15000, 1, 0)])
def test_reference_style_writer_matches_reference(tmp_path, scene, n, separate, ppf, writer):
    """CloudOutput + LasHeader::writePoint/writeHeader (cloudoutput.cpp:119-246, las.cpp:822-904):
    our writers (records made by wb_encode on the GPU, or by the C++ LasHeader::writePoint mirror with
    --host-writer) against the reference's own write path (oracle/_ref/ref_driver -w).  Which of a
    class's files a bucket lands in depends on the reference's block numbering, so files are
    compared per class as multisets of records, plus the header fields that do not depend on it."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(ref):
        pytest.skip("compiled reference not present")
    cloud = synth.generate(scene, n, seed=scene + 50)
    las = str(tmp_path / "in.las")
    cloud.write(las)
    rdir, odir = tmp_path / "ref", tmp_path / "ours"
    rdir.mkdir()
    odir.mkdir()
    subprocess.run([ref, "-t", "1", "-c", "-o", str(rdir / "r"), "-w", str(rdir / "out"), "-s", str(separate),
                    "-p", str(ppf), las], capture_output=True, text=True, check=True, cwd=str(rdir))
    cmd = [CLI, "-o", str(odir / "out"), "--separate-classes", str(separate), "--points-per-file", str(ppf),
           "--dump", str(odir / "dump"), las] + (["--host-writer"] if writer == "host" else [])
    subprocess.run(cmd, capture_output=True, text=True, check=True, cwd=str(odir))
    rfiles = sorted(f for f in os.listdir(rdir) if f.startswith("out") and f.endswith(".las"))
    ofiles = sorted(f for f in os.listdir(odir) if f.startswith("out") and f.endswith(".las"))
    assert rfiles == ofiles and len(rfiles) >= 1

    def group(d, files):
        g = {}
        for f in files:
            key = f.split("-")[1] if separate else "all"
            info, recs = _las14(str(d / f))
            g.setdefault(key, []).append((info, recs))
        return g

    gr, go = group(rdir, rfiles), group(odir, ofiles)
    assert gr.keys() == go.keys()
    for key in gr:
        ri, oi = gr[key][0][0], go[key][0][0]
        for f in ("version", "header_size", "offset_to_points", "fmt", "len", "scale", "offset", "system_id"):
            assert ri[f] == oi[f], (key, f, ri[f], oi[f])
        assert ri["version"] == (1, 4) and ri["header_size"] == 375
        rr = np.concatenate([x[1] for x in gr[key]])
        oo = np.concatenate([x[1] for x in go[key]])
        assert rr.shape == oo.shape
        rs = rr[np.lexsort(rr.T[::-1])]
        os_ = oo[np.lexsort(oo.T[::-1])]
        assert (rs == os_).all(), key
        # totals and the union of the bounding boxes
        assert sum(x[0]["n"] for x in gr[key]) == sum(x[0]["n"] for x in go[key])
        rb = np.array([x[0]["maxmin"] for x in gr[key]])
        ob = np.array([x[0]["maxmin"] for x in go[key]])
        assert (rb[:, 0::2].max(axis=0) == ob[:, 0::2].max(axis=0)).all() and (rb[:, 1::2].min(axis=0) == ob[:, 1::2].min(axis=0)).all()
        assert np.array([x[0]["by_return"] for x in gr[key]]).sum(axis=0).tolist() == \
            np.array([x[0]["by_return"] for x in go[key]]).sum(axis=0).tolist()
        if ppf:
            assert all(x[0]["n"] <= ppf + 537 for x in go[key])


def _scramble(cloud, seed):
    """Random intensity, return/flag bits, scan angle, user data and source id (return number kept
    non-zero) so that every field conversion of readPoint/writePoint is exercised."""
    rng = np.random.default_rng(seed)
    recs = cloud.records.copy()
    n = cloud.n
    last = 20 if cloud.fmt < 6 else 22
    recs[:, 12:last] = rng.integers(0, 256, (n, last - 12), dtype=np.uint8)
    if cloud.fmt < 6:
        recs[:, 14] = (recs[:, 14] & 0xf8) | rng.integers(1, 8, n, dtype=np.uint8)
    else:
        recs[:, 14] = (recs[:, 14] & 0xf0) | rng.integers(1, 16, n, dtype=np.uint8)
    return synth.Cloud(cloud.desc, cloud.header, recs, cloud.bbox)


def _write_both(tmp_path, inputs, extra):
    out = {}
    for w in ("device", "host"):
        d = tmp_path / w
        d.mkdir()
        cmd = [CLI, "-o", str(d / "out"), "--dump", str(d / "dump")] + extra + inputs + \
            (["--host-writer"] if w == "host" else [])
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(d))
        assert r.returncode == 0, r.stdout + r.stderr
        out[w] = {f: open(str(d / f), "rb").read() for f in sorted(os.listdir(d)) if f.endswith(".las")}
    return out


@pytest.mark.parametrize("case", ["mixed_1_6", "mixed_3_6", "duplicates", "zero_returns", "split_one_class"])
def test_device_writer_equals_host_writer(tmp_path, case):
    """wb_encode (records made on the GPU) produces the same bytes as the C++ mirror of
    LasHeader::readPoint/writePoint + CloudOutput::writeFiles — whole files, headers included — on
    format conversion (legacy -> 1.4 layouts, scan angle, RGB), identical locations (the last
    record's attributes in the first one's place), return number 0 (forced to 1) and file splitting."""
    extra = []
    if case.startswith("mixed"):
        sa, sb = (2, 3) if case == "mixed_1_6" else (5, 3)            # formats 1+6 -> 6, 3+6 -> 7
        d = synth.describe(sa, 20000)
        a = synth.generate(sa, 20000, seed=61)
        b = synth.generate(sb, 20000, seed=62, gps_base=a.n)
        a, b = _scramble(a, 1), _scramble(b, 2)
        clouds = [a, b]
        assert {a.fmt, b.fmt} == ({1, 6} if case == "mixed_1_6" else {3, 6})
        extra = ["--points-per-file", "9000"]
    elif case == "duplicates":
        clouds = [synth.with_duplicates(synth.with_duplicates(_scramble(synth.generate(2, 20000, seed=63), 3),
                                                              6000, 1), 6000, 2)]
    elif case == "zero_returns":
        c = synth.generate(2, 20000, seed=64)
        recs = c.records.copy()
        recs[:, 14] &= 0xf8                                            # return number 0 everywhere: kept, written as 1
        clouds = [synth.Cloud(c.desc, c.header, recs, c.bbox)]
    else:
        clouds = [_scramble(synth.generate(1, 30000, seed=65), 4)]
        extra = ["--separate-classes", "0", "--points-per-file", "7000"]
    inputs = []
    for i, c in enumerate(clouds):
        p = str(tmp_path / ("in%d.las" % i))
        c.write(p)
        inputs.append(p)
    out = _write_both(tmp_path, inputs, extra)
    assert list(out["device"].keys()) == list(out["host"].keys()) and len(out["device"]) >= 1
    for f in out["device"]:
        a, b = out["device"][f], out["host"][f]
        assert len(a) == len(b), f
        if a != b:
            aa, bb = np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8)
            bad = np.nonzero(aa != bb)[0]
            raise AssertionError("%s differs at %d bytes, first at offset %d" % (f, len(bad), bad[0]))
    if case == "zero_returns":
        info, recs = _las14(str(tmp_path / "device" / sorted(out["device"])[0]))
        assert ((recs[:, 14] & 7) == 1).all()
    if case.startswith("mixed"):
        info, recs = _las14(str(tmp_path / "device" / sorted(out["device"])[0]))
        assert info["fmt"] == (6 if case == "mixed_1_6" else 7)


def _strip_files(tmp_path, scene, n, strips, seed=61):
    d = synth.describe(scene, n)
    cuts = [d.grid_nx * k // strips for k in range(strips + 1)]
    clouds, names, base = [], [], 0
    for k in range(strips):
        c = synth.generate(scene, n, seed=seed, region=(cuts[k], 0, cuts[k + 1] - cuts[k], d.grid_ny), gps_base=base)
        base += c.n
        name = str(tmp_path / ("strip%d.las" % k))
        c.write(name)
        clouds.append(c)
        names.append(name)
    return clouds, names


@pytest.mark.parametrize("gpus,strips", [(2, 3), (4, 4)])
def test_cli_gpus_equals_oracle_and_one_gpu(tmp_path, gpus, strips):
    """wolkencli --gpus N (startThreads(N): one worker per GPU, x-strips, wb_shard_run) against the oracle's labels
    for the whole cloud and against the same command line on one GPU.  The files are given in shuffled order: the
    CLI puts them in ascending x.  On a one-GPU box the workers share the device over the LOCAL transport
    (WOLKEN_TRANSPORT=local); with one GPU per worker the transport is NCCL (tests/test_multigpu_nccl.py)."""
    import torch
    clouds, names = _strip_files(tmp_path, 2, 60000, strips)
    env = dict(os.environ)
    if os.environ.get("WB_EMULATED") or torch.cuda.device_count() < gpus:
        env["WOLKEN_TRANSPORT"] = "local"
    shuffled = names[1:] + names[:1]
    outs = {}
    for g in (gpus, 1):
        o = subprocess.run([CLI, "--gpus", str(g), "-o", str(tmp_path / ("o%d" % g)), "--lossless", "--separate-classes",
                            "0", "--dump", str(tmp_path / ("d%d" % g))] + (shuffled if g > 1 else names), capture_output=True,
                           text=True, env=env)
        assert o.returncode == 0, o.stdout[-2000:] + o.stderr[-2000:]
        outs[g] = o.stdout
    assert "%d GPUs" % min(gpus, strips) in outs[gpus] and "GPUs" not in outs[1]
    res = O.run([O.file_from_cloud(c) for c in clouds])
    _, recs, _ = _read_las(str(tmp_path / ("o%d.las" % gpus)))
    assert recs.shape[0] == sum(c.n for c in clouds)
    assert ((recs[:, 15] & 31) == res.labels).all()
    assert open(tmp_path / ("o%d.las" % gpus), "rb").read() == open(tmp_path / "o1.las", "rb").read()
    assert open(tmp_path / ("d%d" % gpus), encoding="utf-8").read() == res.dump


def test_cli_census_after_writing(tmp_path):
    """censusPoints after every write (threads.cpp:613): the synthetic files carry the point number as GPS time; a
    file with records cut out must be reported with exactly those numbers missing."""
    cloud = synth.generate(2, 12000, seed=71)
    gone = [5, 700, 701]
    recs = np.delete(cloud.records, gone, axis=0)
    cut = synth.Cloud(cloud.desc, cloud.header, np.ascontiguousarray(recs), cloud.bbox)
    whole, holes = str(tmp_path / "whole.las"), str(tmp_path / "holes.las")
    cloud.write(whole)
    cut.write(holes)
    a = subprocess.run([CLI, "-o", str(tmp_path / "a"), "--dump", str(tmp_path / "da"), whole], capture_output=True, text=True)
    assert a.returncode == 0, a.stdout[-1500:] + a.stderr[-1500:]
    assert "Max point %d" % cloud.n in a.stdout and "Missing points" not in a.stdout and "Duplicate point" not in a.stdout
    b = subprocess.run([CLI, "-o", str(tmp_path / "b"), "--dump", str(tmp_path / "db"), holes], capture_output=True, text=True)
    assert b.returncode == 0, b.stdout[-1500:] + b.stderr[-1500:]
    assert "Max point %d" % cloud.n in b.stdout and "Missing points: 5,700,701" in b.stdout
    # the host walk (records not on the device) says the same
    c = subprocess.run([CLI, "-o", str(tmp_path / "c"), "--host-writer", "--census", "--dump", str(tmp_path / "dc"), holes],
                       capture_output=True, text=True)
    assert c.returncode == 0 and "Missing points: 5,700,701" in c.stdout, c.stdout[-1500:] + c.stderr[-1500:]


@pytest.mark.parametrize("gpus,files", [(3, 1), (4, 2)])
def test_cli_gpus_on_fewer_files_than_gpus(tmp_path, gpus, files):
    """One big file (or fewer files than GPUs): every worker reads every file and keeps its x-interval (wb_set_window);
    the class bytes come back in input order.  Output identical to --gpus 1, labels equal to the oracle's."""
    import torch
    clouds, names = _strip_files(tmp_path, 2, 50000, files, seed=63)
    env = dict(os.environ)
    if os.environ.get("WB_EMULATED") or torch.cuda.device_count() < gpus:
        env["WOLKEN_TRANSPORT"] = "local"
    outs = {}
    for g in (gpus, 1):
        o = subprocess.run([CLI, "--gpus", str(g), "-o", str(tmp_path / ("o%d" % g)), "--lossless", "--separate-classes",
                            "0", "--dump", str(tmp_path / ("d%d" % g))] + names, capture_output=True, text=True, env=env)
        assert o.returncode == 0, o.stdout[-2000:] + o.stderr[-2000:]
        outs[g] = o.stdout
    assert "%d GPUs" % gpus in outs[gpus]
    res = O.run([O.file_from_cloud(c) for c in clouds])
    _, recs, _ = _read_las(str(tmp_path / ("o%d.las" % gpus)))
    assert ((recs[:, 15] & 31) == res.labels).all()
    assert open(tmp_path / ("o%d.las" % gpus), "rb").read() == open(tmp_path / "o1.las", "rb").read()

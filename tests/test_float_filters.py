"""The two single-precision shortcuts of the classify kernel (wb_kernels.cuh: wb_sector64f and the
(chunk, query) reach test at expansion), mirrored in numpy float32 and checked against what they
stand in for: the sector of the exact atan2i (angle.cpp:117-155) and the double-precision wb_reach.
The shortcuts may only ever answer "ask the exact routine" / "maybe": a wrong definite answer would
change labels.  (The GPU tests prove the labels; this pins the error analysis on the CPU.)"""
import numpy as np

from oracle import wb_oracle as O

f32 = np.float32
T = [0.09849140335716425, 0.198912367379658, 0.3033466836073424, 0.41421356237309503,
     0.5345111359507916, 0.6681786379192989, 0.8206787908286602]


def sector64f(dxd, dyd):
    t = [f32(v) for v in T]
    ax, ay = np.abs(dxd.astype(f32)), np.abs(dyd.astype(f32))
    sw = ay > ax
    lo, hi = np.where(sw, ax, ay), np.where(sw, ay, ax)
    c = hi * t[3]
    b1 = lo >= c
    m = np.abs(lo - c)
    c = hi * np.where(b1, t[5], t[1])
    b2 = lo >= c
    m = np.minimum(m, np.abs(lo - c))
    c = hi * np.where(b1, np.where(b2, t[6], t[4]), np.where(b2, t[2], t[0]))
    b3 = lo >= c
    m = np.minimum(m, np.abs(lo - c))
    m = np.minimum(m, np.minimum(lo, hi - lo))
    undecided = ~(m > f32(4e-6) * hi) | ~(hi > f32(1e-30)) | ~(hi < f32(1e30))
    sub = b1 * 4 + b2 * 2 + b3 * 1
    s1 = np.where(sw, 15 - sub, sub)
    s = np.where(dyd >= 0, np.where(dxd >= 0, s1, 31 - s1), np.where(dxd < 0, 32 + s1, 63 - s1))
    return np.where(undecided, -1, s)


def test_float_sector_is_exact_or_undecided():
    rng = np.random.default_rng(5)
    L = O.lib()
    n = 300000
    ang = rng.uniform(0, 2 * np.pi, n)
    # a third of the directions within 1e-5..1e-9 rad of a sector edge (multiples of 5.625 degrees)
    k = rng.integers(0, 64, n // 3)
    ang[: n // 3] = k * (np.pi / 32) + rng.choice([1e-5, 1e-6, 1e-7, 1e-8, 1e-9], n // 3) * rng.uniform(-1, 1, n // 3)
    r = 10 ** rng.uniform(-3, 4, n)
    dx, dy = r * np.cos(ang), r * np.sin(ang)
    s = sector64f(dx, dy)
    exact = np.array([(L.wbo_atan2i(float(y), float(x)) & 0x7fffffff) >> 25 for x, y in zip(dx, dy)])
    decided = s >= 0
    assert (s[decided] == exact[decided]).all()
    assert decided.mean() > 0.6                      # the shortcut answers nearly always on random bearings...
    assert decided[n // 3:].mean() > 0.999
    assert (~decided[: n // 3]).sum() > 1000         # ...and defers right at the edges


def _reach_double(qx, qy, qcz, qpor2, s2, b):
    dx = np.maximum(0.0, np.maximum(b[0] - qx, qx - b[1]))
    dy = np.maximum(0.0, np.maximum(b[2] - qy, qy - b[3]))
    zl = qcz - b[4]
    d2 = (dx * dx + dy * dy) * s2
    return (zl > 0) & (zl * zl * (1 + 1e-12) - d2 * (1 - 1e-12) >= qpor2 * (1 - 1e-12))


def _reach_float(qx, qy, qcz, qpor2, s2, b, org, folded=True):
    ox, oy, oz = org
    por = np.sqrt(qpor2)
    fx, fy = (qx - ox).astype(f32), (qy - oy).astype(f32)
    fh = ((qcz - por) - oz).astype(f32)
    f2p = np.minimum((2 * por).astype(f32), f32(1e30))
    gh = f32(np.max(np.maximum(np.abs(fx), np.abs(fy))))
    zq = f32(np.max(np.abs(fh)))
    x0, x1 = (b[0] - ox).astype(f32), (b[1] - ox).astype(f32)
    y0, y1 = (b[2] - oy).astype(f32), (b[3] - oy).astype(f32)
    z0 = (b[4] - oz).astype(f32)
    c = f32(9.5367431640625e-7)
    ed = c * (np.maximum(np.maximum(np.abs(x0), np.abs(x1)), np.maximum(np.abs(y0), np.abs(y1))) + gh)
    ez = c * (np.abs(z0) + zq) + f32(1e-6)
    if not folded:                                            # WB_CL_F4=0: the slack applied per (child, query)
        fs2 = f32(s2) * f32(0.999998)
        dx = np.maximum(f32(0), np.maximum(x0 - fx, fx - x1) - ed)
        dy = np.maximum(f32(0), np.maximum(y0 - fy, fy - y1) - ed)
        a = (fh - z0) + ez
        return (a >= 0) & (a * (a + f2p) * f32(1.000002) >= (dx * dx + dy * dy) * fs2)
    # the shipped form: slack folded into the child's box before the loop over the queries
    bx0, bx1, by0, by1, bz0 = x0 - ed, x1 + ed, y0 - ed, y1 + ed, z0 - ez
    k2 = f32(s2) * f32(0.999995)
    dx = np.maximum(f32(0), np.maximum(bx0 - fx, fx - bx1))
    dy = np.maximum(f32(0), np.maximum(by0 - fy, fy - by1))
    a = fh - bz0
    return (a >= 0) & (a * (a + f2p) >= (dx * dx + dy * dy) * k2)


def test_float_reach_admits_everything_the_double_test_admits():
    rng = np.random.default_rng(1)
    total = extra = 0
    for _ in range(120):
        base = rng.choice([0.0, 1e3, 5e5, 4.2e6])              # local, UTM easting, UTM northing
        span = rng.choice([50.0, 2000.0, 50000.0])
        n = 20000
        org = (base + rng.uniform(0, span), base * 0.7 + rng.uniform(0, span), rng.uniform(-50, 3000))
        qx, qy = org[0] + rng.uniform(-3, 3, n), org[1] + rng.uniform(-3, 3, n)
        qz = org[2] + rng.uniform(-10, 10, n)
        s = rng.choice([0.1, 0.5, 1.0, 2.0, 7.0])
        s2 = s * s
        por = 10 ** rng.uniform(-1, 3.2, n) * s2               # hyperboloidSize 0.1 m .. 1.6 km
        qpor2, qcz = por * por, qz + por
        w = 10 ** rng.uniform(-1, 2, n)
        cx, cy = org[0] + rng.uniform(-1, 1, n) * span, org[1] + rng.uniform(-1, 1, n) * span
        b = [cx - w, cx + w, cy - w, cy + w, None]
        dx = np.maximum(0.0, np.maximum(b[0] - qx, qx - b[1]))
        dy = np.maximum(0.0, np.maximum(b[2] - qy, qy - b[3]))
        surf = qcz - np.sqrt(qpor2 + (dx * dx + dy * dy) * s2)
        # the box's lowest point right on the surface, give or take ulps .. metres
        eps = rng.choice([0, 1e-12, 1e-9, 1e-6, 1e-3, 1.0], n) * rng.uniform(-1, 1, n)
        b[4] = surf + eps * np.maximum(1, np.abs(surf))
        d = _reach_double(qx, qy, qcz, qpor2, s2, b)
        f = _reach_float(qx, qy, qcz, qpor2, s2, b, org)
        assert not (d & ~f).any()
        assert not (d & ~_reach_float(qx, qy, qcz, qpor2, s2, b, org, folded=False)).any()
        total += n
        extra += int((f & ~d).sum())
    assert extra < 0.25 * total                              # and it still prunes (these inputs sit ON the surface)


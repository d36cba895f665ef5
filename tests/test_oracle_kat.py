"""Pin the CPU oracle against the reference's own known-answer tests (wolkentest.cpp).

Each test restates the constants of one reference test; citations give the lines."""
import ctypes as C

import numpy as np

from oracle import wb_oracle as O


def _hin(v, r, s, p):
    return bool(O.lib().wbo_hyperboloid_in((C.c_double * 3)(*v), r, s, (C.c_double * 3)(*p)))


def test_hyperboloid_kat():
    # testhyperboloid, wolkentest.cpp:223-246: 72^2 = 78^2-18^2-24^2 = 97^2-25^2-60^2
    ver = (100, 200, 300)
    a, b = (118, 176, 294), (75, 260, 275)
    c, d = (118, 176, 295), (75, 260, 276)
    e, f = (118, 176, 293), (75, 260, 274)
    assert _hin(ver, 72, 1, a) and _hin(ver, 72, 1, b)
    assert not _hin(ver, 72, 1, c) and not _hin(ver, 72, 1, d)
    assert _hin(ver, 72, 1, e) and _hin(ver, 72, 1, f)


def test_cylinder_kat():
    # testcylinder, wolkentest.cpp:265-279: Cylinder((0,0),13)
    L = O.lib()
    inn = lambda p: bool(L.wbo_cylinder_in(0.0, 0.0, 13.0, float(p[0]), float(p[1])))
    assert inn((0, 0)) and inn((5, 12)) and not inn((5, 13)) and inn((-9, -9)) and not inn((-11, -7))
    cube = lambda c, s: bool(L.wbo_cylinder_intersects_cube(0.0, 0.0, 13.0, (C.c_double * 3)(*c), s))
    assert cube((0, 0, 13), 1) and cube((5, 12, 0), 0.1) and cube((-11, -7, 2), 1 / 16.)
    assert not cube((-11, -7, 2), 1 / 32.)


def test_paraboloid_sphere_column_kat():
    # testparaboloid, wolkentest.cpp:188-202: Paraboloid((0,0,13),13)
    par = (0, 0, 13, 13)
    v, a, b, c, d, e = (0, 0, 13), (5, 0, 12), (5, 2, 12), (13, 13, 0), (-14, -12, 0), (9, 16, 0)
    assert O.shape_in(1, par, v) and O.shape_in(1, par, a) and not O.shape_in(1, par, b)
    assert O.shape_in(1, par, c) and not O.shape_in(1, par, d) and O.shape_in(1, par, e)
    assert O.shape_intersects_cube(1, par, v, 1) and not O.shape_intersects_cube(1, par, b, 0.1)
    # testsphere, wolkentest.cpp:204-221: 89^2 = 15^2+36^2+80^2 = 39^2+48^2+64^2
    sph = (100, 200, 300, 89)
    for p, want in [((100, 200, 300), True), ((115, 164, 220), True), ((115, 163, 220), False), ((36, 239, 252), True),
                    ((36, 240, 252), False), ((64, 120, 315), True), ((63, 120, 315), False), ((164, 248, 339), True),
                    ((164, 248, 340), False)]:
        assert O.shape_in(0, sph, p) == want, p
    # the generic entry agrees with the dedicated ones (testhyperboloid, testcylinder)
    assert O.shape_in(2, (100, 200, 300, 72, 1), (118, 176, 294)) and not O.shape_in(2, (100, 200, 300, 72, 1), (118, 176, 295))
    assert O.shape_in(3, (0, 0, 13), (5, 12, 0)) and not O.shape_in(3, (0, 0, 13), (5, 13, 4))
    assert O.shape_intersects_cube(3, (0, 0, 13), (-11, -7, 2), 1 / 16.) and not O.shape_intersects_cube(3, (0, 0, 13), (-11, -7, 2), 1 / 32.)
    # Column, shape.cpp:252-255: inclusive on both sides
    assert O.shape_in(4, (10, 20, 2), (11, 19, 5)) and O.shape_in(4, (10, 20, 2), (9, 21, -3)) and not O.shape_in(4, (10, 20, 2), (11.0001, 20, 0))


def test_least_squares_kat():
    # testleastsquares, wolkentest.cpp:744-760
    x = O.least_squares([[1, 3], [2, 4], [1, 6]], [4, 1, 3])
    assert np.hypot(x[0] + 29 / 77., x[1] - 51 / 77.) < 1e-9


def test_flowsnake_roundtrip():
    # testflowsnake, wolkentest.cpp:792-796: baseSeven(baseFlow(n)) == n, here through the
    # composed toFlowsnake and our inverse, plus adjacency of successive tiles (flowsnake.h:22-25)
    rng = np.random.default_rng(5)
    prev = None
    for n in list(range(-900, 1501)) + rng.integers(-24242424, 3603000, 300).tolist():
        ex, ey = O.to_flowsnake(int(n))
        assert O.from_flowsnake(ex, ey) == n
    for n in range(-214, 129):
        e = O.to_flowsnake(n)
        if prev is not None:
            dx, dy = e[0] - prev[0], e[1] - prev[1]
            assert dx * dx + dy * dy - dx * dy == 1      # Eisenstein norm of the step
        prev = e
    # quadraticSize table, wolkentest.cpp:776: tiles of the order-i island = 7^i
    lo = [0, -4, -18, -214, -900, -10504]
    hi = [0, 2, 30, 128, 1500, 6302]
    for i in range(6):
        assert hi[i] - lo[i] + 1 == 7 ** i


def test_atan2i_roundtrip():
    # testintegertrig, wolkentest.cpp:122-132: foldangle(atan2i(cossin(i))) == foldangle(i)
    L = O.lib()
    L.wbo_fill_tan_tables()

    def fold(a):
        a &= 0xffffffff
        if (a >> 30) % 3:
            a ^= 0x80000000
        return a

    rng = np.random.default_rng(1)
    i = -2147400000
    pi = np.longdouble(np.pi) if False else np.arctan(np.longdouble(1)) * 4
    while i < 2147400000:
        ang = np.longdouble(i) * pi / np.longdouble(1073741824.)
        c, s = float(np.cos(ang)), float(np.sin(ang))
        assert fold(L.wbo_atan2i(s, c)) == fold(i), hex(i & 0xffffffff)
        i += int(rng.integers(0, 65536)) * 300 + 18000
    assert L.wbo_atan2i(0.0, 0.0) == 0


def test_pairwise_sum_matches_definition():
    # pairwisesum, manysum.cpp:120-154 restated literally in numpy (binary counter)
    def ref(a):
        sums = {}
        for i, v in enumerate(a):
            lvl = 0
            j = i
            v = np.float64(v)
            while j & 1:
                v = sums[lvl] + v
                j >>= 1
                lvl += 1
            sums[lvl] = v
        s = np.float64(0)
        n = len(a)
        for lvl in range(32):
            if (n >> lvl) & 1:
                s = s + sums[lvl]
        return s

    rng = np.random.default_rng(3)
    for n in [1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 100, 537, 1000]:
        a = rng.normal(size=n) * 10.0 ** rng.integers(-3, 4, size=n)
        got = O.lib().wbo_pairwise_sum(np.ascontiguousarray(a).ctypes.data, n)
        assert got == ref(a)


def test_surround():
    # surround, classify.cpp:67-94: at least two directions, no gap >= 144 degrees (2^31 = 360)
    D = 1 << 31
    deg = lambda x: int(round(x / 360.0 * D)) - (D if x >= 180 else 0)
    assert not O.surround([])
    assert not O.surround([deg(10)])
    assert not O.surround([deg(0), deg(120)])
    assert O.surround([deg(0), deg(120), deg(240)])
    assert not O.surround([deg(0), deg(100), deg(200)])          # 160 degree gap 200 -> 360
    assert O.surround([deg(0), deg(143), deg(250)])
    assert not O.surround([deg(0), deg(145), deg(250)])
    assert not O.surround([deg(5), deg(5), deg(5)])


def test_ldecimal():
    # outputs of the reference's own ldecimal() (ldecimal.cpp, linked from oracle/_ref) for these inputs
    for x, s in [(5e5, "5e5"), (42e5, "42e5"), (500006.0, "500006"), (124.0, "124"), (0.25, ".25"),
                 (4.0, "4"), (-0.5, "-.5"), (1024.0, "1024"), (0.001953125, ".001953125"), (0.0, "0"),
                 (1e-7, "1e-7"), (123456.789, "123456.789"), (2000.0, "2e3"), (30000.0, "3e4"),
                 (200.0, "200"), (0.1, ".1"), (1e22, "1e22"), (-1234.5e-9, "-1.2345e-6"), (3.0e-5, "3e-5")]:
        assert O.ldecimal(x) == s, (x, O.ldecimal(x), s)

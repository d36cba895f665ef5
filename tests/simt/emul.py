"""ctypes front end of the SIMT emulator build of the classify kernels (tests/simt/classify_emul.cpp).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(variant="", out="libwb_simt.so"):
    """make the library; a change of the -D flags behind the same file name forces a rebuild (make only sees files)."""
    stamp = os.path.join(_HERE, out + ".flags")
    old = open(stamp).read() if os.path.exists(stamp) else None
    cmd = ["make", "-s", "-C", _HERE, "OUT=" + out, "VARIANT=" + variant]
    if old != variant:
        cmd.insert(1, "-B")
    subprocess.check_call(cmd + [out])
    with open(stamp, "w") as f:
        f.write(variant)
    return os.path.join(_HERE, out)


_LIBS = {}


def lib(variant="", out="libwb_simt.so"):
    key = (variant, out)
    if key not in _LIBS:
        L = C.CDLL(build(variant, out))
        from oracle import wb_oracle as O
        OL = O.lib()
        OL.wbo_fill_tan_tables()
        for f in ("wbo_tan_table", "wbo_cos_table", "wbo_sin_table"):
            getattr(OL, f).restype = C.c_void_p
        L.simt_set_tables(C.c_void_p(OL.wbo_tan_table()), C.c_void_p(OL.wbo_cos_table()), C.c_void_p(OL.wbo_sin_table()))
        L.simt_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_double,
                                    C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIBS[key] = L
    return _LIBS[key]


_COLUMNS = {}


def _columns(points_sorted, hyp):
    """Contiguous x, y, z, hyp columns, kept while the same arrays are passed again (the emulator library keeps
    its hierarchy per column pointer: sampling warps of a large cloud must not rebuild it per call)."""
    key = (id(points_sorted), id(hyp))
    if key not in _COLUMNS:
        _COLUMNS.clear()
        for L in _LIBS.values():
            L.simt_reset()
        _COLUMNS[key] = (points_sorted, hyp,
                         np.ascontiguousarray(points_sorted[:, 0]), np.ascontiguousarray(points_sorted[:, 1]),
                         np.ascontiguousarray(points_sorted[:, 2]), np.ascontiguousarray(hyp, dtype=np.float64))
    return _COLUMNS[key][2:]


def classify(points_sorted, hyp, max_slope=1.0, thickness=0.0, chunks=None, variant="", out="libwb_simt.so"):
    """Labels (canonical order) for the queries of chunks [chunks[0], chunks[1]) — all by default; 0 for points in
    no tile and (partial runs) for chunks not asked for.  Also returns the kernel's work counters and the number
    of warp-wide intrinsics executed."""
    L = lib(variant, out)
    n = len(points_sorted)
    sx, sy, sz, hyp = _columns(points_sorted, hyp)
    c0, c1 = chunks if chunks else (0, 0xffffffff)
    lab = np.zeros(n, dtype=np.uint8)
    counters = np.zeros(24, dtype=np.uint64)
    coll = C.c_ulonglong()
    L.simt_classify(sx.ctypes.data, sy.ctypes.data, sz.ctypes.data, n, hyp.ctypes.data, max_slope, thickness,
                    c0, c1, lab.ctypes.data, counters.ctypes.data, C.byref(coll))
    work = {"margin": int(counters[0]), "untiled": int(counters[1]), "second_walk_points": int(counters[6]),
            "nodes": int(counters[8]), "chunks": int(counters[9]), "pairs": int(counters[10]),
            "nodes2": int(counters[11]), "chunks2": int(counters[12]), "pairs2": int(counters[13]),
            "collectives": coll.value}
    return lab, work


def site_counts(variant="", out="libwb_simt.so", clear=True):
    """{source line of wb_kernels.cuh: times a warp reached the warp-wide intrinsic on it} since the last clear."""
    L = lib(variant, out)
    c = np.zeros(4096, dtype=np.uint64)
    L.simt_site_counts(c.ctypes.data, 1 if clear else 0)
    return {int(i): int(c[i]) for i in np.nonzero(c)[0]}


def emu_counts(variant="", out="libwb_simt.so", clear=True):
    """Loop-trip counters (WB_EMU_COUNT slots): 0 = (chunk, query) reach tests at expansion, 1 = the same for
    internal nodes (WB_CL_XWANTS)."""
    L = lib(variant, out)
    c = np.zeros(16, dtype=np.uint64)
    L.simt_emu_counts(c.ctypes.data, 1 if clear else 0)
    return [int(x) for x in c]


SIMT_TILE = np.dtype([("n", "<i4"), ("nPoints", "<i4"), ("treeFlags", "<i4"), ("_pad", "<i4"),
                      ("density", "<f8"), ("hyperboloidSize", "<f8"), ("height", "<f8")])


def scan_classify(points_sorted, cube, tile_size=1.0, min_hyperboloid_size=0.1, max_slope=1.0, thickness=0.0,
                  postscan=True, classify=True, variant="", out="libwb_simt.so"):
    """Membership, tile scan, postscan and classify from the kernel sources, starting from the canonical order.
    Returns (non-empty tiles ascending n, labels in canonical order or None)."""
    L = lib(variant, out)
    n = len(points_sorted)
    sx = np.ascontiguousarray(points_sorted[:, 0])
    sy = np.ascontiguousarray(points_sorted[:, 1])
    sz = np.ascontiguousarray(points_sorted[:, 2])
    cap = 3 * n + 16
    tiles = np.zeros(cap, dtype=SIMT_TILE)
    nt = C.c_uint64()
    lab = np.zeros(n, dtype=np.uint8) if classify else None
    cube4 = (C.c_double * 4)(*cube)
    L.simt_scan_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.simt_scan_classify(sx.ctypes.data, sy.ctypes.data, sz.ctypes.data, n, cube4, tile_size, min_hyperboloid_size,
                         max_slope, thickness, 1 if postscan else 0, tiles.ctypes.data, cap, C.byref(nt),
                         lab.ctypes.data if classify else None)
    assert nt.value <= cap
    return tiles[:nt.value].copy(), lab


def radix_sort(keys, vals, begin_bit=0, end_bit=64, variant="", out="libwb_simt.so"):
    """wb_radix_sort (wb_sort.cuh) under the emulator: returns the sorted (keys, vals) copies."""
    L = lib(variant, out)
    k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    L.simt_radix_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    L.simt_radix_sort(k.ctypes.data, v.ctypes.data, len(k), begin_bit, end_bit)
    return k, v


def decode(records, fmt, drop_zeros=False, misalign=0, variant="", out="libwb_simt.so"):
    """wb_decode_kernel under the emulator: (x, y, z int32, class u8, return number u8, dropped count)."""
    L = lib(variant, out)
    recs = np.ascontiguousarray(records)
    n, rec_len = recs.shape
    x, y, z = (np.zeros(n, dtype=np.int32) for _ in range(3))
    cls, ret = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
    dropped = C.c_ulonglong()
    L.simt_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    L.simt_decode(recs.ctypes.data, n, fmt, rec_len, 1 if drop_zeros else 0, misalign, x.ctypes.data, y.ctypes.data,
                  z.ctypes.data, cls.ctypes.data, ret.ctypes.data, C.byref(dropped))
    return x, y, z, cls, ret, dropped.value

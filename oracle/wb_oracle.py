"""ctypes front end of the CPU oracle (oracle/wb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by wolkenbase_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Tile(C.Structure):
    _fields_ = [("n", C.c_int32), ("ex", C.c_int32), ("ey", C.c_int32), ("nPoints", C.c_int32),
                ("treeFlags", C.c_int32), ("density", C.c_double), ("hyperboloidSize", C.c_double),
                ("height", C.c_double)]


class Leaf(C.Structure):
    _fields_ = [("first", C.c_uint64), ("count", C.c_uint32), ("depth", C.c_int32),
                ("cx", C.c_double), ("cy", C.c_double), ("cz", C.c_double), ("half", C.c_double)]


TILE_DTYPE = np.dtype([("n", "<i4"), ("ex", "<i4"), ("ey", "<i4"), ("nPoints", "<i4"), ("treeFlags", "<i4"),
                       ("_pad", "<i4"), ("density", "<f8"), ("hyperboloidSize", "<f8"), ("height", "<f8")])
LEAF_DTYPE = np.dtype([("first", "<u8"), ("count", "<u4"), ("depth", "<i4"),
                       ("cx", "<f8"), ("cy", "<f8"), ("cz", "<f8"), ("half", "<f8")])
assert TILE_DTYPE.itemsize == C.sizeof(Tile) and LEAF_DTYPE.itemsize == C.sizeof(Leaf)


def build(force=False):
    so = os.path.join(_HERE, "_build", "libwb_oracle.so")
    src = os.path.join(_HERE, "wb_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "port"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.wbo_atan2i.argtypes = [C.c_double, C.c_double]
        L.wbo_hyperboloid_in.argtypes = [dp, C.c_double, C.c_double, dp]
        L.wbo_cylinder_in.argtypes = [C.c_double] * 5
        L.wbo_cylinder_intersects_cube.argtypes = [C.c_double, C.c_double, C.c_double, dp, C.c_double]
        L.wbo_to_flowsnake.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.wbo_from_flowsnake.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.wbo_base_seven.argtypes = [C.c_int, C.c_int]
        L.wbo_pairwise_sum.argtypes = [C.c_void_p, C.c_uint]
        L.wbo_pairwise_sum.restype = C.c_double
        L.wbo_least_squares.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.wbo_surround.argtypes = [C.c_void_p, C.c_int]
        L.wbo_ldecimal.argtypes = [C.c_double, C.c_char_p, C.c_int]
        L.wbo_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.wbo_coords.argtypes = [C.c_void_p, C.c_uint64, dp, dp, C.c_double, C.c_void_p]
        L.wbo_size_fit.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.wbo_bbox_cube.argtypes = [C.c_void_p, C.c_int, dp]
        L.wbo_snake_set_size.argtypes = [C.c_double, C.c_double, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.wbo_morton_key.argtypes = [dp, dp, C.c_double]
        L.wbo_morton_key.restype = C.c_uint64
        L.wbo_sort.argtypes = [C.c_void_p, C.c_uint64, dp, C.c_double, C.c_void_p, C.c_void_p]
        L.wbo_leaves.argtypes = [C.c_void_p, C.c_uint64, dp, C.c_double, C.c_void_p, C.c_int64]
        L.wbo_leaves.restype = C.c_int64
        L.wbo_dump.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]
        L.wbo_dump.restype = C.c_int64
        L.wbo_scan.argtypes = [C.c_void_p, C.c_uint64, dp, C.c_double, C.c_double, C.c_void_p, C.c_int64]
        L.wbo_scan.restype = C.c_int64
        L.wbo_postscan.argtypes = [C.c_void_p, C.c_int64, C.c_double]
        L.wbo_point_hyperboloid_sizes.argtypes = [C.c_void_p, C.c_uint64, dp, C.c_double, C.c_void_p, C.c_int64,
                                                  C.c_void_p]
        L.wbo_classify.argtypes = [C.c_void_p, C.c_uint64, dp, C.c_double, C.c_double, C.c_double,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.wbo_classify_sel.argtypes = [C.c_void_p, C.c_uint64, dp, C.c_double, C.c_double, C.c_double,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.wbo_shape_in.argtypes = [C.c_int, dp, dp]
        L.wbo_shape_intersects_cube.argtypes = [C.c_int, dp, dp, C.c_double]
        L.wbo_shape_filter.argtypes = [C.c_int, dp, C.c_void_p, C.c_uint64, C.c_void_p]
        L.wbo_tan_table.restype = dp
        L.wbo_cos_table.restype = dp
        L.wbo_sin_table.restype = dp
        _LIB = L
    return _LIB


def _d3(v):
    return (C.c_double * len(v))(*v)


def tan_tables():
    """(tan[511], cos[512], sin[512]) exactly as fillTanTables (angle.cpp:305-320) makes them."""
    L = lib()
    return (np.ctypeslib.as_array(L.wbo_tan_table(), (511,)).copy(),
            np.ctypeslib.as_array(L.wbo_cos_table(), (512,)).copy(),
            np.ctypeslib.as_array(L.wbo_sin_table(), (512,)).copy())


def ldecimal(x):
    buf = C.create_string_buffer(64)
    lib().wbo_ldecimal(x, buf, 64)
    return buf.value.decode()


def to_flowsnake(n):
    ex, ey = C.c_int(), C.c_int()
    lib().wbo_to_flowsnake(n, C.byref(ex), C.byref(ey))
    return ex.value, ey.value


def from_flowsnake(ex, ey):
    n = C.c_int64()
    if lib().wbo_from_flowsnake(ex, ey, C.byref(n)):
        return None
    return n.value


def least_squares(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(a.shape[1])
    lib().wbo_least_squares(a.ctypes.data, b.ctypes.data, a.shape[0], a.shape[1], x.ctypes.data)
    return x


def shape_in(kind, params, p):
    q = list(params) + [0.0] * (6 - len(params))
    return bool(lib().wbo_shape_in(kind, _d3(q), _d3(p)))


def shape_intersects_cube(kind, params, center, side):
    q = list(params) + [0.0] * (6 - len(params))
    return bool(lib().wbo_shape_intersects_cube(kind, _d3(q), _d3(center), side))


def shape_filter(kind, params, pts):
    """Shape::in over an (n,3) array -> bool mask."""
    q = list(params) + [0.0] * (6 - len(params))
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.zeros(len(pts), dtype=np.uint8)
    lib().wbo_shape_filter(kind, _d3(q), pts.ctypes.data, len(pts), out.ctypes.data)
    return out.astype(bool)


def surround(dirs):
    d = np.ascontiguousarray(dirs, dtype=np.int32)
    return bool(lib().wbo_surround(d.ctypes.data, len(d)))


class Result:
    pass


def run(files, tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1, unit=1.0,
        classify=True):
    """Run the whole path on a list of file images.

    files: list of dicts {records (n,rec_len) uint8, fmt, scale[3], offset[3], min[3], max[3]}
           (min/max = the LAS header's corners, which feed sizeFit, wolkencanvas.cpp:502-519).
    Returns dump text, leaves, tiles (sorted by n), labels in INPUT (concatenated file) order.
    """
    L = lib()
    pts_list, cls_list, corners = [], [], []
    for f in files:
        recs = np.ascontiguousarray(f["records"])
        n, rec_len = recs.shape
        xyz = np.empty((n, 3), dtype=np.int32)
        cls = np.empty(n, dtype=np.uint8)
        ret = np.empty(n, dtype=np.uint8)
        if L.wbo_decode(recs.ctypes.data, n, f["fmt"], rec_len, xyz.ctypes.data, cls.ctypes.data, ret.ctypes.data):
            raise ValueError("bad format")
        pts = np.empty((n, 3), dtype=np.float64)
        L.wbo_coords(xyz.ctypes.data, n, _d3(f["scale"]), _d3(f["offset"]), unit, pts.ctypes.data)
        # threads.cpp:477-530: points whose return number is 0 are dropped iff point 0 has a
        # non-zero return number; otherwise they are kept (return number forced to 1).
        if n and ret[0] != 0:
            keep = ret != 0
        else:
            keep = np.ones(n, dtype=bool)
        f["_keep"] = keep
        pts_list.append(pts[keep])
        cls_list.append(cls[keep])
        corners.append([c * unit for c in f["min"]])
        corners.append([c * unit for c in f["max"]])
    pts_all = np.ascontiguousarray(np.concatenate(pts_list)) if pts_list else np.zeros((0, 3))
    corners = np.ascontiguousarray(np.array(corners, dtype=np.float64))
    res = Result()
    # OctBuffer::put (octree.cpp:620-662) overwrites a stored point whose location is identical:
    # of several points with the same XYZ only one stays in the store (at the position of the first
    # inserted, i.e. the lowest input index in canonical order).  The others are "lost"; here they
    # simply inherit the survivor's label.
    _, first_idx, inverse = np.unique(pts_all, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    rep = first_idx[inverse]                       # input index of each point's representative
    is_rep = rep == np.arange(len(pts_all))
    res.n_duplicates = int((~is_rep).sum())
    res.representative = rep
    rep_rank = np.cumsum(is_rep) - 1                # index among the kept points
    pts = np.ascontiguousarray(pts_all[is_rep])
    n = pts.shape[0]
    center = (C.c_double * 3)()
    side = C.c_double()
    L.wbo_size_fit(corners.ctypes.data, len(corners), center, C.byref(side))
    cube = (C.c_double * 4)()
    L.wbo_bbox_cube(corners.ctypes.data, len(corners), cube)
    res.root_center = tuple(center)
    res.root_side = side.value
    res.cube = tuple(cube)
    spacing, lo, hi = C.c_double(), C.c_int(), C.c_int()
    res.snake_index = L.wbo_snake_set_size(cube[3], tile_size, C.byref(spacing), C.byref(lo), C.byref(hi))
    res.spacing, res.lo, res.hi = spacing.value, lo.value, hi.value
    keys = np.empty(n, dtype=np.uint64)
    order = np.empty(n, dtype=np.uint32)
    L.wbo_sort(pts.ctypes.data, n, center, side.value, keys.ctypes.data, order.ctypes.data)
    res.keys, res.order = keys, order
    res.order_input = np.nonzero(is_rep)[0][order]  # canonical order in terms of ORIGINAL input indices
    cap = max(16, n // 32 + 16)
    leaves = np.zeros(cap, dtype=LEAF_DTYPE)
    nl = L.wbo_leaves(keys.ctypes.data, n, center, side.value, leaves.ctypes.data, cap)
    assert nl <= cap
    res.leaves = leaves[:nl]
    buf = C.create_string_buffer(int(nl) * 120 + 64)
    ln = L.wbo_dump(leaves.ctypes.data, nl, buf, len(buf))
    res.dump = buf.raw[:ln].decode("utf-8")
    sorted_pts = np.ascontiguousarray(pts[order])
    res.points_sorted = sorted_pts
    capt = max(16, 2 * n + 16)
    tiles = np.zeros(capt, dtype=TILE_DTYPE)
    nt = L.wbo_scan(sorted_pts.ctypes.data, n, cube, tile_size, min_hyperboloid_size, tiles.ctypes.data, capt)
    assert nt <= capt
    tiles = tiles[:nt].copy()
    res.tiles_scan = tiles.copy()
    L.wbo_postscan(tiles.ctypes.data, nt, spacing.value)
    res.tiles = tiles
    if classify:
        lab_sorted = np.zeros(n, dtype=np.uint8)
        margins = C.c_uint64()
        L.wbo_classify(sorted_pts.ctypes.data, n, cube, tile_size, max_slope, thickness,
                       tiles.ctypes.data, nt, lab_sorted.ctypes.data, C.byref(margins))
        labels = np.zeros(n, dtype=np.uint8)
        labels[order] = lab_sorted
        res.labels_sorted = lab_sorted
        res.labels_kept = labels                    # one per distinct location
        res.labels = labels[rep_rank[rep]]          # one per input point (duplicates share the survivor's)
        res.margin_count = margins.value
    return res


def point_hyperboloid_sizes(res, tile_size=1.0):
    """Per point of res.points_sorted: hyperboloidSize of the tile that classifies it (NaN = in no tile)."""
    L = lib()
    n = len(res.points_sorted)
    out = np.empty(n, dtype=np.float64)
    tiles = np.ascontiguousarray(res.tiles)
    cube = (C.c_double * 4)(*res.cube)
    L.wbo_point_hyperboloid_sizes(res.points_sorted.ctypes.data, n, cube, tile_size, tiles.ctypes.data, len(tiles),
                                  out.ctypes.data)
    return out


def file_from_cloud(cloud):
    """Adapter from wolkenbase_b200.synth.Cloud to the dict `run` takes."""
    return {"records": cloud.records, "fmt": cloud.fmt, "scale": cloud.scale, "offset": cloud.offset,
            "min": cloud.min_corner, "max": cloud.max_corner}

"""ctypes binding of libwolken_b200.so (include/wolken_b200.h).

Thin by design: every method is one C-ABI call, with the reference phase it replaces named in
the docstring.  The library has no CPU fallback; constructing a Context without the .so or
without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WB_LIB") or os.path.join(_HERE, "libwolken_b200.so")
_LIB = None

WB_RECORDS = 537
WB_LEVELS = 21

LEAF_DTYPE = np.dtype([("first", "<u8"), ("count", "<u4"), ("depth", "<i4"), ("cx", "<f8"), ("cy", "<f8"),
                       ("cz", "<f8"), ("half", "<f8"), ("low", "<f8"), ("high", "<f8")])
TILE_DTYPE = np.dtype([("n", "<i4"), ("ex", "<i4"), ("ey", "<i4"), ("nPoints", "<i4"), ("treeFlags", "<i4"),
                       ("pad_", "<i4"), ("density", "<f8"), ("hyperboloidSize", "<f8"), ("height", "<f8")])


class Geometry(C.Structure):
    _fields_ = [("root_center", C.c_double * 3), ("root_side", C.c_double), ("cube", C.c_double * 4),
                ("spacing", C.c_double), ("radius", C.c_double), ("snake_index", C.c_int32),
                ("snake_lo", C.c_int32), ("snake_hi", C.c_int32), ("pad_", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_points", "n_dropped", "n_duplicates", "n_leaves", "n_tiles_nonempty",
                                          "n_memberships", "n_margin", "n_untiled", "n_second_walk",
                                          "cl_nodes", "cl_chunks", "cl_pairs", "cl_nodes2", "cl_chunks2",
                                          "cl_pairs2", "cl_warps2", "kernel_launches")] + \
               [(k, C.c_double) for k in ("ms_h2d", "ms_decode", "ms_build", "ms_scan", "ms_postscan",
                                          "ms_classify", "ms_d2h", "ms_sort", "ms_leaves", "ms_hier",
                                          "ms_pairs", "ms_classify_kernel", "ms_encode", "ms_encode_d2h",
                                          "ms_classify_order")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ShardStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_own", "own_first", "n_halo_scan", "n_halo_classify", "grid_cells",
                                          "bytes_sent", "bytes_received")] + \
               [(k, C.c_double) for k in ("por_max", "ms_setup", "ms_select", "ms_exchange", "ms_build_scan", "ms_scan",
                                          "ms_grid", "ms_postscan", "ms_build_classify", "ms_assign", "ms_classify")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_u64p = C.POINTER(C.c_uint64)
ALL_GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
ALL_TO_ALL_V_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, _u64p, _u64p, C.c_void_p, _u64p, _u64p)
ALL_REDUCE_MAX_U8_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)


class CensusResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("pad_", C.c_int32), ("n_stored", C.c_uint64), ("max_point", C.c_uint64),
                ("n_missing", C.c_uint64), ("n_duplicate", C.c_uint64)]


class CommOps(C.Structure):
    _fields_ = [("user", C.c_void_p), ("all_gather", ALL_GATHER_FN), ("all_to_all_v", ALL_TO_ALL_V_FN),
                ("all_reduce_max_u8", ALL_REDUCE_MAX_U8_FN)]


class OutSpec(C.Structure):
    _fields_ = [("format", C.c_int32), ("rec_len", C.c_int32), ("n_classes", C.c_int32), ("separate", C.c_int32),
                ("scale", C.c_double * 3), ("offset", C.c_double * 3), ("unit", C.c_double),
                ("classes", C.c_uint8 * 256)]


class FileStats(C.Structure):
    _fields_ = [("n_points", C.c_uint64 * 16), ("imin", C.c_int32 * 3), ("imax", C.c_int32 * 3),
                ("pad_", C.c_int32 * 2)]


SHAPE_DTYPE = np.dtype([("type", "<i4"), ("pad_", "<i4"), ("p", "<f8", (6,))])
SPHERE, PARABOLOID, HYPERBOLOID, CYLINDER, COLUMN = range(5)


def shapes(kind, params):
    """(n, k) parameter rows -> wb_shape array (parameters = the reference constructors' arguments)."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    out = np.zeros(len(params), dtype=SHAPE_DTYPE)
    out["type"] = kind
    out["p"][:, :params.shape[1]] = params
    return out


class WolkenError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; fails loudly if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise WolkenError("libwolken_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        vp, u64, dp = C.c_void_p, C.c_uint64, C.POINTER(C.c_double)
        sig = {
            "wb_create": [C.c_int, C.POINTER(vp)],
            "wb_destroy": [vp],
            "wb_reserve": [vp, u64],
            "wb_clear": [vp],
            "wb_set_params": [vp, C.c_double, C.c_double, C.c_double, C.c_double],
            "wb_add_extent": [vp, dp, dp],
            "wb_add_las": [vp, vp, u64, C.c_int, C.c_int, dp, dp, C.c_double],
            "wb_add_las_device": [vp, vp, u64, C.c_int, C.c_int, dp, dp, C.c_double],
            "wb_add_las_file": [vp, C.c_char_p, u64, u64, C.c_int, C.c_int, dp, dp, C.c_double],
            "wb_add_points_device": [vp, vp, vp, vp, vp, u64, dp, dp, C.c_double],
            "wb_set_own_range": [vp, u64, u64],
            "wb_max_hyperboloid_size": [vp, dp],
            "wb_assign": [vp],
            "wb_set_geometry": [vp, dp, C.c_double, dp],
            "wb_get_geometry": [vp, C.POINTER(Geometry)],
            "wb_build": [vp],
            "wb_num_leaves": [vp, C.POINTER(u64)],
            "wb_get_leaves": [vp, vp, u64],
            "wb_get_order": [vp, vp, vp],
            "wb_get_decoded": [vp, vp, vp, vp, vp],
            "wb_get_points_sorted": [vp, vp, vp, vp],
            "wb_scan": [vp],
            "wb_postscan": [vp],
            "wb_num_tiles": [vp, C.POINTER(u64)],
            "wb_get_tiles": [vp, vp, u64],
            "wb_set_tiles": [vp, vp, u64],
            "wb_classify": [vp],
            "wb_get_labels": [vp, vp],
            "wb_count_classes": [vp, vp],
            "wb_patch_records": [vp, vp, u64, u64, C.c_int, C.c_int],
            "wb_test_math": [vp, u64, vp, vp, vp, vp, vp],
            "wb_run": [vp],
            "wb_get_stats": [vp, C.POINTER(Stats)],
            "wb_sync": [vp],
            "wb_mark": [vp, C.c_int],
            "wb_set_return_zero_rule": [vp, C.c_int],
            "wb_mark_elapsed": [vp, C.c_int, C.c_int, dp],
            "wb_host_alloc": [C.POINTER(vp), u64],
            "wb_host_free": [vp],
            "wb_size_fit": [vp, C.c_int, dp, dp],
            "wb_bbox_cube": [vp, C.c_int, dp],
            "wb_bound_rect": [vp, C.c_int, dp],
            "wb_snake_set_size": [C.c_double, C.c_double, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
            "wb_ldecimal": [C.c_double, C.c_char_p, C.c_int],
            "wb_format_dump": [vp, u64, C.c_char_p, u64],
            "wb_keep_records": [vp, C.c_int],
            "wb_census": [vp, C.POINTER(CensusResult), vp, C.c_uint64],
            "wb_device_count": [C.POINTER(C.c_int)],
            "wb_set_window": [vp, C.c_double, C.c_double],
            "wb_num_loaded": [vp, C.POINTER(C.c_uint64)],
            "wb_leaf_class_counts": [vp, vp, C.c_int, C.c_int, vp],
            "wb_encode": [vp, C.POINTER(OutSpec), vp, vp, C.c_uint32, vp, u64, vp],
            "wb_get_duplicates": [vp, vp, vp, u64],
            "wb_write_encoded": [vp, C.c_int, u64, u64, u64],
            "wb_query_batch": [vp, vp, u64, vp, vp, vp],
            "wb_query_points": [vp, vp, u64, C.POINTER(u64), vp, vp, vp, vp, vp],
            "wb_comm_get_id": [vp],
            "wb_comm_init": [vp, vp, C.c_int, C.c_int, C.POINTER(vp)],
            "wb_local_group_create": [C.c_int, C.POINTER(vp)],
            "wb_comm_init_local": [vp, vp, C.c_int, C.POINTER(vp)],
            "wb_comm_init_custom": [vp, C.POINTER(CommOps), C.c_int, C.c_int, C.POINTER(vp)],
            "wb_shard_run": [vp, vp],
            "wb_shard_get_labels": [vp, vp],
            "wb_shard_get_stats": [vp, C.POINTER(ShardStats)],
            "wb_set_labels": [vp, vp],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int
        L.wb_destroy.restype = None
        for name in ("wb_comm_destroy", "wb_local_group_destroy"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = None
        L.wb_last_error.argtypes = [vp]
        L.wb_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


EXPORTS = ["wb_create", "wb_destroy", "wb_last_error", "wb_reserve", "wb_clear", "wb_set_params", "wb_add_extent",
           "wb_add_las", "wb_add_las_device", "wb_set_geometry", "wb_get_geometry", "wb_build", "wb_num_leaves",
           "wb_get_leaves", "wb_get_order", "wb_get_decoded", "wb_scan", "wb_postscan", "wb_num_tiles",
           "wb_get_tiles", "wb_set_tiles", "wb_classify", "wb_get_labels", "wb_count_classes", "wb_patch_records",
           "wb_run", "wb_get_stats", "wb_sync", "wb_host_alloc", "wb_host_free", "wb_size_fit", "wb_bbox_cube",
           "wb_snake_set_size", "wb_ldecimal", "wb_format_dump", "wb_add_points_device",
           "wb_set_own_range", "wb_max_hyperboloid_size",
           "wb_assign", "wb_get_points_sorted", "wb_test_math", "wb_bound_rect", "wb_keep_records",
           "wb_leaf_class_counts", "wb_encode", "wb_get_duplicates", "wb_add_las_file", "wb_write_encoded", "wb_query_batch",
           "wb_query_points", "wb_mark", "wb_mark_elapsed", "wb_set_return_zero_rule",
           "wb_comm_get_id", "wb_comm_init", "wb_local_group_create", "wb_local_group_destroy", "wb_comm_init_local",
           "wb_comm_init_custom", "wb_comm_destroy", "wb_shard_run", "wb_shard_get_labels", "wb_shard_get_stats",
           "wb_set_labels", "wb_device_count", "wb_census", "wb_set_window", "wb_num_loaded"]


def _d(v):
    return (C.c_double * len(v))(*[float(x) for x in v])


def ldecimal(x):
    buf = C.create_string_buffer(64)
    lib().wb_ldecimal(x, buf, 64)
    return buf.value.decode()


class PinnedBuffer:
    """Pinned host memory (cudaHostAlloc) viewed as a numpy uint8 array."""

    def __init__(self, nbytes):
        self._p = C.c_void_p()
        if lib().wb_host_alloc(C.byref(self._p), nbytes) != 0:
            raise WolkenError("cudaHostAlloc(%d) failed" % nbytes)
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), (nbytes,))

    def free(self):
        if self._p:
            self.array = None
            lib().wb_host_free(self._p)
            self._p = C.c_void_p()


class Context:
    """One GPU's worker-pool replacement (threads.h:92-113 -> wb_*)."""

    def __init__(self, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.wb_create(device, C.byref(self._h))
        if rc != 0:
            raise WolkenError("wb_create(device=%d) failed with %d: no usable CUDA device "
                              "(this library has no CPU path)" % (device, rc))
        self.device = device

    def close(self):
        if self._h:
            self._L.wb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise WolkenError("%s (code %d)" % (self._L.wb_last_error(self._h).decode(), rc))

    # ---- configuration
    def set_params(self, tile_size=1.0, max_slope=1.0, thickness=0.0, min_hyperboloid_size=0.1):
        self._ck(self._L.wb_set_params(self._h, tile_size, max_slope, thickness, min_hyperboloid_size))

    def reserve(self, n):
        self._ck(self._L.wb_reserve(self._h, n))

    def clear(self):
        self._ck(self._L.wb_clear(self._h))

    # ---- read
    def add_extent(self, mn, mx):
        self._ck(self._L.wb_add_extent(self._h, _d(mn), _d(mx)))

    def add_las(self, records, fmt, scale, offset, unit=1.0):
        """ACT_READ (threads.cpp:477-566): records is an (n, rec_len) uint8 host array."""
        assert records.dtype == np.uint8 and records.ndim == 2 and records.flags.c_contiguous
        self._ck(self._L.wb_add_las(self._h, records.ctypes.data, records.shape[0], fmt, records.shape[1],
                                    _d(scale), _d(offset), unit))

    def add_las_file(self, path, point_offset, n, fmt, rec_len, scale, offset, unit=1.0):
        """ACT_READ from the file itself: pread on worker threads -> pinned ring -> H2D -> decode."""
        self._ck(self._L.wb_add_las_file(self._h, os.fsencode(path), point_offset, n, fmt, rec_len,
                                         _d(scale), _d(offset), unit))

    def add_las_device(self, dptr, n, fmt, rec_len, scale, offset, unit=1.0):
        self._ck(self._L.wb_add_las_device(self._h, dptr, n, fmt, rec_len, _d(scale), _d(offset), unit))

    def add_cloud(self, cloud, unit=1.0):
        """Convenience for synth.Cloud: header corners + records."""
        self.add_extent([c * unit for c in cloud.min_corner], [c * unit for c in cloud.max_corner])
        self.add_las(cloud.records, cloud.fmt, cloud.scale, cloud.offset, unit)

    def add_points_device(self, dx, dy, dz, dcls, n, scale, offset, unit=1.0):
        self._ck(self._L.wb_add_points_device(self._h, dx, dy, dz, dcls, n, _d(scale), _d(offset), unit))

    def set_own_range(self, first, end):
        self._ck(self._L.wb_set_own_range(self._h, first, end))

    # ---- several GPUs (wb_shard_run)
    def shard_run(self, comm):
        """The whole path for this rank's strip; collective over the ranks of `comm` (a Comm)."""
        self._ck(self._L.wb_shard_run(self._h, comm._h))

    def shard_labels(self, n, out=None):
        lab = out if out is not None else np.empty(n, dtype=np.uint8)
        self._ck(self._L.wb_shard_get_labels(self._h, lab.ctypes.data))
        return lab

    def shard_stats(self):
        s = ShardStats()
        self._ck(self._L.wb_shard_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def set_labels(self, labels):
        labels = np.ascontiguousarray(labels, dtype=np.uint8)
        self._ck(self._L.wb_set_labels(self._h, labels.ctypes.data))

    def max_hyperboloid_size(self):
        v = C.c_double()
        self._ck(self._L.wb_max_hyperboloid_size(self._h, C.byref(v)))
        return v.value

    def assign(self):
        self._ck(self._L.wb_assign(self._h))

    def set_geometry(self, root_center, root_side, cube):
        self._ck(self._L.wb_set_geometry(self._h, _d(root_center), root_side, _d(cube)))

    def geometry(self):
        g = Geometry()
        self._ck(self._L.wb_get_geometry(self._h, C.byref(g)))
        return g

    # ---- phases
    def build(self):
        self._ck(self._L.wb_build(self._h))

    def scan(self):
        self._ck(self._L.wb_scan(self._h))

    def postscan(self):
        self._ck(self._L.wb_postscan(self._h))

    def classify(self):
        self._ck(self._L.wb_classify(self._h))

    def run(self):
        self._ck(self._L.wb_run(self._h))

    def sync(self):
        self._ck(self._L.wb_sync(self._h))

    def mark(self, slot):
        """CUDA event `slot` on the library's compute stream, behind all work issued so far."""
        self._ck(self._L.wb_mark(self._h, slot))

    def mark_elapsed(self, a, b):
        """Device milliseconds between marks a and b (waits for b)."""
        ms = C.c_double()
        self._ck(self._L.wb_mark_elapsed(self._h, a, b, C.byref(ms)))
        return ms.value

    # ---- results
    def leaves(self):
        n = C.c_uint64()
        self._ck(self._L.wb_num_leaves(self._h, C.byref(n)))
        out = np.zeros(n.value, dtype=LEAF_DTYPE)
        if n.value:
            self._ck(self._L.wb_get_leaves(self._h, out.ctypes.data, n.value))
        return out

    def dump(self):
        """octStore.dump text (octree.cpp:888-891)."""
        lv = self.leaves()
        buf = C.create_string_buffer(len(lv) * 128 + 64)
        ln = self._L.wb_format_dump(lv.ctypes.data, len(lv), buf, len(buf))
        if ln < 0:
            raise WolkenError("wb_format_dump failed")
        return buf.raw[:ln].decode("utf-8")

    # ---- output records (ACT_WRITE)
    def set_return_zero_rule(self, keep_all):
        """keep_all=True: store records with return number 0 too (wolkencli.cpp:104-108 callers)."""
        self._ck(self._L.wb_set_return_zero_rule(self._h, 1 if keep_all else 0))

    def keep_records(self, keep=True):
        """Before add_las: keep the raw records in device memory for encode()."""
        self._ck(self._L.wb_keep_records(self._h, 1 if keep else 0))

    def leaf_class_counts(self, classes=None):
        """counts[leaf, k] = points of bucket `leaf` whose class is classes[k]; classes=None: all together."""
        nl = len(self.leaves())
        sep = classes is not None
        cl = np.ascontiguousarray(classes if sep else [0], dtype=np.uint8)
        out = np.zeros((nl, len(cl)), dtype=np.uint32)
        self._ck(self._L.wb_leaf_class_counts(self._h, cl.ctypes.data, len(cl), 1 if sep else 0, out.ctypes.data))
        return out

    def encode(self, fmt, rec_len, scale, offset, dest, file_of, n_files, out_bytes, classes=None, unit=1.0,
               fetch=True):
        """wb_encode: returns (records as a uint8 array of out_bytes, per-file stats)."""
        spec = OutSpec()
        spec.format, spec.rec_len = fmt, rec_len
        spec.separate = 1 if classes is not None else 0
        cl = list(classes) if classes is not None else [0]
        spec.n_classes = len(cl)
        for k, c in enumerate(cl):
            spec.classes[k] = int(c)
        for k in range(3):
            spec.scale[k], spec.offset[k] = float(scale[k]), float(offset[k])
        spec.unit = unit
        dest = np.ascontiguousarray(dest, dtype=np.uint64)
        file_of = np.ascontiguousarray(file_of, dtype=np.uint32)
        out = np.zeros(out_bytes if fetch else 0, dtype=np.uint8)
        stats = (FileStats * n_files)()
        self._ck(self._L.wb_encode(self._h, C.byref(spec), dest.ctypes.data, file_of.ctypes.data, n_files,
                                   out.ctypes.data if fetch else None, out_bytes, stats))
        return out, [{"n_points": list(s.n_points), "imin": list(s.imin), "imax": list(s.imax)} for s in stats]

    def write_encoded(self, fd, file_pos, arena_off, nbytes):
        """Stream records left on the device by encode(..., fetch=False) into an open file."""
        self._ck(self._L.wb_write_encoded(self._h, fd, file_pos, arena_off, nbytes))

    # ---- store queries (OctStore::countPointsIn / hiLoPointsIn / pointsIn)
    def query_batch(self, shp):
        shp = np.ascontiguousarray(shp, dtype=SHAPE_DTYPE)
        n = len(shp)
        count = np.zeros(n, dtype=np.uint64)
        lo = np.zeros(n)
        hi = np.zeros(n)
        self._ck(self._L.wb_query_batch(self._h, shp.ctypes.data, n, count.ctypes.data, lo.ctypes.data, hi.ctypes.data))
        return count, lo, hi

    def query_points(self, shp, cap=None):
        shp = np.ascontiguousarray(shp, dtype=SHAPE_DTYPE)
        assert len(shp) == 1
        n = C.c_uint64()
        if cap is None:
            self._ck(self._L.wb_query_points(self._h, shp.ctypes.data, 0, C.byref(n), None, None, None, None, None))
            cap = n.value
        pos = np.zeros(cap, dtype=np.uint32)
        idx = np.zeros(cap, dtype=np.uint32)
        x, y, z = np.zeros(cap), np.zeros(cap), np.zeros(cap)
        self._ck(self._L.wb_query_points(self._h, shp.ctypes.data, cap, C.byref(n), pos.ctypes.data, idx.ctypes.data,
                                         x.ctypes.data, y.ctypes.data, z.ctypes.data))
        m = min(cap, n.value)
        return n.value, pos[:m], idx[:m], np.stack([x[:m], y[:m], z[:m]], axis=1)

    def duplicates(self):
        n = self.stats()["n_duplicates"]
        dup = np.zeros(n, dtype=np.uint32)
        rep = np.zeros(n, dtype=np.uint32)
        self._ck(self._L.wb_get_duplicates(self._h, dup.ctypes.data, rep.ctypes.data, n))
        return dup, rep

    def order(self, n):
        order = np.empty(n, dtype=np.uint32)
        keys = np.empty(n, dtype=np.uint64)
        self._ck(self._L.wb_get_order(self._h, order.ctypes.data, keys.ctypes.data))
        return order, keys

    def points_sorted(self, n):
        x = np.empty(n); y = np.empty(n); z = np.empty(n)
        self._ck(self._L.wb_get_points_sorted(self._h, x.ctypes.data, y.ctypes.data, z.ctypes.data))
        return x, y, z

    def decoded(self, n):
        x = np.empty(n, dtype=np.int32)
        y = np.empty(n, dtype=np.int32)
        z = np.empty(n, dtype=np.int32)
        c = np.empty(n, dtype=np.uint8)
        self._ck(self._L.wb_get_decoded(self._h, x.ctypes.data, y.ctypes.data, z.ctypes.data, c.ctypes.data))
        return x, y, z, c

    def tiles(self):
        n = C.c_uint64()
        self._ck(self._L.wb_num_tiles(self._h, C.byref(n)))
        out = np.zeros(n.value, dtype=TILE_DTYPE)
        if n.value:
            self._ck(self._L.wb_get_tiles(self._h, out.ctypes.data, n.value))
        return out

    def set_tiles(self, tiles):
        t = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        self._ck(self._L.wb_set_tiles(self._h, t.ctypes.data, len(t)))

    def labels(self, n, out=None):
        lab = out if out is not None else np.empty(n, dtype=np.uint8)
        self._ck(self._L.wb_get_labels(self._h, lab.ctypes.data))
        return lab

    def set_window(self, x_lo, x_hi):
        """Keep only records with x in [x_lo, x_hi) from the following add_las* calls (wb_set_window)."""
        self._ck(self._L.wb_set_window(self._h, float(x_lo), float(x_hi)))

    def num_loaded(self):
        n = C.c_uint64()
        self._ck(self._L.wb_num_loaded(self._h, C.byref(n)))
        return int(n.value)

    def census(self, cap=64):
        """censusPoints() (testpattern.cpp:84-123) over the store: dict with status, max_point, n_missing, n_duplicate
        and up to `cap` missing point numbers.  Needs keep_records() before the records were added."""
        r = CensusResult()
        miss = np.zeros(max(1, cap), dtype=np.uint64)
        self._ck(self._L.wb_census(self._h, C.byref(r), miss.ctypes.data if cap else None, cap))
        return {"status": int(r.status), "n_stored": int(r.n_stored), "max_point": int(r.max_point),
                "n_missing": int(r.n_missing), "n_duplicate": int(r.n_duplicate),
                "missing": miss[:min(cap, int(r.n_missing))].tolist()}

    def count_classes(self):
        c = np.zeros(256, dtype=np.uint64)
        self._ck(self._L.wb_count_classes(self._h, c.ctypes.data))
        return c

    def patch_records(self, records, fmt, first=0):
        self._ck(self._L.wb_patch_records(self._h, records.ctypes.data, first, records.shape[0], fmt, records.shape[1]))

    def test_math(self, y, x):
        y = np.ascontiguousarray(y, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        a = np.empty(len(x), dtype=np.int32)
        h = np.empty(len(x), dtype=np.float64)
        s = np.empty(len(x), dtype=np.int32)
        self._ck(self._L.wb_test_math(self._h, len(x), y.ctypes.data, x.ctypes.data, a.ctypes.data, h.ctypes.data,
                                      s.ctypes.data))
        return a, h, s

    def stats(self):
        s = Stats()
        self._ck(self._L.wb_get_stats(self._h, C.byref(s)))
        return s.as_dict()


def bind_to_gpu_numa(device):
    """Bind this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that pinned buffers allocated
    afterwards — the caller's records, the file reader's ring — are local to the GPU's PCIe root.  What `numactl
    --cpunodebind` does for a one-process-per-GPU job; a no-op where the topology is not exposed.  Returns the node
    or -1."""
    import os
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return -1
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return -1
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return -1


class LocalGroup:
    """The ranks of one process (one host thread each) exchanging without NCCL (wb_comm_init_local)."""

    def __init__(self, world):
        self._L = lib()
        self._h = C.c_void_p()
        if self._L.wb_local_group_create(world, C.byref(self._h)) != 0:
            raise WolkenError("wb_local_group_create(%d) failed" % world)
        self.world = world

    def close(self):
        if self._h:
            self._L.wb_local_group_destroy(self._h)
            self._h = C.c_void_p()


class Comm:
    """One rank's end of the halo exchange (wb_comm).  Build with Comm.nccl, Comm.local or Comm.custom."""

    def __init__(self, ctx, handle, keep=None):
        self._L, self._h, self.ctx, self._keep = lib(), handle, ctx, keep

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        if lib().wb_comm_get_id(buf) != 0:
            raise WolkenError("wb_comm_get_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    @classmethod
    def nccl(cls, ctx, unique_id, rank, world):
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        ctx._ck(lib().wb_comm_init(ctx._h, buf, rank, world, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def local(cls, ctx, group, rank):
        h = C.c_void_p()
        ctx._ck(lib().wb_comm_init_local(ctx._h, group._h, rank, C.byref(h)))
        return cls(ctx, h, keep=group)

    @classmethod
    def custom(cls, ctx, all_gather, all_to_all_v, all_reduce_max_u8, rank, world):
        """Python callables (device pointers as ints; return 0 on success) stand in for the transport."""
        ops = CommOps(None, ALL_GATHER_FN(all_gather), ALL_TO_ALL_V_FN(all_to_all_v),
                      ALL_REDUCE_MAX_U8_FN(all_reduce_max_u8))
        h = C.c_void_p()
        ctx._ck(lib().wb_comm_init_custom(ctx._h, C.byref(ops), rank, world, C.byref(h)))
        return cls(ctx, h, keep=ops)

    def close(self):
        if self._h:
            self._L.wb_comm_destroy(self._h)
            self._h = C.c_void_p()

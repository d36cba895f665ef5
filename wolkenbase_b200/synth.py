"""Synthetic LAS clouds for the BASELINE.json configs (ctypes front end of csrc/synth.c).

The generator is integer-only and counter-based, so a (scene, seed, n) triple names the same
bytes everywhere.  Scenes follow the reference's testpattern.cpp (laserize(),
testpattern.cpp:166-190: return 1 of 1, gpsTime = point index).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SynthDesc(C.Structure):
    _fields_ = [("scene", C.c_int32), ("fmt", C.c_int32), ("scale", C.c_double),
                ("offset", C.c_double * 3), ("n_points", C.c_uint64),
                ("grid_nx", C.c_uint64), ("grid_ny", C.c_uint64),
                ("cell_ticks", C.c_uint64), ("extent_ticks", C.c_uint64)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libwb_synth.so")
        if not os.path.exists(path):
            raise RuntimeError("libwb_synth.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = C.CDLL(path)
        _LIB.wb_synth_describe.argtypes = [C.c_int, C.c_uint64, C.POINTER(SynthDesc)]
        _LIB.wb_synth_generate.argtypes = [C.POINTER(SynthDesc), C.c_uint64, C.c_uint64, C.c_uint64,
                                           C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        _LIB.wb_synth_generate_terrestrial.argtypes = [C.POINTER(SynthDesc), C.c_uint64, C.c_uint64,
                                                       C.c_void_p, C.c_void_p]
        _LIB.wb_synth_header.argtypes = [C.POINTER(SynthDesc), C.c_uint64, C.c_void_p, C.c_void_p]
        _LIB.wb_synth_record_length.argtypes = [C.c_int]
    return _LIB


class Cloud:
    """One LAS file image in memory: header bytes + (n, rec_len) uint8 records."""

    def __init__(self, desc, header, records, bbox):
        self.desc = desc
        self.header = header
        self.records = records
        self.bbox = bbox
        self.fmt = desc.fmt
        self.rec_len = records.shape[1]
        self.n = records.shape[0]
        self.scale = (desc.scale,) * 3
        self.offset = tuple(desc.offset)

    @property
    def min_corner(self):
        return tuple(self.offset[i] + self.scale[i] * float(self.bbox[i]) for i in range(3))

    @property
    def max_corner(self):
        return tuple(self.offset[i] + self.scale[i] * float(self.bbox[3 + i]) for i in range(3))

    def write(self, path):
        with open(path, "wb") as f:
            f.write(self.header.tobytes())
            f.write(self.records.tobytes())

    def ints(self):
        """(n,3) int32 view of X,Y,Z."""
        return np.ascontiguousarray(self.records[:, :12]).view(np.int32).reshape(self.n, 3)


def describe(scene, n_points):
    d = SynthDesc()
    if lib().wb_synth_describe(scene, n_points, C.byref(d)) != 0:
        raise ValueError("unknown scene %r" % scene)
    return d


def generate(scene, n_points, seed=1, region=None, gps_base=0, out=None):
    """Generate about n_points of `scene`.  `region` = (cx0, cy0, ncx, ncy) picks a sub-rectangle
    of the jitter grid (multi-file / multi-GPU sharding); default is the whole scene."""
    L = lib()
    d = describe(scene, n_points)
    rec_len = L.wb_synth_record_length(d.fmt)
    if d.scene == 4:
        n = d.n_points
    else:
        if region is None:
            region = (0, 0, d.grid_nx, d.grid_ny)
        n = region[2] * region[3]
    recs = out if out is not None else np.empty((n, rec_len), dtype=np.uint8)
    assert recs.shape == (n, rec_len) and recs.flags.c_contiguous
    bbox = np.zeros(6, dtype=np.int32)
    if d.scene == 4:
        rc = L.wb_synth_generate_terrestrial(C.byref(d), seed, gps_base, recs.ctypes.data, bbox.ctypes.data)
    else:
        rc = L.wb_synth_generate(C.byref(d), seed, region[0], region[1], region[2], region[3],
                                 gps_base, recs.ctypes.data, bbox.ctypes.data)
    if rc != 0:
        raise RuntimeError("wb_synth_generate failed: %d" % rc)
    hdr = np.zeros(375, dtype=np.uint8)
    size = L.wb_synth_header(C.byref(d), n, bbox.ctypes.data, hdr.ctypes.data)
    return Cloud(d, hdr[:size].copy(), recs, bbox)


def with_duplicates(cloud, n_dup, seed=0):
    """Copy of `cloud` in which n_dup records take the XYZ of other records (own attributes and
    gpsTime kept): the reference stores one point per location (OctBuffer::put, octree.cpp:620-662)."""
    recs = cloud.records.copy()
    rng = np.random.default_rng(seed)
    src = rng.integers(0, cloud.n, n_dup)
    dst = rng.integers(0, cloud.n, n_dup)
    recs[dst, :12] = recs[src, :12]
    return Cloud(cloud.desc, cloud.header, recs, cloud.bbox)

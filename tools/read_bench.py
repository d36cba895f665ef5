"""Time the three ways records reach the GPU: wb_add_las from pageable memory, from an mmap of the
file, and wb_add_las_file (pread threads -> pinned ring).  Usage: python tools/read_bench.py [points]"""
import mmap
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wolkenbase_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cloud = synth.generate(2, n, seed=2)
path = "/tmp/wb_read_bench.las"
cloud.write(path)
hdr = len(cloud.header.tobytes())
ctx = api.Context(0)
ctx.reserve(cloud.n)


def timed(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        ctx.clear()
        ctx.add_extent(cloud.min_corner, cloud.max_corner)
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    gb = cloud.n * cloud.rec_len / 1e9
    print("%-28s %.3f s  %.2f GB/s  %.1f Mpts/s" % (label, best, gb / best, cloud.n / best / 1e6))


timed("add_las pageable array", lambda: ctx.add_las(cloud.records, cloud.fmt, cloud.scale, cloud.offset))
f = open(path, "rb")
mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
arr = np.frombuffer(mm, dtype=np.uint8, count=cloud.n * cloud.rec_len, offset=hdr).reshape(cloud.n, cloud.rec_len)
timed("add_las mmap of the file", lambda: ctx.add_las(arr, cloud.fmt, cloud.scale, cloud.offset))
timed("add_las_file", lambda: ctx.add_las_file(path, hdr, cloud.n, cloud.fmt, cloud.rec_len, cloud.scale, cloud.offset))
pin = api.PinnedBuffer(cloud.n * cloud.rec_len)
pin.array[:] = cloud.records.reshape(-1)
parr = pin.array.reshape(cloud.n, cloud.rec_len)
timed("add_las pinned memory", lambda: ctx.add_las(parr, cloud.fmt, cloud.scale, cloud.offset))
os.remove(path)

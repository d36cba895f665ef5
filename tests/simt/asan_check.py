"""Run the emulated kernels under AddressSanitizer (a CPU-side memcheck of the kernel sources: every global and
shared-memory access of the decode, sort, membership, tile-scan, postscan and classify kernels on ragged and tiny
inputs).  Started by tests/test_simt_blocks.py with LD_PRELOAD=libasan.so; prints ASAN-CHECK-OK at the end."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import emul  # noqa: E402
from oracle import wb_oracle as O  # noqa: E402
from wolkenbase_b200 import synth  # noqa: E402

ASAN_LIB = os.path.join(HERE, "libwb_simt_asan.so")


def build_asan():
    subprocess.check_call(["g++", "-std=gnu++17", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer",
                           "-ffp-contract=off", "-fPIC", "-shared", "-Ishim", "-I.", "-Wno-unused-function",
                           "-o", ASAN_LIB, "classify_emul.cpp"], cwd=HERE)


def main():
    emul.build = lambda variant="", out="libwb_simt.so": ASAN_LIB
    for scene, n in [(2, 4000), (5, 1500), (2, 1), (2, 33), (2, 537), (4, 600)]:
        cloud = synth.generate(scene, n, seed=n)
        res = O.run([O.file_from_cloud(cloud)])
        tiles, lab = emul.scan_classify(res.points_sorted, res.cube)
        assert (lab == res.labels_sorted).all() and len(tiles) == len(res.tiles)
        for mis in (0, 7, 15):
            emul.decode(cloud.records, cloud.fmt, misalign=mis)
    rng = np.random.default_rng(1)
    for n in (1, 255, 3072, 3073, 20000):
        keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
        k, _ = emul.radix_sort(keys, np.arange(n, dtype=np.uint32))
        assert (k == np.sort(keys)).all()
    print("ASAN-CHECK-OK")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build_asan()
    else:
        main()

"""Kernels whose threads meet at __syncthreads, under the block-level SIMT emulator (tests/simt): the stable LSD
radix sort of wb_sort.cuh (upsweep histogram, table scan, downsweep with __match_any_sync ranking staged through
shared memory) against numpy's stable argsort, and the LAS decode kernel against the oracle's decode (las.cpp:735-820).
CPU-side checks of the kernel sources; the GPU tests stay the parity proof of the nvcc build."""
import os
import sys

import numpy as np
import pytest

from oracle import wb_oracle as O
from wolkenbase_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import emul  # noqa: E402


@pytest.mark.parametrize("n,bits", [(1, 64), (31, 64), (3072, 16), (3073, 24), (40000, 64), (20000, 8), (6145, 40)])
def test_emulated_radix_sort_is_a_stable_sort(n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    keys[::5] = keys[0]                                  # many ties: the order of equal keys must survive
    vals = np.arange(n, dtype=np.uint32)
    k, v = emul.radix_sort(keys, vals, 0, (bits + 7) // 8 * 8)
    order = np.argsort(keys, kind="stable")
    assert (k == keys[order]).all() and (v == vals[order]).all()


def test_emulated_sort_gives_the_canonical_order():
    """Morton keys of a real cloud: the emulated sort reproduces the oracle's canonical order (key, input index)."""
    cloud = synth.generate(2, 12000, seed=8)
    res = O.run([O.file_from_cloud(cloud)], classify=False)
    keys = np.empty(cloud.n, dtype=np.uint64)
    keys[res.order] = res.keys                           # keys in input order
    k, v = emul.radix_sort(keys, np.arange(cloud.n, dtype=np.uint32), 0, 64)
    assert (k == res.keys).all() and (v == res.order).all()


@pytest.mark.parametrize("scene,n", [(2, 5000), (3, 3000), (5, 2500), (4, 700)])
@pytest.mark.parametrize("misalign", [0, 5, 12])
def test_emulated_decode_matches_oracle(scene, n, misalign):
    """Formats 1 (28 B), 6 (30 B), 3 (34 B), every attribute byte random, records starting anywhere relative to a
    16-byte boundary; both return-number rules (threads.cpp:485-500, 527-530)."""
    cloud = synth.generate(scene, n, seed=scene)
    recs = cloud.records.copy()
    rng = np.random.default_rng(scene * 16 + misalign)
    recs[:, 12:] = rng.integers(0, 256, recs[:, 12:].shape, dtype=np.uint8)
    xyz = np.empty((cloud.n, 3), dtype=np.int32)
    oc, orr = np.empty(cloud.n, dtype=np.uint8), np.empty(cloud.n, dtype=np.uint8)
    assert O.lib().wbo_decode(recs.ctypes.data, cloud.n, cloud.fmt, cloud.rec_len, xyz.ctypes.data, oc.ctypes.data,
                              orr.ctypes.data) == 0
    x, y, z, cls, ret, dropped = emul.decode(recs, cloud.fmt, drop_zeros=False, misalign=misalign)
    assert (x == xyz[:, 0]).all() and (y == xyz[:, 1]).all() and (z == xyz[:, 2]).all() and (cls == oc).all()
    assert (ret == np.where(orr == 0, 1, orr)).all() and dropped == 0
    x, y, z, cls, ret, dropped = emul.decode(recs, cloud.fmt, drop_zeros=True, misalign=misalign)
    assert (x == xyz[:, 0]).all() and (cls == oc).all() and (ret == orr).all() and dropped == int((orr == 0).sum())


def test_emulated_kernels_are_clean_under_address_sanitizer():
    """The same kernels, compiled with -fsanitize=address, on ragged and tiny inputs: no out-of-bounds access to a
    global or shared array (what compute-sanitizer's memcheck looks for on the GPU, here without one)."""
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt")
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not installed")
    subprocess.check_call([sys.executable, os.path.join(here, "asan_check.py"), "build"])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    out = subprocess.run([sys.executable, os.path.join(here, "asan_check.py")], capture_output=True, text=True, env=env,
                         timeout=600)
    assert "ASAN-CHECK-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    assert "AddressSanitizer" not in out.stderr, out.stderr[-4000:]

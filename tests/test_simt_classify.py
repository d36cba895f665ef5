"""The classify kernels' own source (wolkenbase_b200/csrc/wb_kernels.cuh), compiled as host C++ and run warp by
warp under the SIMT emulator of tests/simt, against the oracle's labels.  This checks the traversal logic —
pruning, sector occupancy, the exact second walk, the single-precision shortcuts — on the CPU; the GPU tests remain
the parity proof of the nvcc build.  Variants behind WB_CL_* macros (the round-1 kernel, each round-2 change
alone, the shipped combination) are all held to the same standard here."""
import os
import sys

import numpy as np
import pytest

from oracle import wb_oracle as O
from wolkenbase_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import emul  # noqa: E402

CASES = [(2, 5000, 3, {}), (1, 5000, 1, {}), (5, 5000, 5, {}), (4, 5000, 4, {}), (3, 4000, 7, {}),
         (2, 4000, 9, {"max_slope": 0.5, "thickness": 0.05, "tile_size": 2.0})]
OFF = "-DWB_CL_XWANTS=0 -DWB_CL_REFILTER=0 -DWB_CL_COMPACT2=0 -DWB_CL_TRANSPOSE=0"     # the round-1 kernel (all on since round 2)
VARIANTS = [(OFF, "libwb_simt_r1.so"),
            (OFF + " -DWB_CL_FSECTOR=0 -DWB_CL_FREACH=0 -DWB_CL_FSPAN=0", "libwb_simt_double.so"),
            ("-DWB_CL_XWANTS=0 -DWB_CL_REFILTER=1 -DWB_CL_COMPACT2=0 -DWB_CL_TRANSPOSE=0", "libwb_simt_refilter.so"),
            ("-DWB_CL_XWANTS=1 -DWB_CL_REFILTER=0 -DWB_CL_COMPACT2=0 -DWB_CL_TRANSPOSE=0", "libwb_simt_xwants.so"),
            ("", "libwb_simt.so")]                                   # the shipped configuration


@pytest.fixture(scope="module")
def scenes():
    out = []
    for scene, n, seed, p in CASES:
        cloud = synth.generate(scene, n, seed=seed)
        res = O.run([O.file_from_cloud(cloud)], **p)
        out.append((res, O.point_hyperboloid_sizes(res, p.get("tile_size", 1.0)), p))
    return out


@pytest.mark.parametrize("variant,lib", VARIANTS)
def test_emulated_kernel_matches_oracle(scenes, variant, lib):
    for res, hyp, p in scenes:
        lab, work = emul.classify(res.points_sorted, hyp, p.get("max_slope", 1.0), p.get("thickness", 0.0),
                                  variant=variant, out=lib)
        mism = int((lab != res.labels_sorted).sum())
        assert mism <= work["margin"] + res.margin_count, (variant, mism)
        assert work["pairs"] > 0 and work["collectives"] > 0


def test_refilter_only_removes_rejected_pops(scenes):
    """The bulk re-filter drops children the per-pop test would reject anyway: fewer pops, the same chunks opened
    and the same (query, chunk) pairs tested."""
    res, hyp, p = scenes[0]
    _, base = emul.classify(res.points_sorted, hyp, variant=VARIANTS[0][0], out=VARIANTS[0][1])
    _, ref = emul.classify(res.points_sorted, hyp, variant=VARIANTS[2][0], out=VARIANTS[2][1])
    assert ref["nodes"] < base["nodes"]
    assert (ref["chunks"], ref["pairs"]) == (base["chunks"], base["pairs"])


def test_xwants_prunes_pops_not_work(scenes):
    """Per-query reach + open-sector test at expansion for every level: fewer nodes popped, the same chunks
    opened and pairs tested."""
    res, hyp, p = scenes[0]
    _, base = emul.classify(res.points_sorted, hyp, variant=VARIANTS[0][0], out=VARIANTS[0][1])
    _, x = emul.classify(res.points_sorted, hyp, variant=VARIANTS[3][0], out=VARIANTS[3][1])
    assert x["nodes"] < base["nodes"] and x["nodes2"] <= base["nodes2"]
    assert (x["chunks"], x["pairs"]) == (base["chunks"], base["pairs"])


@pytest.mark.parametrize("case", ["urban_40k_slope2", "street_30k_tile3"])
def test_emulated_kernel_matches_compiled_reference_labels(case, golden_dir):
    """Straight against the class bytes the UNMODIFIED reference produced (non-default maxSlope / tileSize)."""
    import json
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    p = json.loads(str(g["params"]))
    cloud = synth.generate(int(g["scene"]), int(g["n"]), seed=int(g["seed"]))
    res = O.run([O.file_from_cloud(cloud)], **p)
    assert res.n_duplicates == 0
    hyp = O.point_hyperboloid_sizes(res, p.get("tile_size", 1.0))
    lab, _ = emul.classify(res.points_sorted, hyp, p.get("max_slope", 1.0), p.get("thickness", 0.0))
    in_input_order = np.zeros(len(lab), dtype=np.uint8)
    in_input_order[res.order] = lab
    assert (in_input_order == g["ref_labels"]).all()


SCAN_CASES = [(2, 5000, 3, {}), (5, 5000, 5, {}), (1, 5000, 1, {"tile_size": 3.0}),
              (4, 5000, 4, {"tile_size": 2.0, "min_hyperboloid_size": 0.5, "max_slope": 0.7, "thickness": 0.05})]


@pytest.mark.parametrize("scene,n,seed,p", SCAN_CASES)
def test_emulated_tile_phases_and_classify_match_oracle(scene, n, seed, p):
    """Membership, tile scan (one warp per tile, shuffle-tree pairwise sums), postscan and classify, all from the
    kernel sources and launched as wb_scan / wb_postscan / wb_classify launch them: tile table bit-identical to the
    oracle's, labels identical."""
    cloud = synth.generate(scene, n, seed=seed)
    res = O.run([O.file_from_cloud(cloud)], **p)
    tiles, lab = emul.scan_classify(res.points_sorted, res.cube, p.get("tile_size", 1.0),
                                    p.get("min_hyperboloid_size", 0.1), p.get("max_slope", 1.0), p.get("thickness", 0.0))
    ot = res.tiles
    assert len(tiles) == len(ot)
    for f in ("n", "nPoints", "treeFlags"):
        assert (tiles[f] == ot[f]).all(), f
    for f in ("density", "hyperboloidSize", "height"):
        assert np.abs(tiles[f].view(np.int64) - ot[f].view(np.int64)).max() <= 4, f
    assert (lab == res.labels_sorted).all()


def test_emulated_tile_phases_match_compiled_reference(golden_dir):
    """The tile table the UNMODIFIED reference produced (format 6 fixture), from the kernel sources."""
    import json
    g = np.load(os.path.join(golden_dir, "multitile_fmt6_40k.npz"))
    cloud = synth.generate(int(g["scene"]), int(g["n"]), seed=int(g["seed"]))
    res = O.run([O.file_from_cloud(cloud)], classify=False)
    tiles, _ = emul.scan_classify(res.points_sorted, res.cube, classify=False)
    rt = g["ref_tiles"]
    assert len(tiles) == len(rt)
    for f in ("n", "nPoints", "treeFlags"):
        assert (tiles[f] == rt[f]).all(), f
    for f in ("density", "hyperboloidSize", "height"):
        assert np.abs(tiles[f].view(np.int64) - rt[f].view(np.int64)).max() <= 4, f

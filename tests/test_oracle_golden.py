"""The CPU oracle reproduces the COMPILED, UNMODIFIED reference (1 thread, canonical order) on
the committed golden fixtures: byte-identical octree dump, bit-identical tile table, identical
class bytes.  Fixtures are written by oracle/validate_against_ref.py --golden."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import wb_oracle as O
from wolkenbase_b200 import synth

CASES = sorted(os.path.basename(p)[:-4] for p in
               glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_fixture(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    p = json.loads(str(g["params"]))
    cloud = synth.generate(int(g["scene"]), int(g["n"]), seed=int(g["seed"]))
    dups = int(g["dups"]) if "dups" in g else 0
    if dups:
        cloud = synth.with_duplicates(cloud, dups, int(g["seed"]))
    res = O.run([O.file_from_cloud(cloud)], **p)
    assert list(res.root_center) + [res.root_side] == g["ref_root"].tolist()
    assert res.spacing == float(g["ref_spacing"]) and res.snake_index == int(g["ref_snake_index"])
    assert res.dump == bytes(g["ref_dump"]).decode("utf-8")
    rt = g["ref_tiles"]
    assert len(res.tiles) == len(rt)
    for f in ("n", "ex", "ey", "nPoints", "treeFlags"):
        assert (res.tiles[f] == rt[f]).all(), f
    for f in ("density", "hyperboloidSize", "height"):
        assert (res.tiles[f].view(np.uint64) == rt[f].view(np.uint64)).all(), f
    stored = g["ref_labels"] != 255          # records lost to an identical location are never stored
    assert int((~stored).sum()) == res.n_duplicates and (res.n_duplicates > 0) == (dups > 0)
    assert (res.labels[stored] == g["ref_labels"][stored]).all()


MULTI = sorted(os.path.basename(p)[:-4] for p in
               glob.glob(os.path.join(os.path.dirname(__file__), "golden", "multi", "*.npz")))


@pytest.mark.parametrize("case", MULTI)
def test_oracle_matches_reference_on_several_files(case, golden_dir):
    """Several input files at once — different header offsets, mixed LAS 1.2 format 1 / LAS 1.4 format 6 — as the
    compiled reference read them (fixtures: oracle/validate_against_ref.py, MULTI_CASES)."""
    from oracle.validate_against_ref import clouds_from_parts
    g = np.load(os.path.join(golden_dir, "multi", case + ".npz"))
    clouds = clouds_from_parts(json.loads(str(g["parts"])))
    res = O.run([O.file_from_cloud(c) for c in clouds], **json.loads(str(g["params"])))
    assert list(res.root_center) + [res.root_side] == g["ref_root"].tolist()
    assert res.spacing == float(g["ref_spacing"]) and res.snake_index == int(g["ref_snake_index"])
    assert res.dump == bytes(g["ref_dump"]).decode("utf-8")
    rt = g["ref_tiles"]
    assert len(res.tiles) == len(rt)
    for f in ("n", "ex", "ey", "nPoints", "treeFlags"):
        assert (res.tiles[f] == rt[f]).all(), f
    for f in ("density", "hyperboloidSize", "height"):
        assert (res.tiles[f].view(np.uint64) == rt[f].view(np.uint64)).all(), f
    assert len(g["ref_labels"]) == len(res.labels) and (res.labels == g["ref_labels"]).all()


def test_multi_fixtures_present():
    assert len(MULTI) >= 2


def test_sharded_fixture_is_what_the_oracle_says(golden_dir):
    """tests/golden/sharded/c3_8strips_400k.npz (what bench.py --gpus N checks its ranks against) regenerated."""
    import os
    import numpy as np
    from oracle import wb_oracle as O
    from wolkenbase_b200 import multigpu
    g = np.load(os.path.join(golden_dir, "sharded", "c3_8strips_400k.npz"))
    clouds = multigpu.parity_strips()
    assert [c.n for c in clouds] == g["counts"].tolist()
    res = O.run([O.file_from_cloud(c) for c in clouds], **multigpu.PARAMS)
    assert (res.labels == g["labels"]).all()
    assert float(res.tiles["hyperboloidSize"].max()) == float(g["hyp_max"])

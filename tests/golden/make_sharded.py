"""Golden labels for the N-GPU parity check that bench.py --gpus N runs before its timed region (and
tests/test_multigpu_nccl.py on a multi-GPU box): the C3 scene (multi-tile aerial, LAS format 6) at 400 k points cut
into 8 x-strips = 8 files; the oracle classifies the whole cloud, files in strip order.  With W = 2, 4 or 8 ranks,
rank r holds files [8r/W, 8(r+1)/W).
    python tests/golden/make_sharded.py      -> tests/golden/sharded/c3_8strips_400k.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wb_oracle as O  # noqa: E402
from wolkenbase_b200 import multigpu  # noqa: E402

if __name__ == "__main__":
    clouds = multigpu.parity_strips()
    res = O.run([O.file_from_cloud(c) for c in clouds], **multigpu.PARAMS)
    out = os.path.join(ROOT, "tests", "golden", "sharded", "c3_8strips_400k.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, labels=res.labels, counts=np.array([c.n for c in clouds], dtype=np.int64),
                        margin=int(res.margin_count), hyp_max=float(res.tiles["hyperboloidSize"].max()),
                        n_duplicates=int(res.n_duplicates))
    print(out, os.path.getsize(out), "bytes;", [c.n for c in clouds], np.bincount(res.labels, minlength=3).tolist(),
          "margin", res.margin_count)

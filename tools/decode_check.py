"""Decode kernel vs numpy for every supported record length and ragged counts."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
from wolkenbase_b200 import api
LEN = {0: 20, 1: 28, 2: 26, 3: 34, 6: 30, 7: 36, 8: 38}
rng = np.random.default_rng(0)
ctx = api.Context(0)
bad = 0
for rep in range(3):
    for fmt, ln in LEN.items():
        for n in [1, 2, 3, 5, 17, 255, 256, 257, 1000, 2977, 3999, 70001]:
            recs = rng.integers(0, 256, size=(n, ln), dtype=np.uint8)
            recs[:, 14] |= 1                      # non-zero return number everywhere
            ctx.clear()
            ctx.add_las(recs, fmt, (0.001,) * 3, (0.0,) * 3)
            x, y, z, c = ctx.decoded(n)
            ints = np.ascontiguousarray(recs[:, :12]).view(np.int32).reshape(n, 3)
            want_c = (recs[:, 15] & 31) if fmt < 6 else recs[:, 16]
            ok = (x == ints[:, 0]).all() and (y == ints[:, 1]).all() and (z == ints[:, 2]).all() and (c == want_c).all()
            if not ok:
                bad += 1
                w = np.nonzero((x != ints[:, 0]) | (y != ints[:, 1]) | (z != ints[:, 2]) | (c != want_c))[0]
                print("MISMATCH fmt", fmt, "len", ln, "n", n, "first bad", w[:5], "count", len(w))
print("bad cases:", bad)

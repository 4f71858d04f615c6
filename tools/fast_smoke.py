#!/usr/bin/env python
"""Small invocation of the lean traversal kernel and the seed kernel (for compute-sanitizer runs): 4000 x 128 float L2,
device seeds, edge cap 16, checked against the general kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngt_b200 import _lib, engine, synth  # noqa: E402

n, nq, k = 4000, 96, 10
base, qs = synth.make("sift", n, 1), synth.make("sift", nq, 2)
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, 128)
ix.set_objects(base)
gids, _, gcounts = ix.linear_search(base, 17)
row_ptr = np.zeros(n + 2, np.uint64)
row_ptr[2:] = np.cumsum(gcounts.astype(np.uint64))
col = np.concatenate([gids[i, :gcounts[i]] for i in range(n)]).astype(np.uint32)
ix.set_graph(row_ptr, col)
ix.build_seed_table(128, 1)
out = {}
for fast in (True, False):
    ix.set_fast_kernel(fast)
    out[fast] = ix.search(qs, k, 0.1, edge_size=16, n_seeds=10, with_stats=True)
for a, b in zip(out[True], out[False]):
    assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all()
print("fast smoke ok: mean distance computations/query %.1f" % out[True][3][:, 0].mean())

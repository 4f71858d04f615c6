#!/usr/bin/env python
"""Development probe: is ngtgpu_index_refine_anng reproducible run to run, and equal on the lean and the general
traversal kernel (result lists of 40 keys)? Prints one JSON line; on a mismatch, the first differing self-search."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_rows  # noqa: E402
from ngt_b200 import _lib, build, engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda", 0)
base = make_rows("sift", "f32", n, 1, dev)
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, base.shape[1])
ix.set_objects(base)
g = ix.build_onng(64, 10, 64, True, want_graph=True)
ix.build_seed_table(256, 1)
rp, col, dist = g["graph"]
out = {"n": n, "k": k}
runs = {}
for tag, fast in (("lean1", True), ("lean2", True), ("general1", False), ("general2", False)):
    ix.set_fast_kernel(fast)
    ix.set_graph(rp, col)
    runs[tag] = build.refine_anng(ix, rp, col, dist, 0.1, 0, -1, 10000, k, 10)
    out[tag + "_edges"] = int(runs[tag][1].numel())


def same(a, b):
    return bool(all(x.shape == y.shape and torch.equal(x, y) for x, y in zip(runs[a], runs[b])))


out["lean_reproducible"] = same("lean1", "lean2")
out["general_reproducible"] = same("general1", "general2")
out["lean_equals_general"] = same("lean1", "general1")
# the first batch's self-search, directly
ix.set_graph(rp, col)
q = base[:10000].contiguous()
res = {}
for es in (-1, 40, 64):
    for fast in (True, False):
        ix.set_fast_kernel(fast)
        res[fast] = [t.cpu().numpy() for t in ix.search(q, k, 0.1, edge_size=es, n_seeds=10, with_stats=True)]
    eq = [bool((x.view(np.uint32) == y.view(np.uint32)).all()) for x, y in zip(res[True], res[False])]
    out["self_search_es%d_equal" % es] = eq
    out["self_search_es%d_overflows" % es] = int(ix.last_overflows)
    if not all(eq):
        bad = np.nonzero((res[True][0] != res[False][0]).any(1) | (res[True][2] != res[False][2]))[0]
        out["self_search_es%d_bad" % es] = int(bad.size)
        if bad.size:
            b = int(bad[0])
            out["first_bad"] = {"q": b, "lean_ids": res[True][0][b].tolist(), "general_ids": res[False][0][b].tolist(),
                                "lean_d": res[True][1][b].tolist(), "general_d": res[False][1][b].tolist(),
                                "counts": [int(res[True][2][b]), int(res[False][2][b])],
                                "stats": [res[True][3][b].tolist(), res[False][3][b].tolist()]}
print(json.dumps(out), flush=True)

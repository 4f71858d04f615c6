#!/usr/bin/env python
"""Development probe: the same batched self-search repeated -- is the traversal reproducible run to run, on the lean and
on the general kernel, for k <= 32 and k > 32, on an ONNG and on a refined (high-degree) graph? One JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_rows  # noqa: E402
from ngt_b200 import _lib, build, engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
base = make_rows("sift", "f32", n, 1, dev)
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, base.shape[1])
ix.set_objects(base)
g = ix.build_onng(64, 10, 64, True, want_graph=True)
ix.build_seed_table(256, 1)
rp, col, dist = g["graph"]
ix.set_fast_kernel(False)
rp2, col2, dist2 = build.refine_anng(ix, rp, col, dist, 0.1, 0, -1, 10000, 40, 10)
out = {"n": n, "reps": reps, "cases": []}
for gname, (a, b) in (("onng", (rp, col)), ("refined", (rp2, col2))):
    ix.set_graph(a, b)
    for k in (10, 32, 40, 100):
        for fast in (True, False):
            ix.set_fast_kernel(fast)
            rec = {"graph": gname, "k": k, "lean": fast, "bad_queries": 0, "examples": []}
            first = None
            for r in range(reps):
                q = base[(r % 3) * 10000:(r % 3) * 10000 + 10000]
                res = [t.cpu().numpy() for t in ix.search(q, k, 0.1, edge_size=-1, n_seeds=10, with_stats=True)]
                if r < 3:
                    if first is None:
                        first = {}
                    first[r] = res
                    continue
                ref = first[r % 3]
                bad = np.nonzero((res[0] != ref[0]).any(1) | (res[1].view(np.uint32) != ref[1].view(np.uint32)).any(1) |
                                 (res[2] != ref[2]) | (res[3] != ref[3]).any(1))[0]
                rec["bad_queries"] += int(bad.size)
                for bq in bad[:2]:
                    bq = int(bq)
                    rec["examples"].append({"rep": r, "q": bq, "stats_now": res[3][bq].tolist(), "stats_first": ref[3][bq].tolist(),
                                            "count_now": int(res[2][bq]), "count_first": int(ref[2][bq]),
                                            "ids_differ": int((res[0][bq] != ref[0][bq]).sum())})
            rec["mean_n_dist"] = round(float(first[0][3][:, 0].mean()), 1)
            rec["max_n_dist"] = int(first[0][3][:, 0].max())
            out["cases"].append(rec)
print(json.dumps(out), flush=True)

#!/usr/bin/env python
"""Development / evidence probe: searches with result lists of 33..128 keys -- what the construction
loop and refineANNG ask for, k = edge count -- on the lean traversal kernel (search_fast_kernel<.., KL = 4>) against the
general kernel (sorted array in shared memory): ms per batch, identical answers, and refineANNG end to end on both.
One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_rows  # noqa: E402
from ngt_b200 import _lib, build, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--batch", type=int, default=10000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--kind", default="f32")
ap.add_argument("--refine-edges", type=int, default=40)
a = ap.parse_args()
dev = torch.device("cuda", 0)
base = make_rows("sift", a.kind, a.n, 1, dev)
qs = make_rows("sift", a.kind, a.batch, 2, dev)
otype = _lib.OBJECT_FLOAT if a.kind == "f32" else _lib.OBJECT_UINT8
ix = engine.GpuIndex(otype, _lib.DISTANCE_L2, base.shape[1])
ix.set_objects(base)
g = ix.build_onng(64, 10, 64, True, want_graph=True)
ix.build_seed_table(256, 1)
out = {"n": a.n, "kind": a.kind, "batch": a.batch, "searches": []}
for k, eps, cap in ((10, 0.1, 64), (40, 0.1, 64), (64, 0.1, 64), (100, 0.1, 100), (128, 0.1, 128)):
    rec = {"k": k, "epsilon": eps, "edge_size": cap}
    res = {}
    for fast in (True, False):
        ix.set_fast_kernel(fast)
        for _ in range(2):
            r = ix.search(qs, k, eps, edge_size=cap, n_seeds=10, with_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            ix.search(qs, k, eps, edge_size=cap, n_seeds=10)
        e1.record()
        torch.cuda.synchronize()
        rec["lean_ms" if fast else "general_ms"] = round(e0.elapsed_time(e1) / a.steps, 3)
        res[fast] = [t.cpu().numpy() for t in r]
    rec["identical"] = bool(all((x.view(np.uint32) == y.view(np.uint32)).all() for x, y in zip(res[True], res[False])))
    rec["n_dist"] = round(float(res[True][3][:, 0].mean()), 1)
    out["searches"].append(rec)
# refineANNG (batched self-search with k = edge count) on both kernels
rp, col, dist = g["graph"]
graphs = {}
for fast in (True, False):
    ix.set_fast_kernel(fast)
    ix.set_graph(rp, col)
    torch.cuda.synchronize()
    t = time.time()
    r = build.refine_anng(ix, rp, col, dist, 0.1, 0, -1, 10000, a.refine_edges, 10)
    torch.cuda.synchronize()
    out["refine_anng_lean_s" if fast else "refine_anng_general_s"] = round(time.time() - t, 3)
    graphs[fast] = r
out["refine_anng_identical"] = bool(all(torch.equal(x, y) for x, y in zip(graphs[True], graphs[False])))
out["refine_anng_edges"] = int(graphs[True][1].numel())
print(json.dumps(out), flush=True)

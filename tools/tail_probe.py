#!/usr/bin/env python
"""Development probe: how heavy is the tail of per-query work in the 10k batch, and how much of the kernel time is the
tail (longest queries finishing alone)? Runs the batch in natural order and sorted by descending work."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngt_b200 import _lib, build, engine, synth

dev = torch.device("cuda", 0)
n, nq, K = 1000000, 10000, 100
base = synth.make_device("sift", n, 1, dev)
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, 128)
ix.set_objects(base)
ix.set_search_workspace(14, 512)
ids, dists, counts = build.knn_graph(ix, K)
q = synth.make_device("sift", nq, 2, dev)
ix.build_seed_table(1024, 1)
rp, col, dd = build.reconstruct_graph(ids, dists, torch.clamp(counts, max=K), 10, 100)
ix.set_graph(rp, col)
ix.set_search_property(80, 30, 20)
eps = float(sys.argv[1]) if len(sys.argv) > 1 else 0.08

def timed(qq, reps=5):
    for _ in range(2):
        ix.search(qq, 10, eps, edge_size=80, n_seeds=10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ix.search(qq, 10, eps, edge_size=80, n_seeds=10)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

r = ix.search(q, 10, eps, edge_size=80, n_seeds=10, with_stats=True)
st = r[3].cpu().numpy().astype(np.int64)
for name, c in (("n_dist", 0), ("n_edge", 1), ("n_exp", 2)):
    v = st[:, c]
    print("%s: mean %.1f p50 %d p90 %d p99 %d p99.9 %d max %d" % (name, v.mean(), *np.percentile(v, [50, 90, 99, 99.9]).astype(int), v.max()))
print("natural order        : %.3f ms" % timed(q))
order = torch.from_numpy(np.argsort(-st[:, 0])).to(dev)
print("longest first        : %.3f ms" % timed(q[order].contiguous()))
order = torch.from_numpy(np.argsort(st[:, 0])).to(dev)
print("shortest first       : %.3f ms" % timed(q[order].contiguous()))
# the same total work with the tail cut: only queries below p99 work, repeated to 10k
keep = np.where(st[:, 0] <= np.percentile(st[:, 0], 90))[0]
sel = torch.from_numpy(np.resize(keep, nq)).to(dev)
w = st[np.resize(keep, nq), 0].sum() / st[:, 0].sum()
print("<= p90 queries only  : %.3f ms for %.3f of the work" % (timed(q[sel].contiguous()), w))

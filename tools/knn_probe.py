#!/usr/bin/env python
"""Development probe: one exhaustive kNN pass (ngtgpu_index_knn_graph) of a synthetic set, for ncu launch lists."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_rows
from ngt_b200 import _lib, build, engine
ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="f32")
ap.add_argument("--shape", default="sift")
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--k", type=int, default=128)
ap.add_argument("--queries", type=int, default=0, help="only the first batches covering this many queries (0: all)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
base = make_rows(a.shape, a.kind, a.n, 1, dev)
otype = _lib.OBJECT_FLOAT if a.kind == "f32" else _lib.OBJECT_UINT8
dist = _lib.DISTANCE_HAMMING if a.kind == "ham" else _lib.DISTANCE_NORMALIZED_COSINE if a.shape == "glove" else _lib.DISTANCE_L2
ix = engine.GpuIndex(otype, dist, base.shape[1])
ix.set_objects(base)
torch.cuda.synchronize()
lib = build._fn()
nq = a.queries or a.n
ids = torch.zeros((nq, a.k), dtype=torch.int32, device=dev)
dists = torch.zeros((nq, a.k), dtype=torch.float32, device=dev)
counts = torch.zeros((nq,), dtype=torch.int32, device=dev)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    for s in range(0, nq, 148 * 768):
        m = min(148 * 768, nq - s)
        _lib.check(lib.ngtgpu_index_knn_graph(ix._h, a.k, s + 1, m, ids[s:].data_ptr(), dists[s:].data_ptr(), counts[s:].data_ptr(), 0))
    torch.cuda.synchronize()
    t = time.time() - t0
    print("pass %d: %.3f s, %.1f useful TFLOP/s, tc batches %d" % (rep, t, 2.0 * nq * a.n * base.shape[1] * (8 if a.kind == "ham" else 1) / t / 1e12, ix.tensor_core_batches), flush=True)

#!/usr/bin/env python
"""Times the exhaustive kNN pass (tensor-core filter + exact re-evaluation vs the CUDA-core scan).
Development / evidence tool; prints one line per case."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngt_b200 import _lib, build, engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="sift")
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--k", type=int, default=64)
ap.add_argument("--queries", type=int, default=0, help="0: kNN graph (all rows as queries); else a batch of this size")
ap.add_argument("--cuda-core-sample", type=int, default=65536, help="rows as queries for the CUDA-core timing (0: skip)")
ap.add_argument("--distance", type=int, default=_lib.DISTANCE_L2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
base = synth.make_device(a.shape, a.n, 1, dev)
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, a.distance, base.shape[1])
ix.set_objects(base)
dim = ix.padded_dimension


def run(tc, nrows):
    ix.set_tensor_core(tc)
    torch.cuda.synchronize()
    t = time.time()
    if a.queries:
        q = synth.make_device(a.shape, a.queries, 2, dev)
        out = ix.linear_search(q, a.k)
        nq = a.queries
    else:
        lib = build._fn()
        ids = torch.zeros((nrows, a.k), dtype=torch.int32, device=dev)
        dists = torch.zeros((nrows, a.k), dtype=torch.float32, device=dev)
        counts = torch.zeros((nrows,), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        for s in range(0, nrows, 1 << 17):
            m = min(1 << 17, nrows - s)
            _lib.check(lib.ngtgpu_index_knn_graph(ix._h, a.k, s + 1, m, ids[s:].data_ptr(), dists[s:].data_ptr(),
                                                  counts[s:].data_ptr(), stream))
        out = (ids, dists, counts)
        nq = nrows
    torch.cuda.synchronize()
    return time.time() - t, nq, out


run(True, min(a.n, 1 << 17))          # warm-up: packs the row operand
t, nq, out_tc = run(True, a.n)
flops = 2.0 * nq * a.n * dim
print("%s n=%d k=%d queries=%d tensor-core path: %.3f s  (%.1f TFLOP/s of q.x over padded dim %d; batches on TC: %d)" % (
    a.shape, a.n, a.k, nq, t, flops / t / 1e12, dim, ix.tensor_core_batches), flush=True)
if a.cuda_core_sample:
    m = min(a.cuda_core_sample, a.n) if not a.queries else a.queries
    t2, nq2, out_cc = run(False, m)
    print("   CUDA-core scan on %d queries: %.3f s -> %.2f s for %d queries (x%.1f)" % (
        nq2, t2, t2 * nq / nq2, nq, t2 * nq / nq2 / t), flush=True)
    same = (out_tc[0][:nq2] == out_cc[0]).all().item() and (out_tc[1][:nq2].view(torch.int32) == out_cc[1].view(torch.int32)).all().item()
    print("   results identical on those queries:", bool(same), flush=True)

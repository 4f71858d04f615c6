#!/usr/bin/env python
"""Development / evidence probe: the reference's ANNG construction loop on the device
(ngtgpu_index_insert_batch, batches of 200 searched on the frozen graph) next to the route the bench uses (exact kNN
table on the tensor cores -> ANNG = its symmetric closure): build seconds and recall@10 over epsilon of both graphs,
searched by the same traversal kernel with the same seeds. One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_rows, recall_at_k  # noqa: E402
from ngt_b200 import _lib, build, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--edges", type=int, default=10)
ap.add_argument("--batch", type=int, default=200)
ap.add_argument("--eps-create", type=float, default=0.1)
ap.add_argument("--loop-only", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
base = make_rows("sift", "f32", a.n, 1, dev)
qs = make_rows("sift", "f32", 2000, 2, dev)
out = {"n": a.n, "edge_size_for_creation": a.edges, "batch_size_for_creation": a.batch}


def sweep(ix, tag):
    gt = ix.linear_search(qs, 10)
    ix.build_seed_table(256, 1)
    rows = []
    for eps in (0.0, 0.05, 0.1, 0.15, 0.2):
        torch.cuda.synchronize()
        t = time.time()
        r = ix.search(qs, 10, eps, edge_size=0, n_seeds=10, with_stats=True)
        torch.cuda.synchronize()
        dt = time.time() - t
        rec = recall_at_k(r[0].cpu().numpy().astype(np.uint32), r[1].cpu().numpy(), r[2].cpu().numpy().astype(np.int64),
                          gt[0].cpu().numpy().astype(np.uint32), gt[1].cpu().numpy())
        rows.append({"epsilon": eps, "recall_at_10": round(rec, 4), "n_dist": round(float(r[3][:, 0].float().mean()), 1)})
    out[tag + "_sweep"] = rows


# (1) the reference's loop
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, base.shape[1])
ix.set_objects(base)
torch.cuda.synchronize()
t = time.time()
rp, col, dist = build.insert_objects(ix, 1, a.n, None, a.edges, a.eps_create, -1, a.batch)
torch.cuda.synchronize()
out["insert_loop_s"] = round(time.time() - t, 2)
out["insert_loop_edges"] = int(col.numel())
import ctypes as C
cnt = (C.c_uint64 * 4)()
_lib.check(_lib.load().ngtgpu_construction_counters(cnt))
out["batches_merged"], out["batches_full_sort"], out["blocks_from_cudaMalloc"], out["blocks_from_cache"] = [int(v) for v in cnt]
out["insert_merge"] = os.environ.get("NGTGPU_INSERT_MERGE", "1")
if a.loop_only:
    print(json.dumps(out), flush=True)
    sys.exit(0)
ix.set_graph(rp, col)
sweep(ix, "insert_loop")
ix.close()
# (2) exact kNN table -> symmetric closure
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, base.shape[1])
ix.set_objects(base)
torch.cuda.synchronize()
t = time.time()
ids, dists, counts = build.knn_graph(ix, a.edges)
import ctypes as C
lib = _lib.load()
cap = 2 * a.n * a.edges + 16
rp2 = torch.zeros(a.n + 2, dtype=torch.int64, device=dev)
col2 = torch.zeros(cap, dtype=torch.int32, device=dev)
dist2 = torch.zeros(cap, dtype=torch.float32, device=dev)
nnz = C.c_uint64(0)
_lib.check(lib.ngtgpu_graph_from_knn_table(a.n, ids.data_ptr(), dists.data_ptr(), counts.data_ptr(), a.edges, None, 1, cap,
                                           rp2.data_ptr(), col2.data_ptr(), dist2.data_ptr(), C.byref(nnz),
                                           torch.cuda.current_stream(dev).cuda_stream))
torch.cuda.synchronize()
out["knn_table_s"] = round(time.time() - t, 2)
out["knn_table_edges"] = int(nnz.value)
ix.set_graph(rp2, col2[:nnz.value].clone())
sweep(ix, "knn_table")
print(json.dumps(out), flush=True)

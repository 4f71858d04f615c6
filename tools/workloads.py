#!/usr/bin/env python
"""BASELINE.json configs[2..4] on one B200 (development / evidence tool; bench.py stays the contract for configs[1]).

  python tools/workloads.py glove  [--n 1200000]    C3: 1.2M x 100 float, normalized cosine + angle, exhaustive kNN graph
                                                    build on the tensor cores (TFLOP/s), ANNG search
  python tools/workloads.py gist   [--n 250000]     C4: one shard of the 1M x 960 float L2 set (4-GPU shard by default)
  python tools/workloads.py u8     [--n 12500000]   C5: one shard of the 100M x 128 uint8 L2 set (8-GPU shard by default)
  python tools/workloads.py hamming [--n 12500000]  C5: the Hamming variant (128-bit objects)

Each prints one JSON line: graph construction times (kNN pass with its tensor-core rate), the epsilon that reaches
recall@10 >= 0.95 against the exhaustive scan, QPS of 10k-query batches (CUDA events, device-resident queries) and the
traversal kernel's algorithmic GB/s against the measured HBM peak."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import measured_peak_gbs, recall_at_k  # noqa: E402
from ngt_b200 import _lib, build, engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload", choices=["glove", "glove-angle", "gist", "u8", "hamming", "sift"])
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--knn", type=int, default=0)
ap.add_argument("--outgoing", type=int, default=-1)
ap.add_argument("--incoming", type=int, default=-1)
ap.add_argument("--edge-size", type=int, default=-1)
ap.add_argument("--adjust", type=int, default=1)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--gt", type=int, default=1000)
ap.add_argument("--recall", type=float, default=0.95)
ap.add_argument("--hash-bits", type=int, default=0)
ap.add_argument("--queue-cap", type=int, default=0)
ap.add_argument("--pivots", type=int, default=256)
a = ap.parse_args()

W = {   # shape, object type, distance, default n, knn, outgoing, incoming, edge cap
    "sift": ("sift", _lib.OBJECT_FLOAT, _lib.DISTANCE_L2, 1000000, 128, 10, 120, 80),
    "glove": ("glove", _lib.OBJECT_FLOAT, _lib.DISTANCE_NORMALIZED_COSINE, 1200000, 64, 10, 64, 64),
    "glove-angle": ("glove", _lib.OBJECT_FLOAT, _lib.DISTANCE_ANGLE, 1200000, 64, 10, 64, 64),
    "gist": ("gist", _lib.OBJECT_FLOAT, _lib.DISTANCE_L2, 250000, 64, 10, 64, 64),
    "u8": ("sift", _lib.OBJECT_UINT8, _lib.DISTANCE_L2, 12500000, 64, 10, 64, 64),
    "hamming": ("sift", _lib.OBJECT_UINT8, _lib.DISTANCE_HAMMING, 12500000, 64, 10, 64, 64),
}
shape, otype, dtype, n0, knn0, o0, i0, cap0 = W[a.workload]
n = a.n or n0
knn = a.knn or knn0
outgoing = a.outgoing if a.outgoing >= 0 else o0
incoming = a.incoming if a.incoming >= 0 else i0
cap = a.edge_size if a.edge_size >= 0 else cap0
dev = torch.device("cuda", 0)


def make(count, seed):
    """rows of the workload on the device, generated in pieces (the uint8 / bit sets are large)."""
    out = []
    step = 2000000
    for s in range(0, count, step):
        m = min(step, count - s)
        x = synth.make_device(shape, m, seed * 1000003 + s, dev) if count > step else synth.make_device(shape, m, seed, dev)
        if a.workload == "u8":
            x = x.to(torch.uint8)
        elif a.workload == "hamming":
            bits = (x > 64.0).to(torch.uint8).reshape(m, -1, 8)
            wts = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=dev)
            x = (bits * wts).sum(-1).to(torch.uint8)
        out.append(x)
    return torch.cat(out) if len(out) > 1 else out[0]


t0 = time.time()
base = make(n, 1)
dim = base.shape[1]
ix = engine.GpuIndex(otype, dtype, dim)
ix.set_objects(base)
elem = 1 if otype == _lib.OBJECT_UINT8 else 4
row_bytes = dim * elem
del base
torch.cuda.synchronize()
t1 = time.time()
tc0 = ix.tensor_core_batches
ids, dists, counts = build.knn_graph(ix, knn)
torch.cuda.synchronize()
t2 = time.time()
tc_batches = ix.tensor_core_batches - tc0
n_batches = (n + (1 << 17) - 1) >> 17
kdim = ix.padded_dimension * (8 if a.workload == "hamming" else 1)
row_ptr, col, dist = build.reconstruct_graph(ids, dists, counts, outgoing, incoming)
del ids, dists, counts
torch.cuda.synchronize()
t3 = time.time()
adj = None
if a.adjust:
    row_ptr, col, dist, adj = build.adjust_paths(row_ptr, col, dist, 0, with_stats=True)
torch.cuda.synchronize()
t4 = time.time()
gstats = build.graph_statistics(row_ptr)
ix.set_graph(row_ptr, col)
ix.set_search_property(cap, 30, 20)
# long rows (960-d) need many more distance evaluations per query: a larger visited slab / queue per CTA
hb = a.hash_bits or (16 if a.workload == "gist" else 15 if n > 4000000 else 14)
qc = a.queue_cap or (2048 if a.workload == "gist" else 512)
ix.set_search_workspace(hb, qc)
ix.build_seed_table(a.pivots, 1)
del row_ptr, col, dist
torch.cuda.empty_cache()

nb = 3
q_all = make(a.nq * nb, 2)
batches = [q_all[i * a.nq:(i + 1) * a.nq].contiguous() for i in range(nb)]
gq = batches[0][:a.gt]
gt_ids, gt_d, _ = ix.linear_search(gq, 10)
gt_ids, gt_d = gt_ids.cpu().numpy().astype(np.uint32), gt_d.cpu().numpy()
curve, eps, rec = [], None, 0.0
for step in range(0, 21):
    e = round(0.02 * step, 2)
    r = ix.search(gq, 10, e, edge_size=cap, n_seeds=10, with_stats=True)
    rc = recall_at_k(r[0].cpu().numpy().astype(np.uint32), r[1].cpu().numpy(), r[2].cpu().numpy().astype(np.int64), gt_ids, gt_d)
    curve.append({"epsilon": e, "recall": round(rc, 4), "n_dist": round(float(r[3][:, 0].float().mean()), 1)})
    if rc >= a.recall:
        eps, rec = e, rc
        break
if eps is None:
    eps, rec = curve[-1]["epsilon"], curve[-1]["recall"]
r = ix.search(batches[0], 10, eps, edge_size=cap, n_seeds=10, with_stats=True)
st = r[3].cpu().numpy().astype(np.int64)
bytes_step = int((st[:, 0] * row_bytes + st[:, 1] * 4).sum())
ix.search(batches[0].cpu().numpy(), 10, eps, edge_size=cap, n_seeds=10)    # host-pointer call: reports the overflow count
overflow = ix.last_overflows
lib = _lib.load()
lib.ngtgpu_index_set_timing.argtypes = [C.c_void_p, C.c_int]
lib.ngtgpu_index_pop_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
for w in range(3):
    ix.search(batches[w % nb], 10, eps, edge_size=cap, n_seeds=10)
torch.cuda.synchronize()
lib.ngtgpu_index_set_timing(ix._h, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(a.steps):
    ix.search(batches[s % nb], 10, eps, edge_size=cap, n_seeds=10)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
kms, kcnt = C.c_double(0), C.c_uint64(0)
lib.ngtgpu_index_pop_timing(ix._h, C.byref(kms), C.byref(kcnt))
k_ms = kms.value / max(kcnt.value, 1)
peak, peak_src = measured_peak_gbs()
# exhaustive scan of one batch (linearSearch, the brute-force row of SURVEY 8d)
torch.cuda.synchronize()
tl = time.time()
ix.linear_search(batches[1], 10)
torch.cuda.synchronize()
lin_s = time.time() - tl
knn_s = t2 - t1
out = {
    "workload": a.workload, "n": n, "dim": dim, "object": "uint8" if elem == 1 else "float32", "distance_type": int(dtype),
    "graph_build": {"knn": knn, "knn_pass_s": round(knn_s, 2), "tensor_core_batches": "%d of %d" % (tc_batches, n_batches),
                    "knn_useful_tflops": round(2.0 * n * n * kdim / knn_s / 1e12, 1),
                    "reconstruct_s": round(t3 - t2, 2), "outgoing": outgoing, "incoming": incoming,
                    "adjust_paths_s": round(t4 - t3, 2) if a.adjust else None, "adjust": adj, "graph": gstats,
                    "set_objects_s": round(t1 - t0, 2)},
    "search": {"batch": a.nq, "k": 10, "edge_size": cap, "epsilon": eps, "recall_at_10": round(rec, 4), "gt_queries": a.gt,
               "qps": round(a.nq / ms * 1e3, 1), "ms_per_batch": round(ms, 3), "kernel_ms": round(k_ms, 3),
               "n_dist_per_query": round(float(st[:, 0].mean()), 1), "n_edge_per_query": round(float(st[:, 1].mean()), 1),
               "algorithmic_gbs": round(bytes_step / (k_ms / 1e3) / 1e9, 1) if k_ms > 0 else None,
               "frac_of_measured_hbm_peak": round(bytes_step / (k_ms / 1e3) / 1e9 / peak, 4) if k_ms > 0 else None,
               "peak_gbs": peak, "overflow_queries": overflow, "hash_bits": hb, "queue_cap": qc, "epsilon_sweep": curve},
    "linear_search": {"batch": a.nq, "seconds": round(lin_s, 3), "scan_gbs_per_query_pass": round(n * row_bytes / lin_s / 1e9, 1),
                      "qps": round(a.nq / lin_s, 1)},
}
print(json.dumps(out), flush=True)

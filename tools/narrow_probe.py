#!/usr/bin/env python
"""Development probe: traversal-kernel time of the narrow-row configurations (uint8 L2 / Hamming, 128-byte and 16-byte rows)
under different shapes of the lean kernel (warps per query, resident CTAs per SM, slab size). One JSON line per setting."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_rows, measured_peak_gbs, recall_at_k  # noqa: E402
from ngt_b200 import _lib, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="u8")
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--eps", type=float, default=0.1)
ap.add_argument("--knn", type=int, default=64)
ap.add_argument("--shape", default="")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--batch", type=int, default=10000)
ap.add_argument("--edge", type=int, default=64)
ap.add_argument("--settings", default="2:0:0,1:32:0,1:24:0,1:16:0")   # warps:ctas:hash_bits (0 = default)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.ngtgpu_index_set_timing.argtypes = [C.c_void_p, C.c_int]
lib.ngtgpu_index_pop_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
shape = "glove" if a.kind == "glove" else "sift"
kind = "f32" if a.kind == "glove" else a.kind
base = make_rows(shape, kind, a.n, 1, dev)
dim = base.shape[1]
otype = _lib.OBJECT_FLOAT if kind == "f32" else _lib.OBJECT_UINT8
dist = _lib.DISTANCE_NORMALIZED_COSINE if a.kind == "glove" else _lib.DISTANCE_HAMMING if kind == "ham" else _lib.DISTANCE_L2
ix = engine.GpuIndex(otype, dist, dim)
ix.set_objects(base)
ix.build_onng(64, 10, 64, True)
ix.set_search_property(a.edge, 30, 20)
ix.build_seed_table(256, 1)
qs = make_rows(shape, kind, 3 * a.batch, 2, dev)
batches = [qs[i * a.batch:(i + 1) * a.batch].contiguous() for i in range(3)]
gt = ix.linear_search(batches[0][:1000], 10)
peak, _ = measured_peak_gbs()
elem = 4 if kind == "f32" else 1
ref = None
for s in a.settings.split(","):
    w, ctas, hb = [int(v) for v in s.split(":")]
    _lib.check(lib.ngtgpu_index_set_fast_shape(ix._h, w, ctas))
    if hb:
        ix.set_search_workspace(hb, 512)
    r = ix.search(batches[0], 10, a.eps, edge_size=a.edge, n_seeds=10, with_stats=True)
    st = r[3].cpu().numpy().astype(np.int64)
    ids = r[0].cpu().numpy()
    if ref is None:
        ref = ids
    same = bool((ids == ref).all())
    rec = recall_at_k(ids[:1000].astype(np.uint32), r[1].cpu().numpy()[:1000], r[2].cpu().numpy()[:1000].astype(np.int64),
                      gt[0].cpu().numpy().astype(np.uint32), gt[1].cpu().numpy())
    ix.search(batches[0].cpu().numpy(), 10, a.eps, edge_size=a.edge, n_seeds=10)
    ovf = ix.last_overflows
    bytes_step = int((st[:, 0] * dim * elem + st[:, 1] * 4).sum())
    for i in range(3):
        ix.search(batches[i % 3], 10, a.eps, edge_size=a.edge, n_seeds=10)
    torch.cuda.synchronize()
    lib.ngtgpu_index_set_timing(ix._h, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        ix.search(batches[i % 3], 10, a.eps, edge_size=a.edge, n_seeds=10)
    e1.record()
    torch.cuda.synchronize()
    kms, kc = C.c_double(0), C.c_uint64(0)
    lib.ngtgpu_index_pop_timing(ix._h, C.byref(kms), C.byref(kc))
    lib.ngtgpu_index_set_timing(ix._h, 0)
    k_ms = kms.value / max(kc.value, 1)
    print(json.dumps({"kind": a.kind, "n": a.n, "batch": a.batch, "warps": w, "ctas_cap": ctas, "hash_bits": hb, "eps": a.eps, "recall": round(rec, 4),
                      "step_ms": round(e0.elapsed_time(e1) / a.steps, 3), "kernel_ms": round(k_ms, 3), "overflow": ovf,
                      "gbs": round(bytes_step / k_ms / 1e6, 1), "frac": round(bytes_step / k_ms / 1e6 / peak, 4),
                      "n_dist": round(float(st[:, 0].mean()), 1), "n_exp": round(float(st[:, 2].mean()), 1), "same_ids_as_first": same}), flush=True)

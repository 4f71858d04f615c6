#!/usr/bin/env python
"""Development probe: how much of a 10k-query traversal launch is its tail (the last queries finishing alone), and does
handing the queries out heaviest-first shorten it? Orders: natural, by the actual work (oracle), by predictors that are
known before the traversal starts (distance to the nearest / the tenth seed). One JSON line per order."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_rows, measured_peak_gbs  # noqa: E402
from ngt_b200 import _lib, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="f32")
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--eps", type=float, default=0.08)
ap.add_argument("--edge", type=int, default=80)
ap.add_argument("--steps", type=int, default=6)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.ngtgpu_index_set_timing.argtypes = [C.c_void_p, C.c_int]
lib.ngtgpu_index_pop_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
kind = a.kind
base = make_rows("sift", kind, a.n, 1, dev)
dim = base.shape[1]
otype = _lib.OBJECT_FLOAT if kind == "f32" else _lib.OBJECT_UINT8
ix = engine.GpuIndex(otype, _lib.DISTANCE_L2, dim)
ix.set_objects(base)
ix.build_onng(64, 10, 64, True)
ix.set_search_property(a.edge, 30, 20)
ix.build_seed_table(256, 1)
qs = make_rows("sift", kind, 10000, 2, dev)
r = ix.search(qs, 10, a.eps, edge_size=a.edge, n_seeds=10, with_stats=True)
work = r[3].cpu().numpy().astype(np.int64)[:, 0]
# predictors: exact distances to the selected seeds
qh = qs.cpu().numpy()
seeds = np.zeros((10000, 10), np.uint32)
lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
lib.ngtgpu_select_seeds.restype = C.c_int
_lib.check(lib.ngtgpu_select_seeds(ix._h, qh.ctypes.data, 2 if kind == "f32" else 1, 10000, 10, seeds.ctypes.data))
bq = qs.float()
sd = torch.stack([((base[torch.from_numpy(seeds[:, j].astype(np.int64) - 1).to(dev)].float() - bq) ** 2).sum(1) for j in (0, 9)], 1).cpu().numpy()
print(json.dumps({"corr_work_d_seed0": float(np.corrcoef(work, np.sqrt(sd[:, 0]))[0, 1]),
                  "corr_work_d_seed9": float(np.corrcoef(work, np.sqrt(sd[:, 1]))[0, 1]),
                  "work_mean": float(work.mean()), "work_p50": float(np.percentile(work, 50)), "work_p99": float(np.percentile(work, 99)),
                  "work_max": int(work.max())}), flush=True)
orders = {"natural": np.arange(10000), "oracle_desc": np.argsort(-work, kind="stable"),
          "seed0_desc": np.argsort(-sd[:, 0], kind="stable"), "seed9_desc": np.argsort(-sd[:, 1], kind="stable"),
          "seed9_asc": np.argsort(sd[:, 1], kind="stable")}
for name, order in orders.items():
    q = qs[torch.from_numpy(order).to(dev)].contiguous()
    for _ in range(3):
        ix.search(q, 10, a.eps, edge_size=a.edge, n_seeds=10)
    torch.cuda.synchronize()
    lib.ngtgpu_index_set_timing(ix._h, 1)
    for _ in range(a.steps):
        ix.search(q, 10, a.eps, edge_size=a.edge, n_seeds=10)
    torch.cuda.synchronize()
    kms, kc = C.c_double(0), C.c_uint64(0)
    lib.ngtgpu_index_pop_timing(ix._h, C.byref(kms), C.byref(kc))
    lib.ngtgpu_index_set_timing(ix._h, 0)
    print(json.dumps({"kind": kind, "order": name, "kernel_ms": round(kms.value / max(kc.value, 1), 3)}), flush=True)

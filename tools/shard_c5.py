#!/usr/bin/env python
"""BASELINE.json configs[4]: the synthetic 128-d uint8 L2 set sharded by rows over the GPUs of one box
(SURVEY.md section 8e), 12.5M objects per GPU by default -- 100M over 8 B200.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/shard_c5.py [--n-per-gpu 12500000]

Every rank builds its own shard on its GPU (exact 64-NN graph on the tensor cores, the reference's ONNG recipe
-o 10 -i 64 + shortcut reduction, seed table), all ranks answer the same 10k-query batches, the per-shard top-10 lists
are exchanged with ONE NCCL all-gather and merged on the device (ngt_b200/sharded.py). Ground truth is the exhaustive
scan of every shard merged the same way; epsilon is the smallest of 0.00, 0.02, ... with recall@10 >= 0.95 on the
union. Timing: CUDA events around K batches between barriers, max over ranks. Rank 0 prints one JSON line.
Development / evidence tool; bench.py stays the contract for configs[1]."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import measured_peak_gbs, recall_at_k  # noqa: E402
from ngt_b200 import _lib, build, engine, sharded, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-per-gpu", type=int, default=12500000)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--gt", type=int, default=1000)
ap.add_argument("--recall", type=float, default=0.95)
ap.add_argument("--knn", type=int, default=64)
ap.add_argument("--hamming", action="store_true", help="the 128-bit Hamming variant of the same rows")
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def make(count, seed):
    out, step = [], 2000000
    for s in range(0, count, step):
        m = min(step, count - s)
        x = synth.make_device("sift", m, seed * 1000003 + s, dev)
        if a.hamming:
            bits = (x > 64.0).to(torch.uint8).reshape(m, -1, 8)
            wts = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=dev)
            x = (bits * wts).sum(-1).to(torch.uint8)
        else:
            x = x.to(torch.uint8)
        out.append(x)
    return torch.cat(out) if len(out) > 1 else out[0]


n_local = a.n_per_gpu
t0 = time.time()
base = make(n_local, 1000 + rank)
dim = base.shape[1]
ix = engine.GpuIndex(_lib.OBJECT_UINT8, _lib.DISTANCE_HAMMING if a.hamming else _lib.DISTANCE_L2, dim, device=local)
ix.set_objects(base)
del base
t1 = time.time()
ids, dists, counts = build.knn_graph(ix, a.knn)
torch.cuda.synchronize()
t2 = time.time()
row_ptr, col, dd = build.reconstruct_graph(ids, dists, counts, 10, 64)
del ids, dists, counts
row_ptr, col, dd = build.adjust_paths(row_ptr, col, dd, 0)
torch.cuda.synchronize()
t3 = time.time()
gstats = build.graph_statistics(row_ptr)
ix.set_graph(row_ptr, col)
ix.set_search_property(64, 30, 20)
ix.build_seed_table(256, 1)
del row_ptr, col, dd
torch.cuda.empty_cache()

S = sharded.ShardedSearcher(ix, rank, world, n_local)
nb = 3
q_all = make(a.nq * nb, 2)      # the same queries on every rank
batches = [q_all[i * a.nq:(i + 1) * a.nq].contiguous() for i in range(nb)]
gq = batches[0][:a.gt]


def merged(fn, *args, **kw):
    if world > 1:
        return fn(*args, **kw)
    return (ix.linear_search if fn == S.linear_search else ix.search)(*args, **kw)[:3]


gt_ids, gt_d, _ = merged(S.linear_search, gq, 10)
gt_ids, gt_d = gt_ids.cpu().numpy().astype(np.uint32), gt_d.cpu().numpy()
curve, eps, rec = [], None, 0.0
for step in range(0, 21):
    e = round(0.02 * step, 2)
    r = merged(S.search, gq, 10, e, edge_size=64, n_seeds=10)
    rc = recall_at_k(r[0].cpu().numpy().astype(np.uint32), r[1].cpu().numpy(), r[2].cpu().numpy().astype(np.int64), gt_ids, gt_d)
    curve.append({"epsilon": e, "recall": round(rc, 4)})
    if rc >= a.recall:
        eps, rec = e, rc
        break
if eps is None:
    eps, rec = curve[-1]["epsilon"], curve[-1]["recall"]

# this shard's work per query and its traversal-kernel time
r = ix.search(batches[0], 10, eps, edge_size=64, n_seeds=10, with_stats=True)
st = r[3].cpu().numpy().astype(np.int64)
bytes_step = int((st[:, 0] * dim + st[:, 1] * 4).sum())
lib = _lib.load()
lib.ngtgpu_index_set_timing.argtypes = [C.c_void_p, C.c_int]
lib.ngtgpu_index_pop_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
for w in range(3):
    merged(S.search, batches[w % nb], 10, eps, edge_size=64, n_seeds=10)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
lib.ngtgpu_index_set_timing(ix._h, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(a.steps):
    merged(S.search, batches[s % nb], 10, eps, edge_size=64, n_seeds=10)
e1.record()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
kms, kcnt = C.c_double(0), C.c_uint64(0)
lib.ngtgpu_index_pop_timing(ix._h, C.byref(kms), C.byref(kcnt))
k_ms = torch.tensor([kms.value / max(kcnt.value, 1)], device=dev)
gbs = torch.tensor([bytes_step / (k_ms.item() / 1e3) / 1e9], device=dev)
setup = torch.tensor([t1 - t0, t2 - t1, t3 - t2], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(gbs, op=dist.ReduceOp.MIN)
    dist.all_reduce(setup, op=dist.ReduceOp.MAX)
peak, _ = measured_peak_gbs()
if rank == 0:
    print(json.dumps({
        "workload": "configs[4]: synthetic %dx%d uint8 %s, %d x B200, %d objects per GPU, rows sharded + one NCCL all-gather + device merge"
                    % (n_local * world, dim * (8 if a.hamming else 1) if a.hamming else dim, "Hamming" if a.hamming else "L2", world, n_local),
        "n_gpus": world, "objects": n_local * world, "batch": a.nq, "k": 10, "epsilon": eps, "recall_at_10": round(rec, 4),
        "gt": "exhaustive scan of every shard, merged", "gt_queries": a.gt,
        "qps": round(a.nq / ms.item() * 1e3, 1), "ms_per_batch_max_over_ranks": round(ms.item(), 3),
        "traversal_kernel_ms_max_over_ranks": round(k_ms.item(), 3),
        "traversal_algorithmic_gbs_min_over_ranks": round(gbs.item(), 1), "frac_of_measured_hbm_peak": round(gbs.item() / peak, 4),
        "n_dist_per_query_rank0": round(float(st[:, 0].mean()), 1),
        "setup_s_max_over_ranks": {"objects": round(setup[0].item(), 1), "knn_pass": round(setup[1].item(), 1),
                                   "onng": round(setup[2].item(), 1)},
        "graph_rank0": gstats, "epsilon_sweep": curve}), flush=True)
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python
"""Graph / search parameter sweep on the GPU (development tool, not part of the bench contract).
Builds the exact kNN table once, then for each (outgoing, incoming, knn, edge cap) finds the smallest epsilon
with recall@10 >= target and times the 10k batch."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import recall_at_k  # noqa: E402
from ngt_b200 import _lib, build, engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--kmax", type=int, default=96)
ap.add_argument("--configs", default="64,10,64,64")   # knn,outgoing,incoming,cap;...
ap.add_argument("--eps", default="0.04,0.06,0.08,0.10,0.12")
ap.add_argument("--hash-bits", type=int, default=14)
ap.add_argument("--queue-cap", type=int, default=512)
ap.add_argument("--pivots", type=int, default=1024)
ap.add_argument("--seeds", type=int, default=10)
ap.add_argument("--target", type=float, default=0.95)
ap.add_argument("--prof", action="store_true")   # needs a library built with EXTRA=-DSEARCH_PHASE_PROFILE
ap.add_argument("--stage-bytes", default="16384")
ap.add_argument("--adjust", type=int, default=-1)   # >= 0: adjustPathsEffectively with this minNoOfEdges
a = ap.parse_args()

dev = torch.device("cuda", 0)
base = synth.make_device("sift", a.n, 1, dev)
ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, 128)
ix.set_objects(base)
ix.set_search_workspace(a.hash_bits, a.queue_cap)
t = time.time()
ids, dists, counts = build.knn_graph(ix, a.kmax)
torch.cuda.synchronize()
print("knn graph k=%d: %.1fs" % (a.kmax, time.time() - t), flush=True)
q = synth.make_device("sift", a.nq, 2, dev)
ngt = 2000
gt_ids, gt_d, _ = ix.linear_search(q[:ngt], 10)
gt_ids, gt_d = gt_ids.cpu().numpy().astype(np.uint32), gt_d.cpu().numpy()
ix.build_seed_table(a.pivots, 1)
for cfg in a.configs.split(";"):
    K, o, i, cap = [int(v) for v in cfg.split(",")]
    rp, col, dd = build.reconstruct_graph(ids[:, :K].contiguous(), dists[:, :K].contiguous(),
                                          torch.clamp(counts, max=K), o, i)
    if a.adjust >= 0:
        torch.cuda.synchronize()
        t = time.time()
        rp, col, dd, ast = build.adjust_paths(rp, col, dd, a.adjust, with_stats=True)
        torch.cuda.synchronize()
        print("adjust_paths: %.2fs %s" % (time.time() - t, ast), flush=True)
    st = build.graph_statistics(rp)
    ix.set_graph(rp, col)
    ix.set_search_property(cap, 30, 20)
    for eps, sb in [(float(e), int(b)) for b in a.stage_bytes.split(",") for e in a.eps.split(",")]:
        ix.set_search_workspace(a.hash_bits, a.queue_cap, stage_bytes=sb)
        print("stage_bytes=%d queue=%d hash_bits=%d" % (sb, a.queue_cap, a.hash_bits))
        r = ix.search(q[:ngt], 10, eps, edge_size=cap, n_seeds=a.seeds, with_stats=True)
        rec = recall_at_k(r[0].cpu().numpy().astype(np.uint32), r[1].cpu().numpy(), r[2].cpu().numpy().astype(np.int64),
                          gt_ids, gt_d)
        s = r[3].float().mean(0).cpu().numpy()
        for _ in range(2):
            ix.search(q, 10, eps, edge_size=cap, n_seeds=a.seeds)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ix.search(q, 10, eps, edge_size=cap, n_seeds=a.seeds)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        gbs = (s[0] * 512 + s[1] * 4) * a.nq / (ms / 1e3) / 1e9
        if a.prof:
            import ctypes as C
            lib = _lib.load()
            lib.ngtgpu_index_set_phase_profile.argtypes = [C.c_void_p, C.c_void_p]
            pb = torch.zeros((a.nq, 8), dtype=torch.int32, device=dev)
            lib.ngtgpu_index_set_phase_profile(ix._h, pb.data_ptr())
            ix.search(q, 10, eps, edge_size=cap, n_seeds=a.seeds)
            torch.cuda.synchronize()
            lib.ngtgpu_index_set_phase_profile(ix._h, None)
            pm = pb.float().mean(0).cpu().numpy()
            names = ["merge", "pop", "next-was-2nd-or-3rd", "filter", "next-was-2nd", "row-wait", "dist+sync", "other"]
            print("   cycles/query by phase: " + "  ".join("%s=%.0f" % (n_, v) for n_, v in zip(names, pm)) +
                  "  total=%.0f  per-expansion=%.0f" % (pm.sum(), pm.sum() / max(s[2], 1)), flush=True)
        print("K=%d o=%d i=%d cap=%d deg=%.1f eps=%.2f recall=%.4f ndist=%.0f nedge=%.0f nexp=%.1f  %.2f ms  %.0f QPS  %.0f GB/s(step)" % (
            K, o, i, cap, st["mean_degree"], eps, rec, s[0], s[1], s[2], ms, a.nq / ms * 1e3, gbs), flush=True)
        if rec >= a.target + 0.02:
            break

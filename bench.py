#!/usr/bin/env python
"""bench.py -- QPS of batched graph search at recall@10 >= 0.95 on synthetic 1M x 128 float L2 (sift-shape),
batch 10k, k = 10 (BASELINE.json `metric`, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this engine (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU search on the host cores

One "step" = one batch of 10k queries through the hot path (seed selection + graph traversal, k = 10) at the
smallest epsilon of the sweep 0.00..0.30 (step 0.02) whose recall@10 (Optimizer.h:496-507 definition, ground
truth from the exhaustive scan) is >= 0.95.

  value     QPS with the query batch already in HBM (CUDA events on the launching stream, max over ranks)
  e2e       QPS through the host-pointer C-ABI call ngtgpu_search(): pinned host queries in, host results out,
            both copies inside the timed region
  roofline  traversal kernel only: algorithmic bytes (n_dist * D * 4 + n_edge * 4 per query, from the kernel's
            own counters) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference (oracle/_ref, NGT::Index::search on the SAME index files, same seeds,
            same epsilon) on all host threads, on a bounded sample of the batch

Setup (untimed, identical for both arms): synthetic data (ngt_b200.synth, seeds 1/2), exact kNN graph on the
device, the reference's ONNG recipe on it (reconstructGraph -o 10 -i 120 + shortcut reduction,
GraphReconstructor.h:425-561 and 197-386, both on the device), seed table. N > 1 runs one replica of the
index per GPU with its own 10k batch (weak scaling, no data-path collective); `--mode shard` instead shards
the rows over the ranks and merges per-shard top-k lists after an NCCL all-gather.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="replica", choices=["replica", "shard"])
    ap.add_argument("--n", "--objects", dest="n", type=int, default=1000000)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--shape", default="sift")
    ap.add_argument("--knn", type=int, default=128, help="edges per node of the exact kNN graph")
    ap.add_argument("--outgoing", type=int, default=10)
    ap.add_argument("--incoming", type=int, default=120)
    ap.add_argument("--shortcut-reduction", type=int, default=1,
                    help="1: GraphReconstructor::adjustPathsEffectively after reconstructGraph (the reference's ONNG recipe)")
    ap.add_argument("--edge-size", type=int, default=80, help="edge_size of the search (0 = all edges)")
    ap.add_argument("--pivots", type=int, default=256)
    ap.add_argument("--seeds", type=int, default=10)
    ap.add_argument("--recall", type=float, default=0.95)
    ap.add_argument("--gt-queries", type=int, default=2000)
    ap.add_argument("--cpu-sample", type=int, default=2000)
    ap.add_argument("--epsilon", type=float, default=None, help="skip the sweep and use this epsilon")
    ap.add_argument("--hash-bits", type=int, default=14)
    ap.add_argument("--queue-cap", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--index-dir", default=None)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def recall_at_k(ids, dists, counts, gt_ids, gt_d):
    """lib/NGT/Optimizer.h:496-507: a result is relevant if its id is in the ground truth or its distance is
    <= the farthest ground-truth distance; accuracy = relevant / |ground truth|."""
    k = gt_ids.shape[1]
    valid = np.arange(ids.shape[1])[None, :] < counts[:, None]
    hit = (ids[:, :, None] == gt_ids[:, None, :]).any(-1)
    far = gt_d[:, -1:]
    rel = valid & (hit | ((far > 0) & (dists <= far)))
    return float(rel.sum(1).mean() / k)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = [r for (t, r) in self.rows if t0 is None or (t0 - 0.05 <= t <= t1 + 0.15)]
        if not rows and self.rows:      # region shorter than the sampling period: take the nearest sample
            mid = 0.5 * (t0 + t1)
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - mid))[1]]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def metric_name(a, dim):
    """BASELINE.json's metric on configs[1]; other shapes (development runs of configs[2..4]) say what they are."""
    if a.shape == "sift" and a.n == 1000000:
        return "QPS at recall@10>=0.95, synthetic 1Mx128 float L2 (sift-shape), batch 10k, k=10"
    return "QPS at recall@10>=%.2f, synthetic %dx%d float L2 (%s-shape), batch %d, k=%d" % (a.recall, a.n, dim, a.shape, a.nq, a.k)


def workload_desc(a, dim):
    """config.workload: the same string in both arms."""
    return ("%s: synthetic %dx%d float L2 (%s-shape), ONNG from the exact kNN graph (knn=%d, reconstructGraph "
            "outgoing=%d incoming=%d, shortcut reduction %s), batch %d queries, k=%d" % (
                "configs[1]" if a.shape == "sift" else "configs[3]" if a.shape == "gist" else "configs[2]", a.n, dim, a.shape, a.knn, a.outgoing, a.incoming, "on" if a.shortcut_reduction else "off", a.nq, a.k))


def index_tag(a, rank=0, world=1):
    return "%s_n%d_k%d_o%d_i%d_s%d_r%dof%d_%s" % (a.shape, a.n, a.knn, a.outgoing, a.incoming, a.shortcut_reduction, rank,
                                                  world, a.mode)


# ---------------------------------------------------------------------------------------------------
def build_index(a, dev, rank, world, want_files):
    """Synthetic rows -> device index with graph + seed table. Returns (GpuIndex, info, index_dir|None)."""
    import torch
    from ngt_b200 import build, engine, index_io, synth
    from ngt_b200 import _lib
    t0 = time.time()
    if a.mode == "shard" and world > 1:
        n_local = a.n // world
        base = synth.make_device(a.shape, n_local, 1000 + rank, dev)
    else:
        n_local = a.n
        base = synth.make_device(a.shape, n_local, 1, dev)
    ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, base.shape[1], device=dev.index or 0)
    ix.set_objects(base)
    ix.set_search_workspace(a.hash_bits, a.queue_cap)
    t1 = time.time()
    ids, dists, counts = build.knn_graph(ix, a.knn)
    torch.cuda.synchronize(dev)
    t2 = time.time()
    row_ptr, col, dist = build.reconstruct_graph(ids, dists, counts, a.outgoing, a.incoming)
    del ids, dists, counts
    torch.cuda.synchronize(dev)
    t2b = time.time()
    adj = None
    if a.shortcut_reduction:
        row_ptr, col, dist, adj = build.adjust_paths(row_ptr, col, dist, 0, with_stats=True)
    stats = build.graph_statistics(row_ptr)
    ix.set_graph(row_ptr, col)
    ix.set_search_property(a.edge_size if a.edge_size > 0 else 0, 30, 20)
    ix.build_seed_table(a.pivots, 1)
    torch.cuda.synchronize(dev)
    t3 = time.time()
    info = {"n": n_local, "gen_s": round(t1 - t0, 2), "knn_graph_s": round(t2 - t1, 2),
            "reconstruct_s": round(t2b - t2, 2), "adjust_paths_s": round(t3 - t2b, 2), "graph": stats}
    if adj:
        info["graph"]["shortcut_reduction"] = {"removed_edges": adj["removed"], "candidates": adj["candidates"]}
    index_dir = None
    if want_files:
        index_dir = a.index_dir or os.path.join(tempfile.gettempdir(), "ngt_b200_bench_" + index_tag(a, rank, world))
        os.makedirs(index_dir, exist_ok=True)
        prop = dict(index_io.DEFAULT_PRF)
        prop.update({"Dimension": str(base.shape[1]), "DistanceType": "L2", "ObjectType": "Float-4",
                     "GraphType": "ONNG", "IndexType": "Graph", "EdgeSizeForSearch": str(a.edge_size),
                     "EdgeSizeForCreation": str(a.knn), "OutgoingEdge": str(a.outgoing), "IncomingEdge": str(a.incoming)})
        index_io.write_prf(index_dir, prop)
        index_io.write_objects(index_dir, base.cpu().numpy())
        index_io.write_graph(index_dir, row_ptr.cpu().numpy().astype(np.uint64), col.cpu().numpy().astype(np.uint32),
                             dist.cpu().numpy())
        info["index_dir"] = index_dir
        info["write_s"] = round(time.time() - t3, 2)
    del base, row_ptr, col, dist
    torch.cuda.empty_cache()
    return ix, info, index_dir


def pick_epsilon(a, ix, q_gt, gt_ids, gt_d):
    """smallest epsilon of the sweep with recall >= target (on the ground-truth subset)."""
    if a.epsilon is not None:
        ids, dists, counts, st = ix.search(q_gt, a.k, a.epsilon, edge_size=a.edge_size, n_seeds=a.seeds, with_stats=True)
        return a.epsilon, recall_at_k(ids, dists, counts, gt_ids, gt_d), []
    curve = []
    for step in range(0, 16):
        eps = round(0.02 * step, 2)
        ids, dists, counts, st = ix.search(q_gt, a.k, eps, edge_size=a.edge_size, n_seeds=a.seeds, with_stats=True)
        rec = recall_at_k(ids, dists, counts, gt_ids, gt_d)
        curve.append({"epsilon": eps, "recall": round(rec, 4), "n_dist": round(float(st[:, 0].mean()), 1)})
        if rec >= a.recall:
            return eps, rec, curve
    return curve[-1]["epsilon"], curve[-1]["recall"], curve


def reference_handle(index_dir):
    from oracle import pyoracle as po
    R = po.Ref()
    h = R.open(index_dir, readonly=True)
    return R, h


def time_reference(R, h, queries, seeds, k, eps, edge_size, steps, warmup):
    """NGT::Index::search from explicit seeds on all host threads; returns (qps, ms/step, threads, recall inputs)."""
    threads = R.max_threads()
    times = []
    out = None
    for s in range(warmup + steps):
        out = R.search(h, queries, k, epsilon=eps, edge_size=edge_size, seeds=seeds, threads=threads, stats=False)
        if s >= warmup:
            times.append(out[4])
    sec = sum(times) / len(times)
    return queries.shape[0] / sec, sec * 1e3, threads, out


def time_port(index_dir, queries, seeds, k, eps, edge_cap, steps, warmup):
    from ngt_b200 import index_io
    from oracle import pyoracle as po
    prop = index_io.read_prf(index_dir)
    rows, _ = index_io.read_objects(index_dir, prop)
    row_ptr, col, _, _ = index_io.read_graph(index_dir)
    port = po.Port()
    pobj, pq = po.pad_objects(rows, po.FLOAT), po.pad_queries(queries, po.FLOAT)
    times = []
    out = None
    for s in range(warmup + steps):
        t = time.perf_counter()
        out = port.graph_search(po.L2, po.FLOAT, pobj, row_ptr, col, pq, seeds, k, eps, edge_size=edge_cap)
        if s >= warmup:
            times.append(time.perf_counter() - t)
    sec = sum(times) / len(times)
    return queries.shape[0] / sec, sec * 1e3, os.cpu_count(), out


# ---------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from ngt_b200 import _lib, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lib.ngtgpu_index_set_timing.argtypes = [C.c_void_p, C.c_int]
    lib.ngtgpu_index_pop_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]

    want_cpu = rank == 0 and world == 1 and not a.no_cpu
    clocks = ClockSampler(local)      # started before the setup: nvidia-smi needs a moment to come up
    clocks.start()
    ix, info, index_dir = build_index(a, dev, rank, world, want_cpu)

    # queries: a few distinct batches so successive steps do not replay the same row set
    n_batches = 4
    q_all = synth.make_device(a.shape, a.nq * n_batches, 2 + 7919 * rank if a.mode == "replica" else 2, dev)
    batches = [q_all[i * a.nq:(i + 1) * a.nq].contiguous() for i in range(n_batches)]
    ngt = min(a.gt_queries, a.nq)
    q_gt = batches[0][:ngt].cpu().numpy()

    if a.mode == "shard" and world > 1:
        from ngt_b200 import sharded
        searcher = sharded.ShardedSearcher(ix, rank, world, info["n"])
        gt_ids, gt_d, _ = searcher.linear_search(batches[0][:ngt], a.k)
        gt_ids, gt_d = gt_ids.cpu().numpy().astype(np.uint32), gt_d.cpu().numpy()

        def run_eps(eps):
            i, d, c = searcher.search(batches[0][:ngt], a.k, eps, a.edge_size, a.seeds)
            return recall_at_k(i.cpu().numpy().astype(np.uint32), d.cpu().numpy(), c.cpu().numpy().astype(np.int64), gt_ids, gt_d)
        eps, rec, curve = None, 0.0, []
        for step in range(16):
            e = round(0.02 * step, 2)
            r = run_eps(e)
            curve.append({"epsilon": e, "recall": round(r, 4)})
            if r >= a.recall:
                eps, rec = e, r
                break
        if eps is None:
            eps, rec = curve[-1]["epsilon"], curve[-1]["recall"]
        t = torch.tensor([eps], device=dev)
        dist.broadcast(t, 0)
        eps = round(float(t.item()), 2)
        step_fn = lambda b: searcher.search(b, a.k, eps, a.edge_size, a.seeds)
    else:
        gt_ids, gt_d, _ = ix.linear_search(q_gt, a.k)
        eps, rec, curve = pick_epsilon(a, ix, q_gt, gt_ids, gt_d)
        if world > 1:
            t = torch.tensor([eps], device=dev)
            dist.broadcast(t, 0)
            eps = round(float(t.item()), 2)
        step_fn = lambda b: ix.search(b, a.k, eps, edge_size=a.edge_size, n_seeds=a.seeds)

    # per-query work at the chosen epsilon (the kernel's own counters) -> algorithmic bytes of one step
    ids, dists, counts, st = ix.search(batches[0], a.k, eps, edge_size=a.edge_size, n_seeds=a.seeds, with_stats=True)
    torch.cuda.synchronize(dev)
    st = st.cpu().numpy().astype(np.int64)
    dim = ix.dimension
    bytes_step = int((st[:, 0] * dim * 4 + st[:, 1] * 4).sum())
    overflow = ix.last_overflows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: device-resident batches, CUDA events on the launching stream
    for w in range(max(a.warmup, 3)):
        step_fn(batches[w % n_batches])
    launches0 = ix.launch_count
    lib.ngtgpu_index_set_timing(ix._h, 1)
    barrier()
    t_begin = clocks.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(a.steps):
        step_fn(batches[s % n_batches])
    ev1.record()
    barrier()
    t_end = clocks.mark()
    ms_total = ev0.elapsed_time(ev1)
    launches = ix.launch_count - launches0
    kms, kcnt = C.c_double(0), C.c_uint64(0)
    lib.ngtgpu_index_pop_timing(ix._h, C.byref(kms), C.byref(kcnt))
    lib.ngtgpu_index_set_timing(ix._h, 0)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / a.steps
    # keep the GPU under the same load a little longer when the timed region was shorter than one sample period
    t_hold = time.time()
    while time.time() - t_hold < 0.35 and (t_end - t_begin) < 0.3:
        ix.search(batches[0], a.k, eps, edge_size=a.edge_size, n_seeds=a.seeds)   # rank-local: no collective in here
        torch.cuda.synchronize(dev)
        t_end = clocks.mark()
    clock_info = clocks.stop(t_begin, t_end)
    units = a.nq * (world if a.mode == "replica" else 1)
    value = units / (ms_step / 1e3)

    # ---- e2e: the host-pointer C-ABI call, pinned host queries in, host results out
    hq = [torch.empty((a.nq, dim), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        hq[i].copy_(batches[i])
    torch.cuda.synchronize(dev)
    hqn = [h.numpy() for h in hq]
    e2e_fn = lambda i: ix.search(hqn[i % 2], a.k, eps, edge_size=a.edge_size, n_seeds=a.seeds)
    for w in range(a.warmup):
        e2e_fn(w)
    barrier()
    t0 = time.perf_counter()
    for s in range(a.steps):
        e2e_fn(s)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = units / (e2e_ms / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    k_ms = kms.value / max(kcnt.value, 1)
    achieved = bytes_step / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    out = {
        "metric": metric_name(a, dim),
        "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(a, dim),
                   "epsilon": eps, "edge_size": a.edge_size, "recall_at_10": round(rec, 4), "recall_queries": ngt,
                   "seeds": "nearest %d of %d device pivots" % (a.seeds, a.pivots),
                   "parallelism": ("replica x%d (one 10k batch per GPU)" % world) if a.mode == "replica" else
                                  ("rows sharded x%d + all-gather merge" % world),
                   "l2_policy": "inputs larger than L2 (%d MB of rows vs 126 MB), 4 rotating query batches" % (info["n"] * dim * 4 >> 20),
                   "graph": info["graph"], "setup_s": {k: info[k] for k in ("gen_s", "knn_graph_s", "reconstruct_s", "adjust_paths_s")},
                   "overflow_queries_per_step": overflow, "epsilon_sweep": curve},
        "e2e": {"value": round(e2e_value, 1), "unit": "queries/s", "ms_per_step": round(e2e_ms, 4),
                "h2d_bytes_per_step": a.nq * dim * 4, "d2h_bytes_per_step": a.nq * a.k * 8 + a.nq * 4},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {"kernel": ("search_fast_kernel<F_L2,CH%d> (graph traversal)" % (1 if dim <= 32 else 2 if dim <= 64 else 4)) if 16 < dim <= 128 and a.edge_size <= 128 and 0 < a.edge_size and a.k <= 32 else
                     "search_kernel<F_L2,G32,CPL%d> (graph traversal)" % (1 if dim <= 128 else 2 if dim <= 256 else 4 if dim <= 512 else 8 if dim <= 1024 else 0), "bound": "hbm",
                     "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None,
                     "kernel_ms": round(k_ms, 4), "algorithmic_bytes_per_launch": bytes_step,
                     "n_dist_per_query": round(float(st[:, 0].mean()), 1), "n_edge_per_query": round(float(st[:, 1].mean()), 1),
                     "n_expanded_per_query": round(float(st[:, 2].mean()), 1)},
    }
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and a.shape == "sift" and a.n == 1000000 and a.mode == "replica":   # the capture is of configs[1]
        try:
            out["roofline"]["traffic"] = json.load(open(tp)).get("search_kernel_dram_bytes_per_launch")
        except Exception:
            pass

    # ---- cpu_baseline: the reference itself on the same index files / seeds / epsilon, bounded sample
    if want_cpu:
        try:
            m = min(a.cpu_sample, a.nq)
            qs = hqn[0][:m].copy()
            seeds = np.zeros((m, a.seeds), np.uint32)
            _lib.check(lib.ngtgpu_select_seeds(ix._h, qs.ctypes.data, _lib.OBJECT_FLOAT, m, a.seeds, seeds.ctypes.data))
            kind = "reference"
            try:
                R, h = reference_handle(index_dir)
                qps, ms, threads, o = time_reference(R, h, qs, seeds, a.k, eps, a.edge_size, 2, 1)
                R.close(h)
                rids, rd, rc = o[0], o[1], o[2]
            except (FileNotFoundError, OSError):
                kind = "port"
                cap = 2 ** 31 - 1 if a.edge_size == 0 else a.edge_size
                qps, ms, threads, o = time_port(index_dir, qs, seeds, a.k, eps, cap, 2, 1)
                rids, rd, rc = o[0], o[1], o[2]
            gi, gd, gc = ix.search(qs, a.k, eps, edge_size=a.edge_size, seeds=seeds)
            same = bool((gi == rids).all() and (gc == rc).all() and (gd.view(np.uint32) == rd.view(np.uint32)).all())
            gt_i, gt_dd, _ = ix.linear_search(qs, a.k)
            out["cpu_baseline"] = {"value": round(qps, 1), "unit": "queries/s", "cores": int(threads), "kind": kind,
                                   "sample": "%d queries of the batch, same index files, same seeds, epsilon %.2f, "
                                             "OpenMP over queries" % (m, eps),
                                   "recall_at_10": round(recall_at_k(rids, rd, rc.astype(np.int64), gt_i, gt_dd), 4),
                                   "gpu_results_identical": same}
        except Exception as ex:  # the baseline must never take the GPU number down with it
            out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference",
                                   "sample": "failed: %s" % (str(ex)[:200])}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(a):
    """The reference arm: NGT::Index::search (unmodified reference, oracle/_ref) on the host cores. The index is the
    same one the GPU arm searches (built in the untimed setup and written in NGT's own file format); the timed
    region runs only reference code."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from ngt_b200 import _lib
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib = _lib.load()
    lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    a.mode = "replica"
    ix, info, index_dir = build_index(a, dev, 0, 1, True)
    from ngt_b200 import synth
    q_all = synth.make_device(a.shape, a.nq, 2, dev).cpu().numpy()
    ngt = min(a.gt_queries, a.nq)
    gt_ids, gt_d, _ = ix.linear_search(q_all[:ngt], a.k)
    eps, rec, curve = pick_epsilon(a, ix, q_all[:ngt], gt_ids, gt_d)
    m = min(a.cpu_sample, a.nq)
    qs = np.ascontiguousarray(q_all[:m])
    seeds = np.zeros((m, a.seeds), np.uint32)
    _lib.check(lib.ngtgpu_select_seeds(ix._h, qs.ctypes.data, _lib.OBJECT_FLOAT, m, a.seeds, seeds.ctypes.data))
    ix.close()
    kind = "reference"
    try:
        R, h = reference_handle(index_dir)
        qps, ms, threads, o = time_reference(R, h, qs, seeds, a.k, eps, a.edge_size, a.steps, a.warmup)
        R.close(h)
    except (FileNotFoundError, OSError):
        kind = "port"
        cap = 2 ** 31 - 1 if a.edge_size == 0 else a.edge_size
        qps, ms, threads, o = time_port(index_dir, qs, seeds, a.k, eps, cap, a.steps, a.warmup)
    rrec = recall_at_k(o[0][:ngt], o[1][:ngt], o[2][:ngt].astype(np.int64), gt_ids[:m], gt_d[:m]) if m >= ngt else \
        recall_at_k(o[0], o[1], o[2].astype(np.int64), gt_ids[:m], gt_d[:m])
    out = {
        "impl": "reference",
        "metric": metric_name(a, qs.shape[1]),
        "value": round(qps, 1), "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(a, qs.shape[1]), "index": "the index files the GPU arm searches",
                   "epsilon": eps, "edge_size": a.edge_size, "recall_at_10": round(rrec, 4)},
        "cpu_baseline": {"value": round(qps, 1), "unit": "queries/s", "cores": int(threads), "kind": kind,
                         "sample": "each step = %d queries of the 10k batch (NGT::Index::search, OpenMP over queries, "
                                   "explicit seeds identical to the GPU arm's)" % m},
        "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

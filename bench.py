#!/usr/bin/env python
"""bench.py -- QPS of batched graph search at recall@10 >= 0.95 on synthetic 1M x 128 float L2 (sift-shape),
batch 10k, k = 10 (BASELINE.json `metric`, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this engine (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU search on the host cores

One "step" = one batch of 10k queries through the hot path (seed selection + graph traversal, k = 10) at the
smallest epsilon of the sweep 0.00..0.30 (step 0.02) whose recall@10 (Optimizer.h:496-507 definition, ground
truth from the exhaustive scan) is >= 0.95.

  value     QPS with the query batch already in HBM (CUDA events on the launching stream, max over ranks)
  e2e       QPS through the reference-facing C API, ngt_batch_search_index_as_float() on an index opened with
            ngt_open_index() from NGT's own files: pinned host queries in, host results out, both copies inside the
            timed region; two host threads submit alternate batches (the reference's threading contract: concurrent
            searches on a read-only index), so the upload of one batch runs under the traversal of the other
  roofline  traversal kernel only: algorithmic bytes (n_dist * D * 4 + n_edge * 4 per query, from the kernel's
            own counters) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference (oracle/_ref, NGT::Index::search on the SAME index files, same seeds,
            same epsilon) on all host threads, on a bounded sample of the batch; plus `reference_recipe`: the
            reference on ITS OWN settings (ANNG -E 100 -> ONNG -o 10 -i 120 built by the reference, DVP-tree seeds,
            dynamic edge-size mode) on a 100k subset, with this engine on the same index files beside it
  sharded   BASELINE configs[4]: 12.5M x 128 uint8 rows PER GPU (100M at N = 8), rows sharded over the N ranks, a graph
            per shard, ONE ncclAllGather inside libngtgpu.so, device merge; recall on the union, ms breakdown
  workloads configs[2] (glove-shape, rank 0), configs[3] (gist-shape, rows sharded over the N ranks), the Hamming
            variant of configs[4]

Setup (untimed, identical for both arms): synthetic data (ngt_b200.synth, seeds 1/2), the reference's ONNG recipe on
an exact kNN table, all on the device (ngtgpu_index_build_onng: kNN pass on the tensor cores, reconstructGraph -o 10
-i 120 + shortcut reduction, GraphReconstructor.h:425-561 and 197-386), seed table. N > 1 runs one replica of the
index per GPU with its own 10k batch for the headline (weak scaling, no data-path collective); the `sharded` record of
the same line is the row-sharded path with its collective.
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
KNN_BATCH = 148 * 256 * 3      # query batch of ngtgpu_index_build_onng's exhaustive pass (three waves of 256-query CTAs on 148 SMs)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="replica", choices=["replica"])
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
    ap.add_argument("--extras", default="sharded,hamming,glove,gist",
                    help="sub-records added to the line: sharded (configs[4]), hamming, glove (configs[2]), gist (configs[3]); 'none'")
    ap.add_argument("--shard-rows", type=int, default=12500000, help="uint8 rows per GPU of the sharded record")
    ap.add_argument("--ham-rows", type=int, default=2000000, help="128-bit rows per GPU of the Hamming record")
    ap.add_argument("--glove-n", type=int, default=1200000)
    ap.add_argument("--gist-n", type=int, default=0, help="gist rows in total (default: 250k on one GPU, 1M sharded over N > 1)")
    ap.add_argument("--recipe-n", type=int, default=100000, help="objects of the reference_recipe leg (0: skip)")
    ap.add_argument("--profile-region", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off lists exactly the step's launches)")
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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


def measured_peak_gbs():
    m = measured_peaks()
    if "hbm_gbs" in m:
        return float(m["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_bf16_tflops():
    """(burst, sustained) dense bf16 TFLOP/s of this pool's B200s."""
    m = measured_peaks()
    if "bf16_tflops" in m:
        return float(m["bf16_tflops"]), float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 1650.0, 1400.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """The host threads this process may use: its CPU affinity, not OMP_NUM_THREADS (torchrun sets that to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def metric_name(a, dim):
    """BASELINE.json's metric on configs[1]; other shapes (development runs) say what they are."""
    if a.shape == "sift" and a.n == 1000000:
        return "QPS at recall@10>=0.95, synthetic 1Mx128 float L2 (sift-shape), batch 10k, k=10"
    return "QPS at recall@10>=%.2f, synthetic %dx%d float L2 (%s-shape), batch %d, k=%d" % (a.recall, a.n, dim, a.shape, a.nq, a.k)


def workload_desc(a, dim):
    """config.workload: the same string in both arms."""
    return ("%s: synthetic %dx%d float L2 (%s-shape), ONNG from the exact kNN graph (knn=%d, reconstructGraph "
            "outgoing=%d incoming=%d, shortcut reduction %s), batch %d queries, k=%d" % (
                "configs[1]" if a.shape == "sift" else "configs[3]" if a.shape == "gist" else "configs[2]", a.n, dim, a.shape, a.knn, a.outgoing, a.incoming, "on" if a.shortcut_reduction else "off", a.nq, a.k))


def index_tag(a):
    return "%s_n%d_k%d_o%d_i%d_s%d" % (a.shape, a.n, a.knn, a.outgoing, a.incoming, a.shortcut_reduction)


def bind_capi(so_path):
    """The few lib/NGT/Capi.h entry points the e2e leg drives (python/ngt/base.py binds them the same way)."""
    lib = C.CDLL(so_path)
    P = C.c_void_p
    lib.ngt_create_error_object.restype = P
    lib.ngt_get_error_string.restype, lib.ngt_get_error_string.argtypes = C.c_char_p, [P]
    lib.ngt_destroy_error_object.argtypes = [P]
    lib.ngt_open_index.restype, lib.ngt_open_index.argtypes = P, [C.c_char_p, P]
    lib.ngt_close_index.argtypes = [P]
    lib.ngt_create_empty_results.restype, lib.ngt_create_empty_results.argtypes = P, [P]
    lib.ngt_destroy_results.argtypes = [P]
    lib.ngt_search_index_as_float.restype = C.c_bool
    lib.ngt_search_index_as_float.argtypes = [P, P, C.c_int32, C.c_size_t, C.c_float, C.c_float, P, P]
    lib.ngt_batch_search_index_as_float.restype = C.c_bool
    lib.ngt_batch_search_index_as_float.argtypes = [P, P, C.c_uint32, C.c_int32, C.c_size_t, C.c_float, C.c_float, C.c_int64, P, P, P, P]
    return lib


# ---------------------------------------------------------------------------------------------------
def build_index(a, dev, want_files):
    """Synthetic rows -> device index with graph + seed table. Returns (GpuIndex, info, index_dir|None)."""
    import torch
    from ngt_b200 import build, engine, index_io, synth
    from ngt_b200 import _lib
    t0 = time.time()
    base = synth.make_device(a.shape, a.n, 1, dev)
    ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, base.shape[1], device=dev.index or 0)
    ix.set_objects(base)
    ix.set_search_workspace(a.hash_bits, a.queue_cap)
    torch.cuda.synchronize(dev)
    t1 = time.time()
    tc0 = ix.tensor_core_batches
    b = ix.build_onng(a.knn, a.outgoing, a.incoming, bool(a.shortcut_reduction), 0, want_graph=want_files)
    ix.set_search_property(a.edge_size if a.edge_size > 0 else 0, 30, 20)
    ix.build_seed_table(a.pivots, 1)
    torch.cuda.synchronize(dev)
    row_ptr = b["graph"][0] if want_files else None
    dim = base.shape[1]
    bf_burst, bf_sus, _ = measured_bf16_tflops()
    tf = 2.0 * a.n * a.n * dim / max(b["knn_s"], 1e-9) / 1e12
    info = {"n": a.n, "gen_s": round(t1 - t0, 2), "knn_graph_s": round(b["knn_s"], 3),
            "reconstruct_s": round(b["reconstruct_s"], 3), "adjust_paths_s": round(b["adjust_paths_s"], 3),
            "knn_pass": {"useful_tflops": round(tf, 1), "frac_of_bf16_sustained": round(tf / bf_sus, 4),
                         "tensor_core_batches": "%d of %d" % (ix.tensor_core_batches - tc0, -(-a.n // KNN_BATCH))}}
    index_dir = None
    if want_files:
        info["graph"] = build.graph_statistics(row_ptr)
        t3 = time.time()
        index_dir = a.index_dir or os.path.join(tempfile.gettempdir(), "ngt_b200_bench_" + index_tag(a))
        os.makedirs(index_dir, exist_ok=True)
        prop = dict(index_io.DEFAULT_PRF)
        prop.update({"Dimension": str(dim), "DistanceType": "L2", "ObjectType": "Float-4",
                     "GraphType": "ONNG", "IndexType": "Graph", "EdgeSizeForSearch": str(a.edge_size),
                     "EdgeSizeForCreation": str(a.knn), "OutgoingEdge": str(a.outgoing), "IncomingEdge": str(a.incoming),
                     "SeedSize": str(a.seeds)})
        index_io.write_prf(index_dir, prop)
        index_io.write_objects(index_dir, base.cpu().numpy())
        g = b["graph"]
        index_io.write_graph(index_dir, g[0].cpu().numpy().astype(np.uint64), g[1].cpu().numpy().astype(np.uint32), g[2].cpu().numpy())
        info["index_dir"] = index_dir
        info["write_s"] = round(time.time() - t3, 2)
    del base, b
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
    R = po.Ref("native")
    h = R.open(index_dir, readonly=True)
    return R, h


def time_reference(R, h, queries, seeds, k, eps, edge_size, steps, warmup):
    """NGT::Index::search from explicit seeds on all host threads; returns (qps, ms/step, threads, recall inputs)."""
    threads = host_threads()      # explicit: under torchrun OMP_NUM_THREADS=1 would otherwise leave the reference one core
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
    return queries.shape[0] / sec, sec * 1e3, host_threads(), out


# ---------------------------------------------------------------------------------------------------
# sub-records: BASELINE configs[2..4] under the same clock (rows sharded over the ranks where the config says so)
def _all_reduce(t, op, world):
    import torch.distributed as dist
    if world > 1:
        dist.all_reduce(t, op=op)
    return t


def make_rows(shape, kind, count, seed, dev):
    """rows of one workload on the device, generated in pieces: kind f32 | u8 (rounded, clipped) | ham (bit j = [x_j > 64])."""
    import torch
    from ngt_b200 import synth
    out, step = [], 2000000
    for s in range(0, count, step):
        m = min(step, count - s)
        x = synth.make_device(shape, m, seed * 1000003 + s if count > step else seed, dev)
        if kind == "u8":
            x = x.to(torch.uint8)
        elif kind == "ham":
            bits = (x > 64.0).to(torch.uint8).reshape(m, -1, 8)
            wts = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=dev)
            x = (bits * wts).sum(-1).to(torch.uint8)
        out.append(x)
    return torch.cat(out) if len(out) > 1 else out[0]


def union_check_integer(kind, base, id_offset, queries, m_ids, m_d, m_c, world, dev):
    """merged result == linearSearch over the union, proved without materialising the union: for every checked query,
    (a) the rows of ALL shards whose (distance, id) key is <= the merged list's last key number exactly as many as the
    list holds, and (b) every merged entry's distance equals the integer formula on its own row, bit for bit.
    (a)+(b) => the list is the k smallest keys of the union. Distances are recomputed here with torch integer arithmetic,
    independently of the engine's kernels. -> (ok, queries checked)"""
    import torch
    import torch.distributed as dist
    n_local = base.shape[0]
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    pop = None
    if kind == "ham":
        pop = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=dev)
    for q in range(queries.shape[0]):
        cnt = int(m_c[q])
        d_parts = []
        for s in range(0, n_local, 1 << 20):
            blk = base[s:s + (1 << 20)]
            if kind == "ham":
                d_parts.append(pop[torch.bitwise_xor(blk, queries[q][None, :]).long()].sum(1).float())
            else:
                diff = blk.to(torch.int32) - queries[q][None, :].to(torch.int32)
                d_parts.append(torch.sqrt((diff * diff).sum(1).double()).float())
        d = torch.cat(d_parts)
        gid = torch.arange(1, n_local + 1, device=dev, dtype=torch.int64) + id_offset
        kd, kid = m_d[q, cnt - 1], m_ids[q, cnt - 1].long()
        below = ((d < kd) | ((d == kd) & (gid <= kid))).sum().to(torch.int64).reshape(1)
        _all_reduce(below, dist.ReduceOp.SUM, world)
        if int(below) != cnt:
            ok.zero_()
        mine = (m_ids[q, :cnt].long() > id_offset) & (m_ids[q, :cnt].long() <= id_offset + n_local)
        loc = (m_ids[q, :cnt].long()[mine] - id_offset - 1)
        if not bool((d[loc].view(torch.int32) == m_d[q, :cnt][mine].view(torch.int32)).all()):
            ok.zero_()
    _all_reduce(ok, dist.ReduceOp.MIN, world)
    return bool(int(ok)), int(queries.shape[0])


def run_workload(spec, a, dev, rank, world, lib):
    """One configuration, rows sharded over `world` ranks (world == 1: one GPU): per-shard ONNG built on the device,
    search through libngtgpu.so's sharded entry point (ngtgpu_shard_search_device: keys written by the traversal kernel,
    ONE ncclAllGather, device merge), recall on the union, CUDA-event timing between barriers, max over ranks."""
    import torch
    import torch.distributed as dist
    from ngt_b200 import _lib, engine, sharded
    kind, n_local, nq, k = spec["kind"], spec["n_local"], a.nq, a.k
    otype = _lib.OBJECT_FLOAT if kind == "f32" else _lib.OBJECT_UINT8
    t0 = time.time()
    base = make_rows(spec["shape"], kind, n_local, 1000 + rank if world > 1 else 1, dev)
    dim = base.shape[1]
    elem = 4 if kind == "f32" else 1
    ix = engine.GpuIndex(otype, spec["distance"], dim, device=dev.index or 0)
    ix.set_objects(base)
    if spec.get("hash_bits"):
        ix.set_search_workspace(spec["hash_bits"], spec.get("queue_cap", 512))
    torch.cuda.synchronize(dev)
    t1 = time.time()
    tc0 = ix.tensor_core_batches
    b = ix.build_onng(spec["knn"], spec["outgoing"], spec["incoming"], True, 0)
    cap = spec["edge_size"]
    ix.set_search_property(cap, 30, 20)
    ix.build_seed_table(spec.get("pivots", 256), 1)
    torch.cuda.synchronize(dev)
    tc_batches = ix.tensor_core_batches - tc0
    S = sharded.LibShardedSearcher(ix, rank, world, rank * n_local)
    nb = 3
    q_all = make_rows(spec["shape"], kind, nq * nb, 2, dev)      # the same queries on every rank
    batches = [q_all[i * nq:(i + 1) * nq].contiguous() for i in range(nb)]
    ngt = min(spec.get("gt", 1000), nq)
    gq = batches[0][:ngt]
    gt_ids, gt_d, gt_c = S.linear_search(gq, k)
    # merged exhaustive answer == linearSearch on the union
    if kind in ("u8", "ham"):
        exact, checked = union_check_integer(kind, base, rank * n_local, gq[:spec.get("union_queries", 8)], gt_ids, gt_d, gt_c, world, dev)
        how = "count of rows under the k-th key over all shards + integer formula, bit for bit"
    else:
        # float rows: the library's all-gather + merge against torch.distributed's gather + a host merge of the per-shard lists
        li, ld, lc = ix.linear_search(gq, k)
        keys = torch.from_numpy(sharded.pack_keys_host(li.cpu().numpy(), ld.cpu().numpy(), lc.cpu().numpy(), rank * n_local).view(np.int64)).to(dev)
        gathered = sharded.all_gather_keys(keys, world) if world > 1 else keys[None]
        mi, md, mc = sharded.merge_keys_host(gathered.cpu().numpy().view(np.uint64), k)
        exact = bool((mi == gt_ids.cpu().numpy().astype(np.uint32)).all() and (md.view(np.uint32) == gt_d.cpu().numpy().view(np.uint32)).all())
        checked, how = ngt, "host merge of the per-shard exhaustive lists (torch.distributed gather), bit for bit"
    del base
    gt_ids_h, gt_d_h = gt_ids.cpu().numpy().astype(np.uint32), gt_d.cpu().numpy()
    curve, eps, rec = [], None, 0.0
    for step in range(0, 21):
        e = round(0.02 * step, 2)
        r = S.search(gq, k, e, edge_size=cap, n_seeds=a.seeds)
        rc = recall_at_k(r[0].cpu().numpy().astype(np.uint32), r[1].cpu().numpy(), r[2].cpu().numpy().astype(np.int64), gt_ids_h, gt_d_h)
        curve.append({"epsilon": e, "recall": round(rc, 4)})
        if rc >= a.recall:
            eps, rec = e, rc
            break
    if eps is None:
        eps, rec = curve[-1]["epsilon"], curve[-1]["recall"]
    # this shard's work per query (the kernel's own counters) -> algorithmic bytes of one batch
    r = ix.search(batches[0], k, eps, edge_size=cap, n_seeds=a.seeds, with_stats=True)
    st = r[3].cpu().numpy().astype(np.int64)
    bytes_step = int((st[:, 0] * dim * elem + st[:, 1] * 4).sum())
    for w in range(max(a.warmup, 3)):
        S.search(batches[w % nb], k, eps, edge_size=cap, n_seeds=a.seeds)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    lib.ngtgpu_index_set_timing(ix._h, 1)
    S.set_timing(True)
    steps = max(a.steps, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        S.search(batches[s % nb], k, eps, edge_size=cap, n_seeds=a.seeds)
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    kms, kcnt = C.c_double(0), C.c_uint64(0)
    lib.ngtgpu_index_pop_timing(ix._h, C.byref(kms), C.byref(kcnt))
    lib.ngtgpu_index_set_timing(ix._h, 0)
    cms, calls = S.pop_timing()
    S.set_timing(False)
    k_ms = kms.value / max(kcnt.value, 1)
    per = {n: v / max(calls, 1) for n, v in cms.items()}
    red = torch.tensor([e0.elapsed_time(e1) / steps, k_ms, max(per["search_ms"] - k_ms, 0.0), per["allgather_ms"], per["merge_ms"],
                        b["knn_s"], b["reconstruct_s"] + b["adjust_paths_s"], t1 - t0], device=dev, dtype=torch.float64)
    _all_reduce(red, dist.ReduceOp.MAX, world)
    gbs = torch.tensor([bytes_step / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0], device=dev, dtype=torch.float64)
    gbs_min = _all_reduce(gbs.clone(), dist.ReduceOp.MIN, world)
    gbs_max = _all_reduce(gbs.clone(), dist.ReduceOp.MAX, world)
    S.close()
    ix.close()
    del q_all, batches
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, _ = measured_peak_gbs()
    bf_burst, bf_sus, _ = measured_bf16_tflops()
    ms, knn_s = float(red[0]), float(red[5])
    kdim = dim * 8 if kind == "ham" else dim
    tf = 2.0 * n_local * n_local * kdim / max(knn_s, 1e-9) / 1e12
    n_batches = -(-n_local // KNN_BATCH)
    return {
        "workload": spec["desc"] % {"total": n_local * world, "world": world, "n_local": n_local},
        "n_gpus": world, "objects": n_local * world, "objects_per_gpu": n_local, "batch": nq, "k": k, "edge_size": cap,
        "epsilon": eps, "recall_at_10": round(rec, 4), "gt": "exhaustive scan of every shard, merged in the library", "gt_queries": ngt,
        "qps": round(nq / ms * 1e3, 1), "ms_per_batch": round(ms, 4),
        "ms_breakdown_max_over_ranks": {"seeds_and_query_preparation": round(float(red[2]), 4), "traversal_kernel": round(float(red[1]), 4),
                                        "pack": 0.0, "allgather": round(float(red[3]), 4), "merge": round(float(red[4]), 4),
                                        "note": "pack is fused into the traversal kernel's result write; the all-gather includes the wait for the slowest shard"},
        "merged_linear_search_equals_union": exact, "union_check": "%d queries: %s" % (checked, how),
        "roofline": {"bound": "hbm", "achieved_min_over_ranks": round(float(gbs_min), 1), "achieved_max_over_ranks": round(float(gbs_max), 1),
                     "peak": peak, "unit": "GB/s", "frac": round(float(gbs_min) / peak, 4), "algorithmic_bytes_per_launch_rank0": bytes_step,
                     "n_dist_per_query_rank0": round(float(st[:, 0].mean()), 1), "n_edge_per_query_rank0": round(float(st[:, 1].mean()), 1)},
        "knn_pass": {"k": spec["knn"], "seconds_max_over_ranks": round(knn_s, 2), "useful_tflops_per_gpu": round(tf, 1),
                     "frac_of_bf16_sustained": round(tf / bf_sus, 4), "bf16_tflops_sustained": bf_sus,
                     "operands": ("u8 x u8 -> s32 on tcgen05 kind::i8 (its dense rate is twice the bf16 one the fraction is quoted against)"
                                  if spec["kind"] in ("u8", "ham") else "bf16 (split floats: three segments) on tcgen05 kind::f16"),
                     "tensor_core_batches_rank0": "%d of %d" % (tc_batches, n_batches)},
        "setup_s_max_over_ranks": {"objects": round(float(red[7]), 2), "knn_pass": round(knn_s, 2), "onng": round(float(red[6]), 2)},
        "epsilon_sweep": curve,
    }


def extra_records(a, dev, rank, world, lib):
    """-> (sharded record | None, {name: record}) on rank 0; (None, {}) elsewhere."""
    import torch.distributed as dist
    from ngt_b200 import _lib
    want = [w for w in a.extras.split(",") if w and w != "none"]
    sharded_rec, wl = None, {}

    def guarded(name, fn):
        try:
            return fn()
        except Exception as ex:      # an extra must never take the headline down with it
            return {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])} if rank == 0 else None

    if "sharded" in want:
        spec = {"kind": "u8", "shape": "sift", "distance": _lib.DISTANCE_L2, "n_local": a.shard_rows, "knn": 64, "outgoing": 10,
                "incoming": 64, "edge_size": 64, "pivots": 256,
                "desc": "configs[4]: synthetic %(total)dx128 uint8 L2, rows sharded over %(world)d x B200 (%(n_local)d per GPU), per-shard ONNG, "
                        "one ncclAllGather + device merge in libngtgpu.so"}
        sharded_rec = guarded("sharded", lambda: run_workload(spec, a, dev, rank, world, lib))
    if "hamming" in want:
        spec = {"kind": "ham", "shape": "sift", "distance": _lib.DISTANCE_HAMMING, "n_local": a.ham_rows, "knn": 64, "outgoing": 10,
                "incoming": 64, "edge_size": 64, "pivots": 256,
                "desc": "configs[4], Hamming variant: %(total)d x 128-bit objects, rows sharded over %(world)d x B200 (%(n_local)d per GPU)"}
        wl["hamming"] = guarded("hamming", lambda: run_workload(spec, a, dev, rank, world, lib))
    if "gist" in want:
        total = a.gist_n or (250000 if world == 1 else 1000000)
        spec = {"kind": "f32", "shape": "gist", "distance": _lib.DISTANCE_L2, "n_local": total // world, "knn": 64, "outgoing": 10,
                "incoming": 64, "edge_size": 64, "pivots": 256, "hash_bits": 16, "queue_cap": 2048,
                "desc": "configs[3]: synthetic %(total)dx960 float L2 (gist-shape), rows sharded over %(world)d x B200 (%(n_local)d per GPU)"
                        + (" -- one GPU: a 250k shard of the 1M set" if world == 1 else "")}
        wl["gist"] = guarded("gist", lambda: run_workload(spec, a, dev, rank, world, lib))
    if "glove" in want:
        spec = {"kind": "f32", "shape": "glove", "distance": _lib.DISTANCE_NORMALIZED_COSINE, "n_local": a.glove_n, "knn": 64,
                "outgoing": 10, "incoming": 64, "edge_size": 64, "pivots": 256,
                "desc": "configs[2]: synthetic %(total)dx100 float normalized cosine (glove-shape), one B200, exhaustive kNN graph on the tensor cores"}
        if rank == 0:      # one GPU by definition: the other ranks wait at the barrier below
            wl["glove"] = guarded("glove", lambda: run_workload(spec, a, dev, 0, 1, lib))
        if world > 1:
            dist.barrier()
    return (sharded_rec, wl) if rank == 0 else (None, {})


# ---------------------------------------------------------------------------------------------------
def reference_recipe(a, dev, k, n_queries=1000):
    """The reference on its OWN settings, beside this engine on the same index files: ANNG `-E 100 -S 40` built by
    NGT::Index::createIndex, ONNG by GraphOptimizer::execute `-o 10 -i 120` with path adjustment (both reference code, all
    host threads), DVP-tree seeds, the dynamic edge-size mode `ngt reconstruct-graph` leaves in prf for this data
    (EdgeSizeForSearch -2, base 32, rate 8: SURVEY.md Appendix A; its timed tuning step is not re-run), and for each side
    the smallest epsilon of the sweep with recall@10 >= 0.95. A bound on how much of the headline ratio is the graph /
    edge-size choice of the GPU arm; not a target."""
    import shutil
    import torch
    from ngt_b200 import _lib, engine, index_io, synth
    from oracle import pyoracle as po
    n = a.recipe_n
    R = po.Ref("native")
    threads = host_threads()
    base = synth.make("sift", n, 1)
    qs = synth.make("sift", n_queries, 2)
    tmp = tempfile.mkdtemp(prefix="ngt_b200_recipe_")
    try:
        anng, onng = os.path.join(tmp, "anng"), os.path.join(tmp, "onng")
        t0 = time.time()
        R.build_index(anng, base, objtype="f", disttype=po.L2, edge_creation=100, edge_search=40, indextype="t", threads=threads)
        t1 = time.time()
        R.build_onng(anng, onng, outgoing=10, incoming=120, shortcut=True)
        t2 = time.time()
        lines = open(os.path.join(onng, "prf")).read().splitlines()
        sub = {"EdgeSizeForSearch": "-2", "DynamicEdgeSizeBase": "32", "DynamicEdgeSizeRate": "8"}
        open(os.path.join(onng, "prf"), "w").write("\n".join((l.split("\t")[0] + "\t" + sub[l.split("\t")[0]]) if l.split("\t")[0] in sub else l
                                                              for l in lines) + "\n")
        # ground truth (untimed) from the engine's exhaustive scan
        prop = index_io.read_prf(onng)
        rows, _ = index_io.read_objects(onng, prop)
        row_ptr, col, _, _ = index_io.read_graph(onng)
        ix = engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, rows.shape[1], device=dev.index or 0)
        ix.set_objects(rows)
        ix.set_graph(row_ptr[:n + 2], col)
        ix.set_search_property(-2, 32, 8)
        ix.build_seed_table(a.pivots, 1)
        gt_ids, gt_d, _ = ix.linear_search(qs, k)
        h = R.open(onng, readonly=True)
        ref = None
        for step in range(16):
            e = round(0.02 * step, 2)
            o = R.search(h, qs, k, epsilon=e, edge_size=-1, seeds=None, threads=threads, stats=False)     # tree seeds, prf edge mode
            rc = recall_at_k(o[0], o[1], o[2].astype(np.int64), gt_ids, gt_d)
            if rc >= a.recall or step == 15:
                secs = [R.search(h, qs, k, epsilon=e, edge_size=-1, seeds=None, threads=threads, stats=False)[4] for _ in range(3)]
                ref = {"epsilon": e, "recall_at_10": round(rc, 4), "qps": round(n_queries / min(secs), 1), "seeds": "DVP-tree",
                       "edge_size": "-2 (dynamic: 32 + 10^(8 eps))"}
                break
        R.close(h)
        ours = None
        qd = torch.from_numpy(np.tile(qs, (10, 1))).to(dev)       # a 10k batch (the 1000 queries ten times)
        for step in range(16):
            e = round(0.02 * step, 2)
            i, d, c = ix.search(qs, k, e, edge_size=-1, n_seeds=a.seeds)
            rc = recall_at_k(i, d, c, gt_ids, gt_d)
            if rc >= a.recall or step == 15:
                for _ in range(3):
                    ix.search(qd, k, e, edge_size=-1, n_seeds=a.seeds)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    ix.search(qd, k, e, edge_size=-1, n_seeds=a.seeds)
                e1.record()
                torch.cuda.synchronize(dev)
                ours = {"epsilon": e, "recall_at_10": round(rc, 4), "qps": round(qd.shape[0] * 5 / e0.elapsed_time(e1) * 1e3, 1),
                        "seeds": "nearest %d of %d device pivots" % (a.seeds, a.pivots), "edge_size": "-2 (same prf)"}
                break
        ix.close()
        return {"objects": n, "queries": n_queries, "index": "ANNG -E 100 (createIndex) -> ONNG -o 10 -i 120 + path adjustment (GraphOptimizer::execute), "
                "both built by the unmodified reference", "reference_build_s": {"anng": round(t1 - t0, 1), "onng": round(t2 - t1, 1)},
                "cores": threads, "isa": R.isa, "reference": ref, "this_engine_on_the_same_index_files": ours,
                "ratio": round(ours["qps"] / ref["qps"], 1) if ours and ref else None}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


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
    ix, info, index_dir = build_index(a, dev, rank == 0)      # rank 0 also writes the index in NGT's own file format

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # queries: a few distinct batches so successive steps do not replay the same row set
    n_batches = 4
    q_all = synth.make_device(a.shape, a.nq * n_batches, 2 + 7919 * rank, dev)
    batches = [q_all[i * a.nq:(i + 1) * a.nq].contiguous() for i in range(n_batches)]
    ngt = min(a.gt_queries, a.nq)
    q_gt = batches[0][:ngt].cpu().numpy()
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
    ix.search(batches[0].cpu().numpy(), a.k, eps, edge_size=a.edge_size, n_seeds=a.seeds)    # host-pointer call: reports the overflow count
    overflow = ix.last_overflows

    # ---- value: device-resident batches, CUDA events on the launching stream
    for w in range(max(a.warmup, 3)):
        step_fn(batches[w % n_batches])
    launches0 = ix.launch_count
    lib.ngtgpu_index_set_timing(ix._h, 1)
    barrier()
    t_begin = clocks.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.profile_region:
        torch.cuda.profiler.start()
    ev0.record()
    for s in range(a.steps):
        step_fn(batches[s % n_batches])
    ev1.record()
    barrier()
    if a.profile_region:
        torch.cuda.profiler.stop()
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
    units = a.nq * world
    value = units / (ms_step / 1e3)

    # ---- e2e: the reference-facing C API on an index opened from NGT's own files (written by rank 0), pinned host
    # queries in, pinned host results out; two host threads submit alternate batches
    if world > 1:
        bc = [index_dir]
        dist.broadcast_object_list(bc, 0)
        index_dir = bc[0]
    barrier()
    capi = bind_capi(_lib.SO_PATH)
    os.environ["NGTGPU_PIVOTS"] = str(a.pivots)
    err = capi.ngt_create_error_object()
    cix = capi.ngt_open_index(index_dir.encode(), err)
    e2e = {"value": None, "unit": "queries/s", "error": None}
    if not cix:
        e2e["error"] = capi.ngt_get_error_string(err).decode()
    else:
        n_threads = 2
        hq = [torch.empty((a.nq, dim), dtype=torch.float32).pin_memory() for _ in range(n_threads)]
        out = [(torch.zeros((a.nq, a.k), dtype=torch.int32).pin_memory(), torch.zeros((a.nq, a.k), dtype=torch.float32).pin_memory(),
                torch.zeros((a.nq,), dtype=torch.int32).pin_memory()) for _ in range(n_threads)]
        for i in range(n_threads):
            hq[i].copy_(batches[i])
        torch.cuda.synchronize(dev)
        errs = [capi.ngt_create_error_object() for _ in range(n_threads)]
        failed = []

        def call(t):
            ok = capi.ngt_batch_search_index_as_float(cix, hq[t].data_ptr(), a.nq, dim, a.k, eps, -1.0, a.edge_size if a.edge_size > 0 else 0,
                                                      out[t][0].data_ptr(), out[t][1].data_ptr(), out[t][2].data_ptr(), errs[t])
            if not ok:
                failed.append(capi.ngt_get_error_string(errs[t]).decode())

        def timed(nthreads, steps):
            def worker(t):
                for s in range(t, steps, nthreads):
                    call(t)
            ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
            t0 = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            torch.cuda.synchronize(dev)
            return (time.perf_counter() - t0) * 1e3 / steps

        timed(n_threads, max(a.warmup, 2) * n_threads)
        # the C API's answers are the engine's answers
        call(0)
        same = bool((out[0][0].numpy().astype(np.uint32)[:ngt] == ix.search(q_gt, a.k, eps, edge_size=a.edge_size, n_seeds=a.seeds)[0]).all())
        barrier()
        e2e_ms1 = timed(1, a.steps)
        barrier()
        e2e_ms = timed(n_threads, a.steps)
        barrier()
        if world > 1:
            t = torch.tensor([e2e_ms, e2e_ms1], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms, e2e_ms1 = float(t[0]), float(t[1])
        # the single-query call of lib/NGT/Capi.h (a batch of one): latency, for callers that cannot batch
        res1 = capi.ngt_create_empty_results(err)
        for i in range(20):
            capi.ngt_search_index_as_float(cix, hq[0].data_ptr() + (i % a.nq) * dim * 4, dim, a.k, eps, -1.0, res1, err)
        t1 = time.perf_counter()
        n1 = 200
        for i in range(n1):
            capi.ngt_search_index_as_float(cix, hq[0].data_ptr() + (i % a.nq) * dim * 4, dim, a.k, eps, -1.0, res1, err)
        single_us = (time.perf_counter() - t1) * 1e6 / n1
        capi.ngt_destroy_results(res1)
        e2e = {"value": round(units / (e2e_ms / 1e3), 1), "unit": "queries/s", "ms_per_step": round(e2e_ms, 4),
               "single_query_call": {"entry_point": "ngt_search_index_as_float", "mean_latency_us": round(single_us, 1)},
               "h2d_bytes_per_step": a.nq * dim * 4, "d2h_bytes_per_step": a.nq * a.k * 8 + a.nq * 4,
               "entry_point": "ngt_batch_search_index_as_float (lib/NGT/Capi.h batch twin) on ngt_open_index(<NGT index files>)",
               "host_threads": n_threads, "single_thread": {"value": round(units / (e2e_ms1 / 1e3), 1), "ms_per_step": round(e2e_ms1, 4)},
               "same_results_as_device_path": same, "errors": failed[:2]}
        capi.ngt_close_index(cix)

    # ---- BASELINE configs[2..4] under the same clock
    sharded_rec, workloads = (None, {})
    if a.extras != "none":
        sharded_rec, workloads = extra_records(a, dev, rank, world, lib)

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
                   "parallelism": "replica x%d (one 10k batch per GPU); the row-sharded path with its all-gather is the `sharded` record" % world,
                   "l2_policy": "inputs larger than L2 (%d MB of rows vs 126 MB), 4 rotating query batches" % (info["n"] * dim * 4 >> 20),
                   "graph": info.get("graph"), "setup_s": {k: info[k] for k in ("gen_s", "knn_graph_s", "reconstruct_s", "adjust_paths_s")},
                   "knn_pass": info["knn_pass"],
                   "overflow_queries_per_step": overflow, "epsilon_sweep": curve},
        "e2e": e2e,
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
    # dram bytes of the traversal launch from the committed `ncu --set full` capture: only valid for the kernel source it
    # was taken from, so it is dropped when search_fast.cuh has changed since
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and a.shape == "sift" and a.n == 1000000:
        try:
            import hashlib
            tj = json.load(open(tp))
            src = hashlib.sha256(open(os.path.join(ROOT, "ngt_b200", "csrc", "search_fast.cuh"), "rb").read()).hexdigest()[:16]
            if tj.get("kernel_source_sha256_16") == src:
                out["roofline"]["traffic"] = tj.get("search_kernel_dram_bytes_per_launch")
                out["roofline"]["traffic_source"] = tj.get("source")
            elif src in tj.get("sass_equivalent_sources", {}):
                # a later source whose measured instantiation compiles to the captured kernel's instructions (evidence named in the note)
                out["roofline"]["traffic"] = tj.get("search_kernel_dram_bytes_per_launch")
                out["roofline"]["traffic_source"] = tj.get("source")
                out["roofline"]["traffic_note"] = "capture of an earlier source: " + tj["sass_equivalent_sources"][src]
            else:
                out["roofline"]["traffic_note"] = "profiles/traffic.json was captured from another version of the kernel source: dropped"
        except Exception:
            pass
    if sharded_rec is not None:
        out["sharded"] = sharded_rec
    if workloads:
        out["workloads"] = workloads

    # ---- cpu_baseline: the reference itself on the same index files / seeds / epsilon, bounded sample
    if want_cpu:
        try:
            m = min(a.cpu_sample, a.nq)
            qs = batches[0][:m].cpu().numpy().copy()
            seeds = np.zeros((m, a.seeds), np.uint32)
            _lib.check(lib.ngtgpu_select_seeds(ix._h, qs.ctypes.data, _lib.OBJECT_FLOAT, m, a.seeds, seeds.ctypes.data))
            kind, isa = "reference", None
            try:
                R, h = reference_handle(index_dir)
                isa = R.isa
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
            out["cpu_baseline"] = {"value": round(qps, 1), "unit": "queries/s", "cores": int(threads), "kind": kind, "isa": isa,
                                   "sample": "%d queries of the batch, same index files, same seeds, epsilon %.2f, "
                                             "OpenMP over queries" % (m, eps),
                                   "recall_at_10": round(recall_at_k(rids, rd, rc.astype(np.int64), gt_i, gt_dd), 4),
                                   "gpu_results_identical": same}
            if a.recipe_n > 0 and kind == "reference":
                try:
                    out["cpu_baseline"]["reference_recipe"] = reference_recipe(a, dev, a.k)
                except Exception as ex:
                    out["cpu_baseline"]["reference_recipe"] = {"error": str(ex)[:300]}
        except Exception as ex:  # the baseline must never take the GPU number down with it
            out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": host_threads(), "kind": "reference",
                                   "sample": "failed: %s" % (str(ex)[:200])}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(a):
    """The reference arm: NGT::Index::search (unmodified reference, oracle/_ref) on the host cores. The index is the
    same one the GPU arm searches (built in the untimed setup and written in NGT's own file format); the timed
    region runs only reference code, on every host thread this process may use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from ngt_b200 import _lib
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib = _lib.load()
    lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    ix, info, index_dir = build_index(a, dev, True)
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
    kind, isa = "reference", None
    try:
        R, h = reference_handle(index_dir)
        isa = R.isa
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
        "cpu_baseline": {"value": round(qps, 1), "unit": "queries/s", "cores": int(threads), "kind": kind, "isa": isa,
                         "sample": "each step = %d queries of the 10k batch (NGT::Index::search, OpenMP over queries on %d threads = "
                                   "the process's CPU affinity, explicit seeds identical to the GPU arm's)" % (m, threads)},
        "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

"""Index construction on the device.

The distance work -- the brute-force kNN pass that GraphIndex::searchForKNNGInsertion
(lib/NGT/Index.h:839-856) does one object at a time -- runs in libngtgpu.so (ngtgpu_index_knn_graph).
What is left is adjacency-list surgery with no arithmetic in it (concatenate, sort by (node, distance, id),
drop duplicates); it is done with torch sort/unique primitives on the device, following
GraphReconstructor::reconstructGraph (lib/NGT/GraphReconstructor.h:425-561): keep the first `outgoing`
edges of every node, add the reverse of the first `incoming` edges, sort, dedupe.
"""
import ctypes as C

import numpy as np

from . import _lib

_ready = False


def _fn():
    global _ready
    lib = _lib.load()
    if not _ready:
        P = C.c_void_p
        lib.ngtgpu_index_knn_graph.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, P, P, P, P]
        lib.ngtgpu_index_knn_graph.restype = C.c_int
        _ready = True
    return lib


def knn_graph(ix, k, batch=148 * 256 * 3):
    """Exact k nearest OTHER objects of every stored object. -> (ids [n,k] int32, dists [n,k] float32,
    counts [n] int32) as torch CUDA tensors; lists ascending by (distance,id)."""
    import torch
    lib = _fn()
    n = ix.size
    dev = torch.device("cuda", ix.device)
    ids = torch.zeros((n, k), dtype=torch.int32, device=dev)
    dists = torch.zeros((n, k), dtype=torch.float32, device=dev)
    counts = torch.zeros((n,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for s in range(0, n, batch):
        m = min(batch, n - s)
        _lib.check(lib.ngtgpu_index_knn_graph(ix._h, int(k), s + 1, m, ids[s:].data_ptr(), dists[s:].data_ptr(),
                                              counts[s:].data_ptr(), stream))
    return ids, dists, counts


def _dist_sort_key(dists):
    """float32 -> int64 that sorts like the float (handles negative zero and tiny negative cosines)."""
    import torch
    b = (dists + 0.0).view(torch.int32).to(torch.int64)
    return torch.where(b < 0, -(b & 0x7fffffff) - 1, b)


def reconstruct_graph(ids, dists, counts, outgoing, incoming):
    """GraphReconstructor::reconstructGraph on a [n,k] neighbour table (1-based ids, ascending lists).
    -> (row_ptr [n+2] int64 over ids 0..n, col int32, dist float32), lists ascending by (distance,id)."""
    import torch
    n, k = ids.shape
    dev = ids.device
    outgoing, incoming = min(outgoing, k), min(incoming, k)
    node = torch.arange(1, n + 1, device=dev, dtype=torch.int64)[:, None]
    rank = torch.arange(k, device=dev)[None, :]
    have = rank < counts[:, None]
    # outgoing part: node -> its first `outgoing` neighbours
    mo = have & (rank < outgoing)
    src_o = node.expand(n, k)[mo]
    dst_o = ids.to(torch.int64)[mo]
    d_o = dists[mo]
    # incoming part: the reverse of the first `incoming` edges
    mi = have & (rank < incoming)
    src_i = ids.to(torch.int64)[mi]
    dst_i = node.expand(n, k)[mi]
    d_i = dists[mi]
    src = torch.cat([src_o, src_i])
    dst = torch.cat([dst_o, dst_i])
    d = torch.cat([d_o, d_i])
    del src_o, src_i, dst_o, dst_i, d_o, d_i, mo, mi, have
    # sort by (src, distance, dst): stable sort by src after a sort by (distance, dst)
    key = (_dist_sort_key(d) << 32) | dst
    order = torch.argsort(key, stable=True)
    src, dst, d = src[order], dst[order], d[order]
    order = torch.argsort(src, stable=True)
    src, dst, d = src[order], dst[order], d[order]
    del key, order
    keep = torch.ones_like(src, dtype=torch.bool)
    keep[1:] = (src[1:] != src[:-1]) | (dst[1:] != dst[:-1])
    src, dst, d = src[keep], dst[keep], d[keep]
    deg = torch.bincount(src, minlength=n + 1)
    row_ptr = torch.zeros(n + 2, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(deg, 0)
    return row_ptr, dst.to(torch.int32), d


def reconstruct_graph_csr(row_ptr, col, dist, outgoing, incoming):
    """GraphReconstructor::reconstructGraph (lib/NGT/GraphReconstructor.h:425-561) on adjacency lists of any
    length (an ANNG as `grp` stores it): node i keeps its first `outgoing` edges -- or ALL of them when it has fewer
    than `outgoing` (the reference leaves such nodes untouched, :447-456) -- then gets the reverse of the first
    `incoming` edges of every node; lists are sorted by (distance, id) and de-duplicated.
    row_ptr: [n+2] over ids 0..n, col: 1-based ids, dist: float32. torch tensors (any device)."""
    import torch
    row_ptr = row_ptr.to(torch.int64)
    dev = col.device
    n = row_ptr.numel() - 2
    deg = row_ptr[1:] - row_ptr[:-1]                      # per id 0..n
    src = torch.repeat_interleave(torch.arange(n + 1, device=dev, dtype=torch.int64), deg)
    rank = torch.arange(col.numel(), device=dev, dtype=torch.int64) - row_ptr[src]
    dst = col.to(torch.int64)
    d = dist.to(torch.float32)
    keep_all = deg[src] < outgoing
    mo = (rank < outgoing) | keep_all if outgoing > 0 else torch.zeros_like(rank, dtype=torch.bool)
    mi = rank < incoming
    s2 = torch.cat([src[mo], dst[mi]])
    t2 = torch.cat([dst[mo], src[mi]])
    d2 = torch.cat([d[mo], d[mi]])
    key = (_dist_sort_key(d2) << 32) | t2
    order = torch.argsort(key, stable=True)
    s2, t2, d2 = s2[order], t2[order], d2[order]
    order = torch.argsort(s2, stable=True)
    s2, t2, d2 = s2[order], t2[order], d2[order]
    # the reference drops an entry when its id equals the previous entry's id in the sorted list (:519-533)
    keep = torch.ones_like(s2, dtype=torch.bool)
    keep[1:] = (s2[1:] != s2[:-1]) | (t2[1:] != t2[:-1])
    s2, t2, d2 = s2[keep], t2[keep], d2[keep]
    out_deg = torch.bincount(s2, minlength=n + 1)
    out_ptr = torch.zeros(n + 2, dtype=torch.int64, device=dev)
    out_ptr[1:] = torch.cumsum(out_deg, 0)
    return out_ptr, t2.to(torch.int32), d2


def reconstruct_graph_device(row_ptr, col, dist, outgoing, incoming):
    """reconstruct_graph_csr done by libngtgpu.so (ngtgpu_graph_reconstruct: emit, two radix sorts, compact) -- what
    the C API's optimizer entry points run. CUDA tensors in, CUDA tensors out."""
    import torch
    lib = _lib.load()
    dev = col.device
    if dev.type != "cuda":
        raise _lib.NgtGpuError(_lib.ERR_NO_DEVICE, "reconstruct_graph_device: the graph must be on a CUDA device")
    n = row_ptr.numel() - 2
    rp = row_ptr.to(torch.int64).contiguous()
    c = col.to(torch.int32).contiguous()
    d = dist.to(torch.float32).contiguous()
    cap = 2 * max(c.numel(), 1)
    out_ptr = torch.zeros(n + 2, dtype=torch.int64, device=dev)
    out_col = torch.empty(cap, dtype=torch.int32, device=dev)
    out_dist = torch.empty(cap, dtype=torch.float32, device=dev)
    nnz = C.c_uint64(0)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.ngtgpu_graph_reconstruct(n, rp.data_ptr(), c.data_ptr(), d.data_ptr(), int(outgoing), int(incoming),
                                                cap, out_ptr.data_ptr(), out_col.data_ptr(), out_dist.data_ptr(),
                                                C.byref(nnz), stream))
    return out_ptr, out_col[:nnz.value].clone(), out_dist[:nnz.value].clone()


def refine_anng(ix, row_ptr, col, dist, epsilon=0.1, no_of_edges=0, edge_size=-1, batch_size=10000,
                edge_size_for_creation=10, n_seeds=10):
    """GraphReconstructor::refineANNG (lib/NGT/GraphReconstructor.h:814-924) on the device
    (ngtgpu_index_refine_anng): batched self-search of every object in the current graph, results merged into the
    out-edges, reverse edges added (unless a kNN graph is asked for with no_of_edges != 0). The graph (CUDA CSR with
    distances) is returned refined and is also left set on `ix`."""
    import torch
    lib = _lib.load()
    dev = col.device
    if dev.type != "cuda":
        raise _lib.NgtGpuError(_lib.ERR_NO_DEVICE, "refine_anng: the graph must be on a CUDA device (no CPU path)")
    n = ix.size
    # noOfSearchedEdges, GraphReconstructor.h:825
    k = -no_of_edges if no_of_edges < 0 else max(no_of_edges, edge_size_for_creation)
    nnz = col.numel()
    cap = nnz + 2 * n * k
    rp = row_ptr.to(torch.int64).clone()
    c = torch.zeros(cap, dtype=torch.int32, device=dev)
    d = torch.zeros(cap, dtype=torch.float32, device=dev)
    c[:nnz] = col
    d[:nnz] = dist
    out = C.c_uint64(0)
    torch.cuda.synchronize(dev)
    _lib.check(lib.ngtgpu_index_refine_anng(ix._h, float(epsilon), int(no_of_edges), int(edge_size), int(k), int(batch_size),
                                            int(n_seeds), cap, rp.data_ptr(), c.data_ptr(), d.data_ptr(), C.byref(out)))
    return rp, c[:out.value].clone(), d[:out.value].clone()


def insert_objects(ix, first_id, count, graph=None, edge_size_for_creation=10, epsilon=0.1, edge_size=-1, batch_size=200,
                   n_seeds=10, n_pivots=256, pivot_seed=1):
    """The reference's ANNG construction loop (lib/NGT/Index.cpp:721-792 createIndex: batches of batchSizeForCreation
    objects are searched for on the frozen graph, linked among themselves and inserted with their reverse edges) on the
    device, batch by batch (ngtgpu_index_insert_batch). `graph` = (row_ptr, col, dist) CUDA CSR of the objects already
    in the index (None: start from nothing); the objects first_id .. first_id+count-1 must be stored in `ix`.
    -> (row_ptr, col, dist); the graph is also left set on `ix`. With first_id = 1, count = ix.size this is
    NGT::Index::createIndex; with a graph and the appended ids it is insertion into an existing index."""
    import torch
    lib = _lib.load()
    n = ix.size
    dev = torch.device("cuda", ix.device)
    e = int(edge_size_for_creation)
    old = 0 if graph is None else int(graph[1].numel())
    cap = old + 2 * count * e + 16
    rp = torch.zeros(n + 2, dtype=torch.int64, device=dev)
    c = torch.zeros(cap, dtype=torch.int32, device=dev)
    d = torch.zeros(cap, dtype=torch.float32, device=dev)
    if graph is not None:
        rp.copy_(graph[0].to(torch.int64))
        c[:old] = graph[1]
        d[:old] = graph[2]
    nnz = C.c_uint64(old)
    torch.cuda.synchronize(dev)
    for s in range(first_id, first_id + count, batch_size):
        m = min(batch_size, first_id + count - s)
        _lib.check(lib.ngtgpu_index_insert_batch(ix._h, s, m, e, float(epsilon), int(edge_size), int(n_seeds), int(n_pivots),
                                                 int(pivot_seed), cap, rp.data_ptr(), c.data_ptr(), d.data_ptr(), C.byref(nnz)))
    return rp, c[:nnz.value].clone(), d[:nnz.value].clone()


def adjust_paths(row_ptr, col, dist, min_edges=0, with_stats=False):
    """GraphReconstructor::adjustPathsEffectively (lib/NGT/GraphReconstructor.h:197-386) -- the shortcut reduction
    GraphOptimizer::execute applies after reconstructGraph -- on a device CSR (row_ptr over ids 0..n, lists
    ascending by (distance, id)). The decisions are made by libngtgpu.so (ngtgpu_graph_adjust_paths); compaction of
    the surviving edges is a masked select. -> (row_ptr, col, dist) of the adjusted graph."""
    import torch
    lib = _lib.load()
    dev = col.device
    if dev.type != "cuda":
        raise _lib.NgtGpuError(_lib.ERR_NO_DEVICE, "adjust_paths: the graph must be on a CUDA device (no CPU path)")
    n = row_ptr.numel() - 2
    rp = row_ptr.to(torch.int64).contiguous()
    c = col.to(torch.int32).contiguous()
    d = dist.to(torch.float32).contiguous()
    keep = torch.zeros(max(c.numel(), 1), dtype=torch.uint8, device=dev)
    stats = (C.c_uint64 * 4)()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.ngtgpu_graph_adjust_paths(n, rp.data_ptr(), c.data_ptr(), d.data_ptr(), int(min_edges),
                                                 keep.data_ptr(), stats, stream))
    keep = keep[:c.numel()].bool()
    src = torch.repeat_interleave(torch.arange(n + 1, device=dev, dtype=torch.int64), rp[1:] - rp[:-1])
    deg = torch.bincount(src[keep], minlength=n + 1)
    out_ptr = torch.zeros(n + 2, dtype=torch.int64, device=dev)
    out_ptr[1:] = torch.cumsum(deg, 0)
    out = (out_ptr, c[keep], d[keep])
    if with_stats:
        return out + ({"candidates": int(stats[0]), "removed": int(stats[1]), "sweeps": int(stats[2]),
                       "launches": int(stats[3])},)
    return out


def graph_statistics(row_ptr):
    deg = (row_ptr[2:] - row_ptr[1:-1]).float()
    return {"edges": int(row_ptr[-1]), "mean_degree": float(deg.mean()), "min_degree": int(deg.min()),
            "max_degree": int(deg.max())}


def csr_from_table(ids, counts):
    """[n,k] table -> CSR numpy (row_ptr over ids 0..n)."""
    ids = np.asarray(ids)
    counts = np.asarray(counts).astype(np.int64)
    n, k = ids.shape
    row_ptr = np.zeros(n + 2, np.uint64)
    row_ptr[2:] = np.cumsum(counts)
    mask = np.arange(k)[None, :] < counts[:, None]
    return row_ptr, ids[mask].astype(np.uint32)

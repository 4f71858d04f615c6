"""ctypes bindings for the parity oracle (TEST INFRASTRUCTURE ONLY).

Two libraries, both built by oracle/Makefile:
  * libngt_oracle.so  -- the C restatement (oracle/ngt_oracle.c), kind "port"
  * _ref/libngt_ref.so -- the unmodified reference compiled from /root/reference plus
                          oracle/ref_shim.cpp, kind "reference"

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. The product package ngt_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libngt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libngt_ref.so")          # -march=x86-64-v3: the reference's AVX2 path (runs anywhere)
REF_SO_V4 = os.path.join(HERE, "_ref", "v4", "libngt_ref.so")  # -march=x86-64-v4: its AVX-512 path


def host_has_avx512():
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
    except (OSError, StopIteration):
        return False
    return all(f in flags for f in ("avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"))

# ObjectSpace.h:166-186
L1, L2, HAMMING, ANGLE, COSINE, NORMALIZED_ANGLE, NORMALIZED_COSINE, JACCARD = range(8)
NORMALIZED_L2 = 9
UINT8, FLOAT = 1, 2

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")


def build(ref=True):
    """Compile the restatement and, when /root/reference is present, the reference itself."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def padded_dimension(dim):
    return ((dim - 1) // 16 + 1) * 16


def pad_objects(x, otype):
    """[n, dim] -> [(n+1), padded] with the dummy row 0 (1-based ids), zero padded (ObjectSpace.h:357-400)."""
    x = np.asarray(x)
    n, dim = x.shape
    dt = np.uint8 if otype == UINT8 else np.float32
    out = np.zeros((n + 1, padded_dimension(dim)), dtype=dt)
    out[1:, :dim] = x.astype(dt)
    return out


def pad_queries(q, otype):
    q = np.asarray(q)
    dt = np.uint8 if otype == UINT8 else np.float32
    out = np.zeros((q.shape[0], padded_dimension(q.shape[1])), dtype=dt)
    out[:, : q.shape[1]] = q.astype(dt)
    return out


class Port:
    """The C restatement."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = lib = C.CDLL(PORT_SO)
        lib.ngto_distance.restype = C.c_double
        lib.ngto_distance.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.ngto_set_simd.argtypes = [C.c_int, C.c_int]
        lib.ngto_normalize.argtypes = [_f32p, C.c_size_t]
        lib.ngto_edge_size.restype = C.c_int64
        lib.ngto_edge_size.argtypes = [C.c_int64, C.c_int64, C.c_float, C.c_int64, C.c_int64]
        lib.ngto_recall.restype = C.c_double
        lib.ngto_recall.argtypes = [_u32p, _f32p, C.c_size_t, _u32p, _f32p, C.c_size_t]
        lib.ngto_batch_linear_search.argtypes = [
            C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_size_t, C.c_double, C.c_size_t, _u32p, _f32p, _u32p]
        lib.ngto_batch_graph_search.argtypes = [
            C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p, _u32p, C.c_void_p,
            C.c_size_t, C.c_size_t, _u32p, C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_int64,
            _u32p, _f32p, _u32p, C.c_void_p]

    def set_simd(self, lanes, fma):
        self.lib.ngto_set_simd(lanes, int(fma))

    def distance(self, dtype, otype, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        return self.lib.ngto_distance(dtype, otype, a.ctypes.data, b.ctypes.data, a.shape[-1])

    def normalize(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32).copy()
        for i in range(x.shape[0]):
            row = x[i]
            if self.lib.ngto_normalize(row, row.shape[0]) != 0:
                raise ValueError("zero vector")
        return x

    def edge_size(self, sc_edge_size, prop_edge_size_for_search, exploration_coefficient, dyn_base, dyn_rate):
        return self.lib.ngto_edge_size(sc_edge_size, prop_edge_size_for_search, exploration_coefficient,
                                       dyn_base, dyn_rate)

    def recall(self, ids, dists, gt_ids, gt_dists):
        ids = np.ascontiguousarray(ids, np.uint32)
        dists = np.ascontiguousarray(dists, np.float32)
        gt_ids = np.ascontiguousarray(gt_ids, np.uint32)
        gt_dists = np.ascontiguousarray(gt_dists, np.float32)
        return self.lib.ngto_recall(ids, dists, ids.size, gt_ids, gt_dists, gt_ids.size)

    def mean_recall(self, ids, dists, counts, gt_ids, gt_dists):
        tot = 0.0
        for q in range(ids.shape[0]):
            c = int(counts[q])
            tot += self.recall(ids[q, :c], dists[q, :c], gt_ids[q], gt_dists[q])
        return tot / ids.shape[0]

    def linear_search(self, dtype, otype, objects, queries, k, radius=-1.0, valid=None):
        """objects: padded [(n+1), pad]; queries: padded [nq, pad]."""
        objects = np.ascontiguousarray(objects)
        queries = np.ascontiguousarray(queries)
        n = objects.shape[0] - 1
        nq = queries.shape[0]
        ids = np.zeros((nq, k), np.uint32)
        dists = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        vp = None
        if valid is not None:
            valid = np.ascontiguousarray(valid, np.uint8)
            vp = valid.ctypes.data
        self.lib.ngto_batch_linear_search(dtype, otype, objects.ctypes.data, objects.strides[0], n, vp,
                                          objects.shape[1], queries.ctypes.data, queries.strides[0], nq,
                                          float(radius), k, ids, dists, counts)
        return ids, dists, counts

    def graph_search(self, dtype, otype, objects, row_ptr, col, queries, seeds, k, epsilon, radius=-1.0,
                     edge_size=2 ** 31 - 1):
        objects = np.ascontiguousarray(objects)
        queries = np.ascontiguousarray(queries)
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        seeds = np.ascontiguousarray(seeds, np.uint32)
        n = objects.shape[0] - 1
        nq = queries.shape[0]
        ids = np.zeros((nq, k), np.uint32)
        dists = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        stats = np.zeros((nq, 3), np.uint64)
        self.lib.ngto_batch_graph_search(dtype, otype, objects.ctypes.data, objects.strides[0], n,
                                         objects.shape[1], row_ptr, col, queries.ctypes.data,
                                         queries.strides[0], nq, seeds, seeds.shape[1], k, epsilon, radius,
                                         edge_size, ids, dists, counts, stats.ctypes.data)
        return ids, dists, counts, stats


class Ref:
    """The unmodified reference behind oracle/ref_shim.cpp."""

    def __init__(self, isa="avx2"):
        """isa: "avx2" (default: the build every golden vector was made with), or "native": the AVX-512 build when the
        host has AVX-512 (what the reference's own -march=native build would use there), else AVX2."""
        so, self.isa = REF_SO, "avx2 (-march=x86-64-v3)"
        if isa == "native" and os.path.exists(REF_SO_V4) and host_has_avx512():
            so, self.isa = REF_SO_V4, "avx512 (-march=x86-64-v4)"
        if not os.path.exists(so):
            raise FileNotFoundError(so + " (build it with `make -C oracle ref` where /root/reference exists)")
        self.lib = lib = C.CDLL(so)
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_build_index.argtypes = [C.c_char_p, _f32p, C.c_size_t, C.c_int, C.c_char, C.c_int, C.c_int,
                                        C.c_int, C.c_char, C.c_int]
        lib.ref_build_onng.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        lib.ref_open.restype = C.c_void_p
        lib.ref_open.argtypes = [C.c_char_p, C.c_int]
        lib.ref_close.argtypes = [C.c_void_p]
        lib.ref_info.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.int64, flags="C")]
        lib.ref_get_object.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.ref_export_graph.restype = C.c_int64
        lib.ref_export_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_int, C.c_size_t, C.c_float, C.c_float,
                                   C.c_int, C.c_void_p, C.c_size_t, _u32p, _f32p, _u32p, C.c_void_p, C.c_int,
                                   C.POINTER(C.c_double)]
        lib.ref_linear_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_int, C.c_size_t, C.c_float, _u32p,
                                          _f32p, _u32p, C.c_int, C.POINTER(C.c_double)]
        lib.ref_max_threads.restype = C.c_int
        lib.ref_build_anng_fixed_seeds.argtypes = [C.c_char_p, _f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_char, C.c_int,
                                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ref_refine_anng.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_size_t]
        lib.ref_object_distance.restype = C.c_double
        lib.ref_object_distance.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        lib.ref_insert_node.argtypes = [C.c_void_p, C.c_uint32, _u32p, _f32p, C.c_size_t]
        lib.ref_remove.argtypes = [C.c_void_p, C.c_uint32]
        lib.ref_build_onng_with_accuracy_table.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ref_epsilon_from_accuracy_table.restype = C.c_float
        lib.ref_epsilon_from_accuracy_table.argtypes = [C.c_char_p, C.c_double]
        lib.ref_tree_seeds.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_int, C.c_size_t, _u32p, C.c_size_t, _u32p]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def max_threads(self):
        return self.lib.ref_max_threads()

    def build_index(self, path, data, objtype="f", disttype=L2, edge_creation=10, edge_search=40,
                    indextype="t", threads=8):
        data = np.ascontiguousarray(data, np.float32)
        self._check(self.lib.ref_build_index(path.encode(), data, data.shape[0], data.shape[1],
                                             objtype.encode(), disttype, edge_creation, edge_search,
                                             indextype.encode(), threads))

    def build_anng_fixed_seeds(self, path, data, n_first=None, objtype="f", disttype=L2, edge_creation=10,
                               edge_search=40, seed_size=10, batch_size=200, threads=4):
        """createIndex's batched loop on a graph-only index with SeedType FixedNodes (deterministic); rows after
        n_first are inserted into the finished graph by a second createIndex."""
        data = np.ascontiguousarray(data, np.float32)
        n = data.shape[0]
        self._check(self.lib.ref_build_anng_fixed_seeds(path.encode(), data, n, n if n_first is None else n_first,
                                                        data.shape[1], objtype.encode(), disttype, edge_creation,
                                                        edge_search, seed_size, batch_size, threads))

    def refine_anng(self, h, epsilon=0.1, accuracy=0.0, no_of_edges=0, explore_edge_size=-2 ** 31, batch_size=10000):
        self._check(self.lib.ref_refine_anng(h, epsilon, accuracy, no_of_edges, explore_edge_size, batch_size))

    def build_onng_with_accuracy_table(self, anng, onng, outgoing, incoming, shortcut=True, n_queries=100, n_results=20):
        self._check(self.lib.ref_build_onng_with_accuracy_table(anng.encode(), onng.encode(), outgoing, incoming, int(shortcut),
                                                                n_queries, n_results))

    def remove(self, h, object_id):
        self._check(self.lib.ref_remove(h, int(object_id)))

    def epsilon_from_accuracy_table(self, table, accuracy):
        e = self.lib.ref_epsilon_from_accuracy_table(table.encode(), accuracy)
        if e != e:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return e

    def object_distance(self, h, a, b):
        return self.lib.ref_object_distance(h, a, b)

    def insert_node(self, h, node_id, ids, dists):
        ids = np.ascontiguousarray(ids, np.uint32)
        dists = np.ascontiguousarray(dists, np.float32)
        self._check(self.lib.ref_insert_node(h, node_id, ids, dists, ids.shape[0]))

    def build_onng(self, anng, onng, outgoing=10, incoming=120, shortcut=True):
        self._check(self.lib.ref_build_onng(anng.encode(), onng.encode(), outgoing, incoming, int(shortcut)))

    def open(self, path, readonly=False):
        h = self.lib.ref_open(path.encode(), int(readonly))
        if not h:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return h

    def close(self, h):
        self.lib.ref_close(h)

    def info(self, h):
        a = np.zeros(10, np.int64)
        self._check(self.lib.ref_info(h, a))
        keys = ["repo_size", "dim", "padded", "object_type", "distance_type", "edge_size_for_search",
                "dyn_base", "dyn_rate", "seed_size", "byte_size"]
        return dict(zip(keys, [int(v) for v in a]))

    def objects(self, h):
        """Stored object bytes as [n, dim] array of the object type (unpadded; normalised if applicable)."""
        inf = self.info(h)
        n = inf["repo_size"] - 1
        dt = np.uint8 if inf["object_type"] == UINT8 else np.float32
        out = np.zeros((n, inf["byte_size"] // np.dtype(dt).itemsize), dt)
        for i in range(n):
            self._check(self.lib.ref_get_object(h, i + 1, out[i].ctypes.data))
        return out

    def graph(self, h):
        nnz = self.lib.ref_export_graph(h, None, None, None)
        if nnz < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        rs = self.info(h)["repo_size"]
        row_ptr = np.zeros(rs + 1, np.uint64)
        col = np.zeros(max(nnz, 1), np.uint32)
        dist = np.zeros(max(nnz, 1), np.float32)
        self.lib.ref_export_graph(h, row_ptr.ctypes.data, col.ctypes.data, dist.ctypes.data)
        return row_ptr, col[:nnz], dist[:nnz]

    def search(self, h, queries, k, epsilon=0.1, radius=-1.0, edge_size=-1, seeds=None, threads=1, stats=True):
        queries = np.ascontiguousarray(queries, np.float32)
        nq, dim = queries.shape
        ids = np.zeros((nq, k), np.uint32)
        dists = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        st = np.zeros((nq, 2), np.uint64)
        sec = C.c_double(0)
        sp, ns = None, 0
        if seeds is not None:
            seeds = np.ascontiguousarray(seeds, np.uint32)
            sp, ns = seeds.ctypes.data, seeds.shape[1]
        self._check(self.lib.ref_search(h, queries, nq, dim, k, epsilon, radius, edge_size, sp, ns, ids, dists,
                                        counts, st.ctypes.data if stats else None, threads, C.byref(sec)))
        return ids, dists, counts, st, sec.value

    def tree_seeds(self, h, queries, k, max_seeds=128):
        queries = np.ascontiguousarray(queries, np.float32)
        nq, dim = queries.shape
        seeds = np.zeros((nq, max_seeds), np.uint32)
        ns = np.zeros(nq, np.uint32)
        self._check(self.lib.ref_tree_seeds(h, queries, nq, dim, k, seeds, max_seeds, ns))
        return seeds, ns

    def linear_search(self, h, queries, k, radius=-1.0, threads=1):
        queries = np.ascontiguousarray(queries, np.float32)
        nq, dim = queries.shape
        ids = np.zeros((nq, k), np.uint32)
        dists = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        sec = C.c_double(0)
        self._check(self.lib.ref_linear_search(h, queries, nq, dim, k, radius, ids, dists, counts, threads,
                                               C.byref(sec)))
        return ids, dists, counts, sec.value


def adjust_paths_loop(row_ptr, col, dist, min_edges=0):
    """Sequential restatement of GraphReconstructor::adjustPathsEffectively
    (lib/NGT/GraphReconstructor.h:197-386) for small graphs (pure-Python loops; test infrastructure).
    row_ptr over ids 0..n, lists ascending by (distance, id). -> list of kept (id, distance) lists per id 0..n."""
    n = len(row_ptr) - 2
    tmp = [[(int(col[e]), float(dist[e])) for e in range(int(row_ptr[i]), int(row_ptr[i + 1]))] for i in range(n + 1)]
    remove_candidates = [[] for _ in range(n + 1)]
    for src in range(1, n + 1):                                     # :236-284
        node = tmp[src]
        neighbors = {nid: (sni, d) for sni, (nid, d) in enumerate(node)}
        cands = []
        for sni, (path, d1) in enumerate(node):
            for (dst, d2) in tmp[path]:
                hit = neighbors.get(dst)
                if hit is not None and d1 < hit[1] and d2 < hit[1]:
                    cands.append((hit[0], (path, dst)))
        cands.sort(reverse=True)
        remove_candidates[src] = [c[1] for c in cands]
    out = [dict() for _ in range(n + 1)]                            # id -> distance, the graph being rebuilt
    ids = list(range(1, n + 1))
    rank = 0
    while ids:                                                      # :299-371
        nxt = []
        for src in ids:
            node = tmp[src]
            if rank >= len(node):
                continue
            rc = remove_candidates[src]
            if rc and (len(out[src]) + len(node) - rank) > min_edges:
                path_exist = False
                while rc and rc[-1][1] == node[rank][0]:
                    path, dst = rc.pop()
                    if path in out[src] and dst in out[path]:
                        path_exist = True
                        while rc and rc[-1][1] == node[rank][0]:
                            rc.pop()
                        break
                if path_exist:
                    nxt.append(src)
                    continue
            out[src][node[rank][0]] = node[rank][1]
            nxt.append(src)
        ids = nxt
        rank += 1
    return [sorted(((d, i) for i, d in o.items())) for o in out]   # :373-383: sorted by (distance, id)


def refine_anng_loop(port, dtype, otype, pobj, row_ptr, col, dist, seeds, epsilon=0.1, no_of_edges=0,
                     edge_size=2 ** 31 - 1, batch_size=10000, edge_size_for_creation=10):
    """Sequential restatement of GraphReconstructor::refineANNG (lib/NGT/GraphReconstructor.h:814-924) over the C
    restatement of the search (Port.graph_search) with explicit seeds per object (seeds[id - 1]); test infrastructure.
    pobj: padded objects [(n+1) x padDim]. -> per id 0..n a list of (distance, id) ascending."""
    n = pobj.shape[0] - 1
    lists = [[(float(dist[e]), int(col[e])) for e in range(int(row_ptr[i]), int(row_ptr[i + 1]))] for i in range(n + 1)]
    k = -no_of_edges if no_of_edges < 0 else max(no_of_edges, edge_size_for_creation)        # :825
    for bid in range(1, n + 1, batch_size):
        ids_b = list(range(bid, min(bid + batch_size, n + 1)))
        rp = np.zeros(n + 2, np.uint64)
        rp[1:] = np.cumsum([len(l) for l in lists])
        cc = np.array([t for l in lists for (_, t) in l], np.uint32)
        r_ids, r_d, r_cnt, _ = port.graph_search(dtype, otype, pobj, rp, cc, pobj[bid:bid + len(ids_b)],
                                                 seeds[bid - 1:bid - 1 + len(ids_b)], k, epsilon, edge_size=edge_size)
        for x, nid in enumerate(ids_b):                                                          # :869-888
            node = lists[nid] + [(float(r_d[x, r]), int(r_ids[x, r])) for r in range(int(r_cnt[x])) if int(r_ids[x, r]) != nid]
            node.sort()
            out, prev = [], 0
            for e in node:
                if e[1] == prev:
                    continue
                prev = e[1]
                out.append(e)
            lists[nid] = out
        if no_of_edges != 0:
            continue
        for x, nid in enumerate(ids_b):                                                          # :893-901, Graph.h:845-875
            for r in range(int(r_cnt[x])):
                t, d = int(r_ids[x, r]), float(r_d[x, r])
                if t == nid:
                    continue
                node = lists[t]
                import bisect
                pos = bisect.bisect_left(node, (d, nid))
                if pos < len(node) and node[pos][1] == nid:
                    continue
                node.insert(pos, (d, nid))
    if no_of_edges > 0:
        lists = [l[:no_of_edges] for l in lists]
    return lists


def fixed_node_seeds(n, seed_size):
    """SeedTypeFixedNodes (lib/NGT/Index.h:1122-1127): every search starts from ids 1..seed_size. [n x seed_size]."""
    return np.tile(np.arange(1, seed_size + 1, dtype=np.uint32), (n, 1))


def build_anng_loop(port, pobj, rows_int, seeds, first_id, count, lists=None, edge_size_for_creation=10, epsilon=0.1,
                    edge_size=40, batch_size=200, dtype=L2, otype=FLOAT):
    """Sequential restatement of the ANNG construction loop (lib/NGT/Index.cpp:631-719, 721-792; Index.h:815-837;
    Graph.h:611-626, 845-886) over the C restatement of the search with explicit seeds (seeds[id - 1]); test
    infrastructure. Pinned to the reference's createIndex by tests/golden/anng_build.npz (test_oracle_pin.py).
    rows_int: [(n+1) x dim] int64 copy of the objects (row 0 dummy) for integer-valued L2 data, or None: the in-batch
    distances then come from the C restatement of the comparator. -> per id 0..n a list of (distance, id) ascending."""
    import bisect
    n = pobj.shape[0] - 1
    e = edge_size_for_creation
    lists = [[] for _ in range(n + 1)] if lists is None else [list(l) for l in lists]
    for s in range(first_id, first_id + count, batch_size):
        ids_b = list(range(s, min(s + batch_size, first_id + count)))
        res = [[] for _ in ids_b]
        nnz = sum(len(l) for l in lists)
        if nnz and s > 1:                                                                   # searchForNNGInsertion
            rp = np.zeros(n + 2, np.uint64)
            rp[1:] = np.cumsum([len(l) for l in lists])
            cc = np.array([t for l in lists for (_, t) in l], np.uint32)
            q = pobj[s:s + len(ids_b)]
            sd = seeds[s - 1:s - 1 + len(ids_b)]
            cap = 2 ** 31 - 1 if edge_size == 0 else edge_size
            r_ids, r_d, r_cnt, _ = port.graph_search(dtype, otype, pobj, rp, cc, q, sd, e, epsilon, edge_size=cap)
            short = [x for x in range(len(ids_b)) if r_cnt[x] < e and r_cnt[x] < s]   # result.size() < repository.size()
            if short and edge_size != 0:                                                    # Index.h:826-836
                a_ids, a_d, a_cnt, _ = port.graph_search(dtype, otype, pobj, rp, cc, q, sd, e, epsilon, edge_size=2 ** 31 - 1)
                for x in short:
                    r_ids[x], r_d[x], r_cnt[x] = a_ids[x], a_d[x], a_cnt[x]
            for x in range(len(ids_b)):
                res[x] = [(float(r_d[x, r]), int(r_ids[x, r])) for r in range(int(r_cnt[x]))]
        for x, i in enumerate(ids_b):                                                       # insertMultipleSearchResults
            objs = list(res[x])
            for j in ids_b[:x]:
                if rows_int is not None:
                    d2 = int(((rows_int[i] - rows_int[j]) ** 2).sum())
                    objs.append((float(np.float32(np.sqrt(np.float64(d2)))), j))
                else:
                    objs.append((float(np.float32(port.distance(dtype, otype, pobj[i], pobj[j]))), j))
            objs.sort()
            res[x] = objs[:e]
        for x, i in enumerate(ids_b):                                                       # insertANNGNode
            lists[i] = list(res[x])
            for (d, t) in res[x]:
                bisect.insort(lists[t], (d, i))
    return lists


def remove_edges_reliably_loop(port, dtype, otype, pobj, lists, node_id):
    """Sequential restatement of NeighborhoodGraph::removeEdgesReliably (lib/NGT/Graph.cpp:641-864), what
    NGT::Index::remove does to the graph: the removed node's back edges go, then its neighbours are chained -- neighbour
    i is linked (both ways, at their own distance) to the nearest of the neighbours after it, which then takes place
    i + 1 -- so that the removal does not disconnect them. `lists`: per id a list of (distance, id) ascending, edited in
    place. Pinned to the reference by tests/golden/remove.npz (test_oracle_pin.py); test infrastructure."""
    import bisect
    node = list(lists[node_id])
    if not node:
        return
    for (d, nid) in node:
        n = lists[nid]
        pos = bisect.bisect_left(n, (d, node_id))
        if pos < len(n) and n[pos][1] == node_id:     # (else: reported and skipped, NGT_FORCED_REMOVE, defines.h.in:36)
            del n[pos]
    order = [nid for (_, nid) in node]
    for i in range(len(order) - 1):
        minj, mind = -1, np.float32(3.4028235e38)
        for j in range(i + 1, len(order)):
            d = np.float32(port.distance(dtype, otype, pobj[order[i]], pobj[order[j]]))
            if d < mind:
                minj, mind = j, d
        a, b = order[i], order[minj]
        ins = []
        for (src, dst) in ((a, b), (b, a)):
            n = lists[src]
            pos = bisect.bisect_left(n, (float(mind), dst))
            if pos == len(n) or n[pos][1] != dst:
                n.insert(pos, (float(mind), dst))
                ins.append(True)
            else:
                ins.append(False)
        if i + 1 != minj:
            order[i + 1], order[minj] = order[minj], order[i + 1]
    lists[node_id] = []


def full_sort_lists(old, new):
    """What insertANNGNode's reverse-edge insertion converges to for ONE node's list as the engine's full sort states it
    (csr_from_triples, ngt_b200/csrc/graph_ops.cu; lib/NGT/Graph.h:845-886 inserts by lower_bound on (distance, id) and
    rejects an id that is already there): the keys of the old list and of the batch's new edges sorted ascending, an
    entry dropped when its target (low 32 bits of the key) equals the entry before it. Test infrastructure."""
    out, prev = [], None
    for k in sorted(list(old) + list(new)):
        if prev is None or (k & 0xFFFFFFFF) != (prev & 0xFFFFFFFF):
            out.append(k)
        prev = k
    return out


def merge_lists_by_rank(old, new):
    """The rank merge of merge_lists_kernel (ngt_b200/csrc/graph_ops.cu), restated: an old key moves up by the number of
    new keys below it, a new key lands behind the old keys not above it; `redo` is raised for a list that is not
    strictly ascending / repeats a target in neighbouring entries, or when a new key's target equals the old entry on
    either side of its place or the new key before it -- the cases in which the full sort (above) is taken instead.
    old, new: ascending lists of 64-bit keys (ordered distance bits << 32 | target). -> (merged keys, redo)."""
    import bisect
    redo = False
    out = [None] * (len(old) + len(new))
    for i, ka in enumerate(old):
        if i + 1 < len(old) and (old[i + 1] <= ka or (old[i + 1] & 0xFFFFFFFF) == (ka & 0xFFFFFFFF)):
            redo = True
        out[i + bisect.bisect_left(new, ka)] = ka
    for j, kb in enumerate(new):
        lo = bisect.bisect_right(old, kb)
        t = kb & 0xFFFFFFFF
        out[j + lo] = kb
        if (lo > 0 and (old[lo - 1] & 0xFFFFFFFF) == t) or (lo < len(old) and (old[lo] & 0xFFFFFFFF) == t) or \
                (j > 0 and (new[j - 1] & 0xFFFFFFFF) == t):
            redo = True
    return out, redo

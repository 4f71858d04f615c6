"""ngt_b200.Index -- the host-side mirror of ngtpy.Index (python/src/ngtpy.cpp:28-360, bindings :500-560)
over the B200 engine: same method names, argument meaning, defaults and id numbering, plus the batch entry
points a GPU needs (batch_search / batch_linear_search). Index directories are NGT's own (`prf`, `obj`,
`grp`; see index_io.py), so an index built by the reference's `ngt create` opens here and vice versa.

What differs, on purpose: seeds come from the device seed table instead of the DVP-tree,
`build_index` builds the graph from an exact kNN pass on the device instead of incremental insertion, and
a single-query `search` is a batch of one.
"""
import os

import numpy as np

from . import _lib, build, index_io
from .engine import NORMALIZED, GpuIndex
from ._lib import NgtGpuError

FLT_MAX = 3.4028234663852886e38
INT_MIN = -2 ** 31

# python/src/ngtpy.cpp:59-99
_DISTANCE_ARG = {
    "L2": "L2", "Normalized L2": "NormalizedL2", "Hamming": "Hamming", "Angle": "Angle",
    "Normalized Angle": "NormalizedAngle", "Cosine": "Cosine", "Normalized Cosine": "NormalizedCosine",
}
_OBJECT_ARG = {"Float": "Float-4", "float": "Float-4", "Byte": "Integer-1", "byte": "Integer-1"}


def create(path, dimension, edge_size_for_creation=10, edge_size_for_search=40, distance_type="L2",
           object_type="Float"):
    """ngtpy.create: an empty index directory."""
    if object_type not in _OBJECT_ARG:
        raise NgtGpuError(_lib.ERR_INVALID, "ngtpy::create: invalid object type. %s" % object_type)
    if distance_type not in _DISTANCE_ARG:
        raise NgtGpuError(_lib.ERR_INVALID, "ngtpy::create: invalid distance type. %s" % distance_type)
    prop = dict(index_io.DEFAULT_PRF)
    prop.update({"Dimension": str(int(dimension)), "EdgeSizeForCreation": str(int(edge_size_for_creation)),
                 "EdgeSizeForSearch": str(int(edge_size_for_search)), "DistanceType": _DISTANCE_ARG[distance_type],
                 "ObjectType": _OBJECT_ARG[object_type]})
    os.makedirs(path, exist_ok=True)
    dt = index_io.object_dtype(prop)
    index_io.write_prf(path, prop)
    index_io.write_objects(path, np.zeros((0, int(dimension)), dt))
    index_io.write_graph(path, np.zeros(2, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.float32))


def epsilon_from_accuracy_table(table, accuracy):
    """Index::AccuracyTable::set + getEpsilon (lib/NGT/Index.h:299-347): "epsilon:accuracy,..." pairs; linear through
    the two entries around `accuracy` (the last two when it lies above the table, the first two below), clamped at -0.9."""
    tokens = [t for t in table.split(",") if t]
    pts = []
    if len(tokens) >= 2:
        for tok in tokens:
            ts = tok.split(":")
            if len(ts) != 2:
                raise NgtGpuError(_lib.ERR_INVALID, "AccuracyTable: Invalid accuracy table string %s:%s" % (tok, table))
            pts.append((float(np.float32(float(ts[0]))), float(ts[1])))
    if len(pts) <= 2:
        raise NgtGpuError(_lib.ERR_STATE, "AccuracyTable: The accuracy table is not set yet. The table size=%d" % len(pts))
    accuracy = min(float(accuracy), 1.0)
    i = 0
    while i < len(pts) and pts[i][1] < accuracy:
        i += 1
    if i == len(pts):
        i -= 2
    elif i != 0:
        i -= 1
    lo, up = pts[i], pts[i + 1]
    e = float(np.float32(lo[0] + (up[0] - lo[0]) * (accuracy - lo[1]) / (up[1] - lo[1])))
    return max(e, float(np.float32(-0.9)))


class Index:
    def __init__(self, path, read_only=False, zero_based_numbering=True, tree_disabled=False, log_disabled=False,
                 device=0, n_pivots=4096):
        self.path = path
        self.read_only = read_only
        self.zero_numbering = zero_based_numbering
        self.device = device
        self.n_pivots = n_pivots
        self.prop = index_io.read_prf(path)
        self.object_type = index_io.OBJECT_TYPE_NAMES[self.prop["ObjectType"]]
        self.distance_type = index_io.DISTANCE_TYPE_NAMES[self.prop["DistanceType"]]
        self.dimension = int(self.prop["Dimension"])
        self._dtype = index_io.object_dtype(self.prop)
        rows, present = index_io.read_objects(path, self.prop)
        self._objects = rows                      # ids 1..n, as stored (normalised when the space normalises)
        self._present = present
        row_ptr, col, dist, _ = index_io.read_graph(path)
        self._graph = self._fit_graph(row_ptr, col, dist)
        self._pending = 0                         # appended objects not yet in the graph
        self._raw_from = None                     # first id whose row is not normalised yet (Normalized* types)
        self._gpu = GpuIndex(self.object_type, self.distance_type, self.dimension, device)
        self._num_dist = 0
        # ngtpy defaults, python/src/ngtpy.cpp:43-48
        self.default_size = 20
        self.default_epsilon = 0.1
        self.default_radius = FLT_MAX
        self.default_edge_size = -1
        self.default_expected_accuracy = -1.0
        self._upload()

    # ---- internals -------------------------------------------------------------------------------
    def _fit_graph(self, row_ptr, col, dist):
        n = self._objects.shape[0]
        rp = np.zeros(n + 2, np.uint64)
        m = min(row_ptr.size, n + 2)
        rp[:m] = row_ptr[:m]
        rp[m:] = row_ptr[-1]
        return rp, col, dist

    def _upload(self):
        n = self._objects.shape[0]
        if n == 0:
            return
        self._gpu.set_objects(self._objects, normalize=False)
        removed = np.nonzero(self._present[1:] == 0)[0].astype(np.uint32) + 1
        if removed.size:
            self._gpu.set_removed(removed)
        rp, col, _ = self._fit_graph(*self._graph)     # objects appended since the last build have no edges yet
        self._gpu.set_graph(rp, col)
        self._gpu.set_search_property(int(self.prop.get("EdgeSizeForSearch", 40)),
                                      int(self.prop.get("DynamicEdgeSizeBase", 30)),
                                      int(self.prop.get("DynamicEdgeSizeRate", 20)))
        self._gpu.build_seed_table(min(self.n_pivots, n), 1)

    def _ids_out(self, ids):
        return ids.astype(np.int64) - 1 if self.zero_numbering else ids.astype(np.int64)

    def _epsilon_from_accuracy(self, accuracy):
        """Index::getEpsilonFromExpectedAccuracy (lib/NGT/Index.h:293-360): piecewise-linear table in prf."""
        return epsilon_from_accuracy_table(self.prop.get("AccuracyTable", ""), accuracy)

    def _params(self, size, epsilon, edge_size, expected_accuracy):
        size = self.default_size if size == 0 else int(size)
        if expected_accuracy is not None and expected_accuracy > 0.0:
            epsilon = self._epsilon_from_accuracy(expected_accuracy)
        elif epsilon is None or epsilon <= -1.0:
            epsilon = self.default_epsilon
        if edge_size is None or edge_size < -2:
            edge_size = self.default_edge_size
        return size, float(epsilon), int(edge_size)

    # ---- ngtpy.Index surface ---------------------------------------------------------------------
    def set(self, num_of_search_objects=0, search_radius=-FLT_MAX, epsilon=-FLT_MAX, edge_size=INT_MIN,
            expected_accuracy=-FLT_MAX):
        if num_of_search_objects > 0:
            self.default_size = num_of_search_objects
        if search_radius > -FLT_MAX:
            self.default_radius = search_radius
        if epsilon > -FLT_MAX:
            self.default_epsilon = epsilon
        if edge_size >= -2:
            self.default_edge_size = edge_size
        if expected_accuracy > -FLT_MAX:
            self.default_expected_accuracy = expected_accuracy

    def batch_search(self, queries, size=0, epsilon=-FLT_MAX, edge_size=INT_MIN, expected_accuracy=-FLT_MAX,
                     with_stats=False):
        """queries [nq, dim] -> (ids [nq,size] (numbering per zero_based_numbering, -1 where fewer results),
        distances [nq,size]); the batch form of search()."""
        if self._pending:
            raise NgtGpuError(_lib.ERR_STATE, "objects were appended: call build_index() before searching")
        size, epsilon, edge_size = self._params(size, epsilon, edge_size, expected_accuracy)
        radius = -1.0 if self.default_radius >= FLT_MAX else self.default_radius
        seed_size = int(self.prop.get("SeedSize", 10)) or 10
        out = self._gpu.search(queries, size, epsilon, radius, edge_size, n_seeds=seed_size, with_stats=True)
        ids, dists, counts, stats = out
        self._num_dist += int(stats[:, 0].sum())
        res = self._ids_out(ids)
        res[np.arange(size)[None, :] >= counts[:, None]] = -1
        return (res, dists, stats) if with_stats else (res, dists)

    def batch_linear_search(self, queries, size=0):
        size = self.default_size if size == 0 else int(size)
        radius = -1.0 if self.default_radius >= FLT_MAX else self.default_radius
        ids, dists, counts = self._gpu.linear_search(queries, size, radius)
        res = self._ids_out(ids)
        res[np.arange(size)[None, :] >= counts[:, None]] = -1
        return res, dists

    def _one(self, ids, dists, with_distance):
        ok = ids[0] >= (0 if self.zero_numbering else 1)
        if not with_distance:
            return ids[0][ok].astype(np.int32)
        return [(int(i), float(d)) for i, d in zip(ids[0][ok], dists[0][ok])]

    def search(self, query, size=0, epsilon=-FLT_MAX, edge_size=INT_MIN, expected_accuracy=-FLT_MAX,
               with_distance=True):
        q = np.asarray(query, np.float32).reshape(1, -1)
        ids, dists = self.batch_search(q, size, epsilon, edge_size, expected_accuracy)
        return self._one(ids, dists, with_distance)

    def linear_search(self, query, size=0, with_distance=True):
        q = np.asarray(query, np.float32).reshape(1, -1)
        ids, dists = self.batch_linear_search(q, size)
        return self._one(ids, dists, with_distance)

    def get_num_of_distance_computations(self):
        return self._num_dist

    def get_object(self, object_id):
        oid = object_id + 1 if self.zero_numbering else object_id
        if oid < 1 or oid > self._objects.shape[0] or not self._present[oid]:
            raise NgtGpuError(_lib.ERR_INVALID, "get_object: no such object %d" % object_id)
        return [float(v) for v in self._objects[oid - 1]]

    def _append(self, objects):
        x = np.asarray(objects, np.float64)
        if x.ndim == 1:
            x = x[None, :]
        if x.shape[1] != self.dimension:
            raise NgtGpuError(_lib.ERR_INVALID, "ngtpy::insert: Error! dimensions are inconsitency. %d:%d" % (
                self.dimension, x.shape[1]))
        x = x.astype(np.float32)
        if self.distance_type in NORMALIZED:
            # ObjectSpace::normalize (lib/NGT/ObjectSpace.h:251-266) happens on the device at upload time:
            # rows are stored raw here and replaced by the device's normalised rows in build_index()
            if (np.abs(x).sum(axis=1) == 0).any():
                raise NgtGpuError(_lib.ERR_ZERO_VECTOR, "normalize: a zero vector cannot be normalised")
        first = self._objects.shape[0] + 1
        self._objects = np.concatenate([self._objects, x.astype(self._dtype)], axis=0)
        self._present = np.concatenate([self._present, np.ones(x.shape[0], np.uint8)])
        self._pending += x.shape[0]
        self._raw_from = first if self._raw_from is None else min(self._raw_from, first)
        return first

    def insert(self, object, debug=False):
        first = self._append(np.asarray(object).reshape(1, -1))
        self._num_dist = 0
        return first - 1 if self.zero_numbering else first

    def batch_insert(self, objects, num_threads=8, debug=False):
        self._append(objects)
        self.build_index(num_threads)
        self._num_dist = 0

    def build_index(self, num_threads=8, target_size_of_graph=0):
        """NGT::Index::createIndex. An index without a graph gets one over all present objects: the reference inserts
        object by object with approximate searches (lib/NGT/Index.cpp:721-792); here the edge candidates of
        every node come from one exact kNN pass on the device and the ANNG is its symmetric closure
        (out-edges + reverse edges, sorted by (distance,id)) -- what insertANNGNode converges to
        (lib/NGT/Graph.h:611-626). Objects appended to an index that already has its graph (loaded ONNG, refined or
        optimised graph included) are INSERTED into it with the reference's construction loop on the device
        (ngtgpu_index_insert_batch), as ngt_create_index of the C API does; with nothing queued it is a no-op
        (createIndex only indexes objects that are not in the graph yet, Index.cpp:645-648)."""
        import torch
        n = self._objects.shape[0]
        if n == 0:
            return
        have_graph = self._graph[1].size > 0
        if self._pending == 0 and have_graph:
            return
        normalize = self.distance_type in NORMALIZED and self._raw_from is not None
        if normalize:
            # normalise only the newly appended rows; stored rows are already unit length
            first = self._raw_from
            tmp = GpuIndex(self.object_type, self.distance_type, self.dimension, self.device)
            tmp.set_objects(self._objects[first - 1:], normalize=True)
            self._objects[first - 1:] = tmp.get_objects(1, n - first + 1)
            tmp.close()
        self._raw_from = None
        self._gpu.set_objects(self._objects, normalize=False)
        removed = np.nonzero(self._present[1:] == 0)[0].astype(np.uint32) + 1
        if removed.size:
            self._gpu.set_removed(removed)
        e = int(self.prop.get("EdgeSizeForCreation", 10))
        if have_graph and 0 < self._pending < n:
            dev = torch.device("cuda", self.device)
            n_old = n - self._pending
            rp, col, dist = self._graph
            rp_new = np.full(n + 2, rp[n_old + 1], np.int64)
            rp_new[:n_old + 2] = rp[:n_old + 2].astype(np.int64)
            g = (torch.from_numpy(rp_new).to(dev), torch.from_numpy(col.astype(np.int32)).to(dev),
                 torch.from_numpy(np.asarray(dist, np.float32)).to(dev))
            self._gpu.set_search_property(int(self.prop.get("EdgeSizeForSearch", 40)),
                                          int(self.prop.get("DynamicEdgeSizeBase", 30)),
                                          int(self.prop.get("DynamicEdgeSizeRate", 20)))
            g = build.insert_objects(self._gpu, n_old + 1, self._pending, g, max(e, 1),
                                     float(self.prop.get("EpsilonForCreation", 0.1)), -1,
                                     max(int(self.prop.get("BatchSizeForCreation", 200)), 1), 10, 1024, 1)
            torch.cuda.synchronize()
            self._graph = (g[0].cpu().numpy().astype(np.uint64), g[1].cpu().numpy().astype(np.uint32), g[2].cpu().numpy())
            self._pending = 0
            self._gpu.build_seed_table(min(self.n_pivots, n), 1)
            return
        k = min(e, max(n - 1 - removed.size, 1))
        ids, dists, counts = build.knn_graph(self._gpu, k)
        row_ptr, col, dist = build.reconstruct_graph(ids, dists, counts, k, k)
        torch.cuda.synchronize()
        self._graph = (row_ptr.cpu().numpy().astype(np.uint64), col.cpu().numpy().astype(np.uint32),
                       dist.cpu().numpy())
        self._pending = 0
        rp, c, _ = self._graph
        self._gpu.set_graph(rp, c)
        self._gpu.set_search_property(int(self.prop.get("EdgeSizeForSearch", 40)),
                                      int(self.prop.get("DynamicEdgeSizeBase", 30)),
                                      int(self.prop.get("DynamicEdgeSizeRate", 20)))
        self._gpu.build_seed_table(min(self.n_pivots, n), 1)
        self.prop["GraphType"] = "ANNG"

    def refine_anng(self, epsilon=0.1, expected_accuracy=0.0, num_of_edges=0, num_of_explored_edges=INT_MIN,
                    batch_size=10000):
        """ngtpy.Index.refine_anng (python/src/ngtpy.cpp:548-553 -> GraphReconstructor::refineANNG,
        lib/NGT/GraphReconstructor.h:814-924) on the device."""
        import torch
        if self._pending:
            raise NgtGpuError(_lib.ERR_STATE, "objects were appended: call build_index() first")
        if expected_accuracy > 0.0:     # GraphReconstructor.h:843-847 -> Index.h:1156-1158
            epsilon = self._epsilon_from_accuracy(expected_accuracy)
        dev = torch.device("cuda", self.device)
        rp, col, dist = self._graph
        g = build.refine_anng(self._gpu, torch.from_numpy(rp.astype(np.int64)).to(dev),
                              torch.from_numpy(col.astype(np.int32)).to(dev), torch.from_numpy(dist).to(dev),
                              epsilon=epsilon, no_of_edges=num_of_edges,
                              edge_size=-1 if num_of_explored_edges == INT_MIN else num_of_explored_edges,
                              batch_size=batch_size, edge_size_for_creation=int(self.prop.get("EdgeSizeForCreation", 10)))
        self._graph = (g[0].cpu().numpy().astype(np.uint64), g[1].cpu().numpy().astype(np.uint32), g[2].cpu().numpy())

    def remove(self, object_id):
        oid = object_id + 1 if self.zero_numbering else object_id
        if oid < 1 or oid > self._objects.shape[0] or not self._present[oid]:
            raise NgtGpuError(_lib.ERR_INVALID, "remove: no such object %d" % object_id)
        self._remove_edges_reliably(oid)
        self._present[oid] = 0
        self._upload()

    def _remove_edges_reliably(self, oid):
        """NeighborhoodGraph::removeEdgesReliably (lib/NGT/Graph.cpp:641-864), as ngt_remove_index of the C API does it:
        back edges go, the neighbours are chained nearest-first with distances from the device; a missing back edge is
        skipped like the reference's NGT_FORCED_REMOVE build does."""
        import bisect
        rp, col, dist = self._graph
        n_graph = rp.size - 2
        if oid > n_graph:
            return
        lists = {}

        def get(nid):
            if nid not in lists:
                lists[nid] = [(float(dist[e]), int(col[e])) for e in range(int(rp[nid]), int(rp[nid + 1]))]
            return lists[nid]
        node = [e for e in get(oid) if e[1] != oid]
        for (d, nid) in node:
            lst = get(nid)
            pos = bisect.bisect_left(lst, (d, oid))
            if pos < len(lst) and lst[pos][1] == oid:
                del lst[pos]
        order = [nid for (_, nid) in node]
        m = len(order)
        if m > 1:
            D = self._gpu.pairwise_distances(np.array(order, np.uint32))
            slot = list(range(m))
            for i in range(m - 1):
                row = D[slot[i]]
                minj, mind = -1, np.float32(3.4028235e38)
                for j in range(i + 1, m):
                    if row[slot[j]] < mind:
                        minj, mind = j, row[slot[j]]
                a, b = order[i], order[minj]
                for (src, dst) in ((a, b), (b, a)):
                    lst = get(src)
                    pos = bisect.bisect_left(lst, (float(mind), dst))
                    if pos == len(lst) or lst[pos][1] != dst:
                        lst.insert(pos, (float(mind), dst))
                if i + 1 != minj:
                    order[i + 1], order[minj] = order[minj], order[i + 1]
                    slot[i + 1], slot[minj] = slot[minj], slot[i + 1]
        lists[oid] = []
        # (edges to the removed node from lists it does not name itself -- asymmetric graphs -- go as well)
        src_all = np.repeat(np.arange(rp.size - 1), np.diff(rp.astype(np.int64)))
        alive = col != oid
        deg = np.bincount(src_all[alive], minlength=rp.size - 1).astype(np.int64)
        for nid, lst in lists.items():
            deg[nid] = len(lst)
        nrp = np.zeros_like(rp)
        nrp[1:] = np.cumsum(deg)
        ncol = np.zeros(int(nrp[-1]), col.dtype)
        ndist = np.zeros(int(nrp[-1]), dist.dtype)
        touched = np.zeros(rp.size - 1, bool)
        touched[list(lists.keys())] = True
        keep = ~touched[src_all] & alive
        rank = np.cumsum(keep) - 1                                  # position among the kept entries ...
        first = np.zeros(rp.size - 1, np.int64)
        kept_deg = np.bincount(src_all[keep], minlength=rp.size - 1)
        first[1:] = np.cumsum(kept_deg)[:-1]                          # ... minus the kept entries of earlier lists
        dst_pos = (nrp[:-1].astype(np.int64)[src_all] + (rank - first[src_all]))[keep]
        ncol[dst_pos] = col[keep]
        ndist[dst_pos] = dist[keep]
        for nid, lst in lists.items():
            b = int(nrp[nid])
            for t, (d, c) in enumerate(lst):
                ncol[b + t], ndist[b + t] = c, d
        self._graph = (nrp, ncol, ndist)

    def save(self):
        if self._pending:
            raise NgtGpuError(_lib.ERR_STATE, "objects were appended: call build_index() before save()")
        self.prop["IndexType"] = "Graph"      # no `tre` is written; the reference opens Graph indexes without one
        index_io.write_prf(self.path, self.prop)
        if os.path.exists(os.path.join(self.path, "tre")):   # a DVP-tree the reference left here describes another object set
            os.remove(os.path.join(self.path, "tre"))
        index_io.write_objects(self.path, self._objects, self._present)
        rp, col, dist = self._graph
        index_io.write_graph(self.path, rp, col, dist, self._present)

    def close(self):
        self._gpu.close()


class Optimizer:
    """ngtpy.Optimizer (python/src/ngtpy.cpp:566-607 over NGT::GraphOptimizer, lib/NGT/GraphOptimizer.h): `execute`
    runs reconstructGraph + shortcut reduction on the device and writes the ONNG in NGT's file format. The
    search-parameter tuning steps (timed host searches) are outside the hot path: `adjust_search_coefficients` and
    `optimize_search_parameters` raise."""

    def __init__(self, num_of_outgoings=-1, num_of_incomings=-1, num_of_queries=-1, num_of_objects=-1,
                 low_accuracy_from=-1.0, low_accuracy_to=-1.0, high_accuracy_from=-1.0, high_accuracy_to=-1.0,
                 gt_epsilon=-1.7976931348623157e308, margin=-1.0, log_disabled=False, device=0):
        # GraphOptimizer::init, lib/NGT/GraphOptimizer.h:60-78
        self.outgoing, self.incoming, self.min_edges = 10, 120, 0
        self.shortcut_reduction = True
        self.device = device
        self.set(num_of_outgoings, num_of_incomings)

    def set(self, num_of_outgoings=-1, num_of_incomings=-1, *_, **__):
        if num_of_outgoings >= 0:
            self.outgoing = num_of_outgoings
        if num_of_incomings >= 0:
            self.incoming = num_of_incomings

    def set_processing_modes(self, shortcut_reduction=True, search_parameter_optimization=True,
                             prefetch_parameter_optimization=True, accuracy_table_generation=True):
        self.shortcut_reduction = bool(shortcut_reduction)

    def execute(self, in_path, out_path):
        import shutil
        import torch
        if os.path.exists(out_path):
            raise NgtGpuError(_lib.ERR_INVALID, "Optimizer::execute: The specified index exists. " + out_path)
        shutil.copytree(in_path, out_path)
        prop = index_io.read_prf(out_path)
        row_ptr, col, dist, present = index_io.read_graph(out_path)
        dev = torch.device("cuda", self.device)
        g = (torch.from_numpy(row_ptr.astype(np.int64)).to(dev), torch.from_numpy(col.astype(np.int32)).to(dev),
             torch.from_numpy(dist).to(dev))
        if self.outgoing > 0 or self.incoming > 0:
            if prop.get("GraphType", "ANNG") != "ANNG":          # convertToANNG, GraphReconstructor.h:389-423
                g = build.reconstruct_graph_device(g[0], g[1], g[2], 0xffffffff, 0xffffffff)
            g = build.reconstruct_graph_device(g[0], g[1], g[2], max(self.outgoing, 0), max(self.incoming, 0))
            prop["GraphType"] = "ONNG"
        if self.shortcut_reduction:
            g = build.adjust_paths(g[0], g[1], g[2], self.min_edges)
        index_io.write_graph(out_path, g[0].cpu().numpy().astype(np.uint64), g[1].cpu().numpy().astype(np.uint32),
                             g[2].cpu().numpy(), present)
        index_io.write_prf(out_path, prop)

    def adjust_search_coefficients(self, path):
        raise NgtGpuError(_lib.ERR_INVALID, "not provided by the B200 engine (search-parameter tuning is outside the hot path)")

    optimize_search_parameters = adjust_search_coefficients

"""GpuIndex -- the device-resident copy of one NGT index (or of one shard of it), over the C ABI.

Thin by design: numpy arrays and torch CUDA tensors go in as raw pointers, results come back as
arrays of the same kind. All arithmetic happens in libngtgpu.so; nothing here computes a distance.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (DISTANCE_ANGLE, DISTANCE_COSINE, DISTANCE_HAMMING, DISTANCE_L2, DISTANCE_NORMALIZED_ANGLE,
                   DISTANCE_NORMALIZED_COSINE, DISTANCE_NORMALIZED_L2, OBJECT_FLOAT, OBJECT_UINT8, NgtGpuError,
                   SearchParams)

NORMALIZED = (DISTANCE_NORMALIZED_ANGLE, DISTANCE_NORMALIZED_COSINE, DISTANCE_NORMALIZED_L2)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _np_type(object_type):
    return np.uint8 if object_type == OBJECT_UINT8 else np.float32


class GpuIndex:
    """Objects 1..n, adjacency lists and the seed table of one index in HBM."""

    def __init__(self, object_type, distance_type, dimension, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.ngtgpu_index_create(C.byref(self._h), int(device), int(object_type), int(distance_type),
                                                 int(dimension)))
        self.object_type = int(object_type)
        self.distance_type = int(distance_type)
        self.dimension = int(dimension)
        self.device = int(device)
        self._keep = []

    # ---- life cycle ------------------------------------------------------------------------------
    def close(self):
        if self._h:
            self._lib.ngtgpu_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self):
        return int(self._lib.ngtgpu_index_size(self._h))

    @property
    def padded_dimension(self):
        return int(self._lib.ngtgpu_index_padded_dimension(self._h))

    @property
    def launch_count(self):
        return int(self._lib.ngtgpu_index_launch_count(self._h))

    @property
    def last_overflows(self):
        return int(self._lib.ngtgpu_index_last_overflows(self._h))

    # ---- data ------------------------------------------------------------------------------------
    def _query_array(self, q):
        """-> (pointer, query_type, nq, on_device, keepalive)"""
        if _is_torch(q):
            import torch
            if q.dim() == 1:
                q = q[None, :]
            if q.dtype not in (torch.float32, torch.uint8):
                q = q.to(torch.float32)
            q = q.contiguous()
            if q.shape[1] != self.dimension:
                raise NgtGpuError(_lib.ERR_INVALID, "dimension mismatch: %d vs %d" % (q.shape[1], self.dimension))
            qt = OBJECT_UINT8 if q.dtype == torch.uint8 else OBJECT_FLOAT
            if not q.is_cuda:
                raise NgtGpuError(_lib.ERR_INVALID, "torch queries must be CUDA tensors (or pass numpy)")
            return q.data_ptr(), qt, q.shape[0], True, q
        q = np.asarray(q)
        if q.ndim == 1:
            q = q[None, :]
        if q.dtype != np.uint8:
            q = q.astype(np.float32, copy=False)
        q = np.ascontiguousarray(q)
        if q.shape[1] != self.dimension:
            raise NgtGpuError(_lib.ERR_INVALID, "dimension mismatch: %d vs %d" % (q.shape[1], self.dimension))
        qt = OBJECT_UINT8 if q.dtype == np.uint8 else OBJECT_FLOAT
        return q.ctypes.data, qt, q.shape[0], False, q

    def set_objects(self, x, normalize=None):
        """x: [n, dim] rows of the object type (numpy, or a torch CUDA tensor already in HBM).
        normalize: divide rows by their norm (default: only for Normalized* distance types, as
        ObjectSpaceRepository does on insertion, lib/NGT/ObjectSpaceRepository.h:560-594)."""
        if normalize is None:
            normalize = self.distance_type in NORMALIZED
        if _is_torch(x):
            import torch
            want = torch.uint8 if self.object_type == OBJECT_UINT8 else torch.float32
            if x.dtype != want:
                x = x.to(want)
            x = x.contiguous()
            assert x.is_cuda and x.dim() == 2 and x.shape[1] == self.dimension
            torch.cuda.current_stream(x.device).synchronize()
            _lib.check(self._lib.ngtgpu_index_set_objects(self._h, x.data_ptr(), x.shape[0], int(normalize), 1))
            return
        x = np.ascontiguousarray(np.asarray(x).astype(_np_type(self.object_type), copy=False))
        assert x.ndim == 2 and x.shape[1] == self.dimension, x.shape
        _lib.check(self._lib.ngtgpu_index_set_objects(self._h, x.ctypes.data, x.shape[0], int(normalize), 0))

    def set_removed(self, ids):
        ids = np.ascontiguousarray(ids, np.uint32)
        _lib.check(self._lib.ngtgpu_index_set_removed(self._h, ids.ctypes.data, ids.size))

    def set_graph(self, row_ptr, col):
        """CSR over ids 0..n (row_ptr has n+2 entries); lists in `grp` order."""
        if _is_torch(row_ptr):
            import torch
            row_ptr = row_ptr.to(torch.int64).contiguous()
            col = col.to(torch.int32).contiguous() if col.dtype != torch.int32 else col.contiguous()
            torch.cuda.current_stream(col.device).synchronize()
            _lib.check(self._lib.ngtgpu_index_set_graph(self._h, row_ptr.data_ptr(), col.data_ptr(), 1))
            return
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        if row_ptr.size != self.size + 2:
            raise NgtGpuError(_lib.ERR_INVALID, "row_ptr must have n+2 = %d entries, got %d" % (self.size + 2, row_ptr.size))
        _lib.check(self._lib.ngtgpu_index_set_graph(self._h, row_ptr.ctypes.data, col.ctypes.data, 0))

    def set_search_property(self, edge_size_for_search=40, dynamic_edge_size_base=30, dynamic_edge_size_rate=20):
        _lib.check(self._lib.ngtgpu_index_set_search_property(self._h, int(edge_size_for_search),
                                                              int(dynamic_edge_size_base), int(dynamic_edge_size_rate)))

    def set_search_workspace(self, hash_bits=14, queue_cap=512, onchip_tiers=2, stage_bytes=None):
        _lib.check(self._lib.ngtgpu_index_set_search_workspace(self._h, int(hash_bits), int(queue_cap)))
        _lib.check(self._lib.ngtgpu_index_set_onchip_tiers(self._h, int(onchip_tiers)))
        if stage_bytes is not None:
            _lib.check(self._lib.ngtgpu_index_set_stage_bytes(self._h, int(stage_bytes)))

    def set_fast_kernel(self, enabled=True):
        """First traversal tier of the common case on the lean kernel (default) or on the general one."""
        _lib.check(self._lib.ngtgpu_index_set_fast_kernel(self._h, int(bool(enabled))))

    def set_fast_shape(self, warps_per_query=0, ctas_per_sm=0):
        """Warps per query of the lean kernel (0: by row width; 1, 2, 4) and a cap on its resident CTAs per SM (0: what fits)."""
        _lib.check(self._lib.ngtgpu_index_set_fast_shape(self._h, int(warps_per_query), int(ctas_per_sm)))

    def set_seed_fusion(self, enabled=True):
        """Seed selection inside the lean traversal kernel (True) or by its own launch (default)."""
        _lib.check(self._lib.ngtgpu_index_set_seed_fusion(self._h, int(bool(enabled))))

    def set_tensor_core(self, enabled=True):
        _lib.check(self._lib.ngtgpu_index_set_tensor_core(self._h, int(bool(enabled))))

    @property
    def tensor_core_batches(self):
        return int(self._lib.ngtgpu_index_tensor_core_batches(self._h))

    def build_onng(self, knn, outgoing=10, incoming=120, shortcut_reduction=True, min_edges=0, want_graph=False):
        """The reference's ONNG recipe on the device for this index's objects (ngtgpu_index_build_onng): exact kNN table,
        reconstructGraph(outgoing, incoming), path adjustment; the graph is left set on the index.
        -> {"knn_s", "reconstruct_s", "adjust_paths_s"[, "graph": (row_ptr, col, dist) torch CUDA tensors]}"""
        import ctypes as C

        class _Graph(C.Structure):
            _fields_ = [("n", C.c_uint64), ("nnz", C.c_uint64), ("row_ptr", C.c_void_p), ("col", C.c_void_p), ("dist", C.c_void_p)]
        g = _Graph()
        sec = (C.c_double * 3)()
        _lib.check(self._lib.ngtgpu_index_build_onng(self._h, int(knn), int(outgoing), int(incoming), int(bool(shortcut_reduction)),
                                                     int(min_edges), C.byref(g) if want_graph else None, sec))
        out = {"knn_s": sec[0], "reconstruct_s": sec[1], "adjust_paths_s": sec[2]}
        if want_graph:
            import torch
            dev = torch.device("cuda", self.device)
            rp = torch.empty(int(g.n) + 2, dtype=torch.int64, device=dev)
            col = torch.empty(int(g.nnz), dtype=torch.int32, device=dev)
            dist = torch.empty(int(g.nnz), dtype=torch.float32, device=dev)
            # device-to-device copies into torch-owned storage, then the library's buffers are released
            for dst, src, nbytes in ((rp, g.row_ptr, (int(g.n) + 2) * 8), (col, g.col, int(g.nnz) * 4), (dist, g.dist, int(g.nnz) * 4)):
                if nbytes:
                    _lib.check(self._lib.ngtgpu_device_copy(dst.data_ptr(), src, nbytes))
                _lib.check(self._lib.ngtgpu_device_free(src))
            out["graph"] = (rp, col, dist)
        return out

    def build_seed_table(self, n_pivots=4096, rng_seed=1):
        _lib.check(self._lib.ngtgpu_index_build_seed_table(self._h, int(n_pivots), int(rng_seed)))

    def get_object(self, object_id):
        out = np.zeros(self.dimension, _np_type(self.object_type))
        _lib.check(self._lib.ngtgpu_index_get_object(self._h, int(object_id), out.ctypes.data))
        return out

    def pairwise_distances(self, ids):
        """[m, m] float32 distances among the stored objects `ids` (the engine's exact distance)."""
        ids = np.ascontiguousarray(ids, np.uint32)
        out = np.zeros((ids.size, ids.size), np.float32)
        _lib.check(self._lib.ngtgpu_index_pairwise_distances(self._h, ids.ctypes.data, ids.size, out.ctypes.data))
        return out

    def get_objects(self, first, count):
        """Stored rows first .. first+count-1, one strided copy."""
        out = np.zeros((int(count), self.dimension), _np_type(self.object_type))
        _lib.check(self._lib.ngtgpu_index_get_objects(self._h, int(first), int(count), out.ctypes.data))
        return out

    # ---- the hot path ----------------------------------------------------------------------------
    def search(self, queries, size=10, epsilon=0.1, radius=-1.0, edge_size=-1, seeds=None, n_seeds=10,
               with_stats=False):
        """Batched GraphIndex::search. Returns (ids [nq,size] uint32, dists [nq,size] float32,
        counts [nq] uint32[, stats [nq,3] uint32]); torch CUDA tensors if the queries were one."""
        ptr, qt, nq, on_dev, keep = self._query_array(queries)
        p = SearchParams(int(size), float(epsilon), float(radius), int(edge_size))
        k = max(int(size), 1)
        if on_dev:
            import torch
            dev = keep.device
            ids = torch.zeros((nq, k), dtype=torch.int32, device=dev)
            dists = torch.zeros((nq, k), dtype=torch.float32, device=dev)
            counts = torch.zeros((nq,), dtype=torch.int32, device=dev)
            stats = torch.zeros((nq, 3), dtype=torch.int32, device=dev) if with_stats else None
            sp, ns = None, int(n_seeds)
            if seeds is not None:
                seeds = seeds.to(torch.int32).contiguous()
                sp, ns = seeds.data_ptr(), seeds.shape[1]
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self._lib.ngtgpu_search_device(self._h, ptr, qt, nq, C.byref(p), sp, ns, ids.data_ptr(),
                                                      dists.data_ptr(), counts.data_ptr(),
                                                      stats.data_ptr() if with_stats else None, stream))
        else:
            ids = np.zeros((nq, k), np.uint32)
            dists = np.zeros((nq, k), np.float32)
            counts = np.zeros(nq, np.uint32)
            stats = np.zeros((nq, 3), np.uint32) if with_stats else None
            sp, ns = None, int(n_seeds)
            if seeds is not None:
                seeds = np.ascontiguousarray(seeds, np.uint32)
                assert seeds.ndim == 2 and seeds.shape[0] == nq
                sp, ns = seeds.ctypes.data, seeds.shape[1]
            _lib.check(self._lib.ngtgpu_search(self._h, ptr, qt, nq, C.byref(p), sp, ns, ids.ctypes.data,
                                               dists.ctypes.data, counts.ctypes.data,
                                               stats.ctypes.data if with_stats else None))
        if int(size) == 0:
            ids, dists = ids[:, :0], dists[:, :0]
        return (ids, dists, counts, stats) if with_stats else (ids, dists, counts)

    def linear_search(self, queries, size=10, radius=-1.0):
        """Batched ObjectSpace::linearSearch: exhaustive, exact."""
        ptr, qt, nq, on_dev, keep = self._query_array(queries)
        k = max(int(size), 1)
        if on_dev:
            import torch
            dev = keep.device
            ids = torch.zeros((nq, k), dtype=torch.int32, device=dev)
            dists = torch.zeros((nq, k), dtype=torch.float32, device=dev)
            counts = torch.zeros((nq,), dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self._lib.ngtgpu_linear_search_device(self._h, ptr, qt, nq, int(size), float(radius),
                                                             ids.data_ptr(), dists.data_ptr(), counts.data_ptr(), stream))
        else:
            ids = np.zeros((nq, k), np.uint32)
            dists = np.zeros((nq, k), np.float32)
            counts = np.zeros(nq, np.uint32)
            _lib.check(self._lib.ngtgpu_linear_search(self._h, ptr, qt, nq, int(size), float(radius), ids.ctypes.data,
                                                      dists.ctypes.data, counts.ctypes.data))
        if int(size) == 0:
            ids, dists = ids[:, :0], dists[:, :0]
        return ids, dists, counts

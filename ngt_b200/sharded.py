"""Row-sharded search over the GPUs of one box.

Rank g owns the objects with global ids g*n_local+1 .. (g+1)*n_local as its own index (own graph, own seed
table; no cross-shard edges). Every rank gets the full query batch, searches its shard with the same
epsilon / k, and the per-shard top-k lists are exchanged with ONE all-gather (NCCL over NVLink) and merged per
query by (distance, id) -- the order of ObjectDistance, lib/NGT/Common.h:1946-1952 -- which is what a single
priority queue over the union keeps. The reference has no distribution; this is the
B200-native addition.

On GPUs the whole exchange lives in libngtgpu.so (csrc/shard.cu): the traversal kernel writes its results as keys into
this rank's slot of the gather buffer, ncclAllGather runs in place, a device kernel merges. `LibShardedSearcher` is the
one-process-per-GPU binding (torchrun: torch.distributed only carries NCCL's 128-byte id from rank 0), `ShardedIndex`
the one-process-many-GPUs handle a C host uses. `ShardedSearcher` / `all_gather_keys` / `*_host` are the same plumbing
written against torch.distributed (gloo in the CPU tests): lists travel as int64 keys = ordered distance bits << 32 |
global id, so one flat tensor per rank is gathered.
"""
import ctypes as C

import numpy as np

from . import _lib

KEY_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)

_ready = False


def _fn():
    global _ready
    lib = _lib.load()
    if not _ready:
        P = C.c_void_p
        lib.ngtgpu_pack_keys.argtypes = [P, P, P, C.c_uint32, C.c_uint32, C.c_uint32, P, P]
        lib.ngtgpu_merge_keys.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, P, P, P, P]
        _ready = True
    return lib


# ---- the same key format on the host (numpy), used by the CPU tests --------------------------------------
def pack_keys_host(ids, dists, counts, id_offset):
    """[nq,k] ids/float32 dists/counts -> uint64 keys; order of keys == order of (distance, id)."""
    d = (np.asarray(dists, np.float32) + np.float32(0.0)).view(np.uint32).astype(np.uint64)
    neg = (d & np.uint64(0x80000000)) != 0
    o = np.where(neg, (~d) & np.uint64(0xFFFFFFFF), d | np.uint64(0x80000000))
    keys = (o << np.uint64(32)) | (np.asarray(ids, np.uint64) + np.uint64(id_offset))
    k = keys.shape[1]
    keys[np.arange(k)[None, :] >= np.asarray(counts)[:, None]] = KEY_NONE
    return keys


def unpack_keys_host(keys):
    keys = np.asarray(keys, np.uint64)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    b = np.where((o & np.uint32(0x80000000)) != 0, o & np.uint32(0x7FFFFFFF), ~o)
    ids = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    valid = keys != KEY_NONE
    return np.where(valid, ids, 0).astype(np.uint32), np.where(valid, b.view(np.float32), 0).astype(np.float32), valid


def merge_keys_host(gathered, k):
    """gathered: [n_lists, nq, k] uint64 -> (ids, dists, counts): the k smallest keys per query."""
    g = np.asarray(gathered, np.uint64)
    allk = np.sort(np.transpose(g, (1, 0, 2)).reshape(g.shape[1], -1), axis=1)[:, :k]
    ids, dists, valid = unpack_keys_host(allk)
    return ids, dists, valid.sum(1).astype(np.uint32)


def all_gather_keys(keys, world):
    """keys: torch int64 [nq, k] on this rank's device -> [world, nq, k] on every rank (one collective)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(out, keys.contiguous()) if keys.is_cuda else dist.all_gather(list(out.unbind(0)), keys.contiguous())
    return out


class ShardedSearcher:
    """One per rank. `ix` is the GpuIndex of this rank's shard (local ids 1..n_local)."""

    def __init__(self, ix, rank, world, n_local):
        self.ix, self.rank, self.world, self.n_local = ix, rank, world, n_local
        self.id_offset = rank * n_local

    def _merge(self, ids, dists, counts, k):
        import torch
        lib = _fn()
        dev = ids.device
        nq = ids.shape[0]
        stream = torch.cuda.current_stream(dev).cuda_stream
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        _lib.check(lib.ngtgpu_pack_keys(ids.data_ptr(), dists.data_ptr(), counts.data_ptr(), nq, k, self.id_offset,
                                        keys.data_ptr(), stream))
        gathered = all_gather_keys(keys, self.world)
        out_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_c = torch.empty((nq,), dtype=torch.int32, device=dev)
        _lib.check(lib.ngtgpu_merge_keys(gathered.data_ptr(), self.world, nq, k, out_ids.data_ptr(), out_d.data_ptr(),
                                         out_c.data_ptr(), stream))
        return out_ids, out_d, out_c

    def search(self, queries, k, epsilon, edge_size=-1, n_seeds=10):
        """queries: torch CUDA tensor [nq, dim], identical on every rank -> merged (global ids, dists, counts)."""
        ids, dists, counts = self.ix.search(queries, k, epsilon, edge_size=edge_size, n_seeds=n_seeds)
        return self._merge(ids, dists, counts, k)

    def linear_search(self, queries, k, radius=-1.0):
        ids, dists, counts = self.ix.linear_search(queries, k, radius)
        return self._merge(ids, dists, counts, k)


class LibShardedSearcher:
    """One per rank (one process per GPU). Search, all-gather and merge all run inside libngtgpu.so
    (ngtgpu_shard_search_device: keys written by the traversal kernel, ncclAllGather in place, device merge);
    torch.distributed is used once, to hand NCCL's unique id from rank 0 to the other ranks."""

    def __init__(self, ix, rank, world, id_offset):
        import torch
        import torch.distributed as dist
        self.ix, self.rank, self.world, self.id_offset = ix, rank, world, int(id_offset)
        self._lib = lib = _lib.load()
        dev = torch.device("cuda", ix.device)
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            _lib.check(lib.ngtgpu_comm_get_unique_id(buf))
            idt = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
        if world > 1:
            idt = idt.to(dev) if dist.get_backend() == "nccl" else idt
            dist.broadcast(idt, 0)
        raw = bytes(idt.cpu().numpy().tobytes())
        h = C.c_void_p()
        _lib.check(lib.ngtgpu_comm_create(C.byref(h), raw, rank, world, ix.device))
        self._h = h

    def set_timing(self, enabled):
        _lib.check(self._lib.ngtgpu_comm_set_timing(self._h, int(enabled)))

    def pop_timing(self):
        """-> ({'search_ms', 'allgather_ms', 'merge_ms'} summed over the recorded calls, number of calls)"""
        ms = (C.c_double * 3)()
        n = C.c_uint64(0)
        _lib.check(self._lib.ngtgpu_comm_pop_timing(self._h, ms, C.byref(n)))
        return {"search_ms": ms[0], "allgather_ms": ms[1], "merge_ms": ms[2]}, int(n.value)

    def _out(self, nq, k, dev):
        import torch
        return (torch.zeros((nq, k), dtype=torch.int32, device=dev), torch.zeros((nq, k), dtype=torch.float32, device=dev),
                torch.zeros((nq,), dtype=torch.int32, device=dev))

    def _query(self, queries):
        import torch
        q = queries.contiguous()
        qt = _lib.OBJECT_UINT8 if q.dtype == torch.uint8 else _lib.OBJECT_FLOAT
        if qt == _lib.OBJECT_FLOAT and q.dtype != torch.float32:
            q = q.float()
        return q, qt

    def search(self, queries, k, epsilon, edge_size=-1, n_seeds=10, radius=-1.0):
        """queries: CUDA tensor [nq, dim], identical on every rank -> merged (global ids, dists, counts) on every rank."""
        import torch
        q, qt = self._query(queries)
        ids, dists, counts = self._out(q.shape[0], k, q.device)
        p = _lib.SearchParams(int(k), float(epsilon), float(radius), int(edge_size))
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(self._lib.ngtgpu_shard_search_device(self.ix._h, self._h, q.data_ptr(), qt, q.shape[0], C.byref(p), int(n_seeds),
                                                        self.id_offset, ids.data_ptr(), dists.data_ptr(), counts.data_ptr(), stream))
        return ids, dists, counts

    def linear_search(self, queries, k, radius=-1.0):
        import torch
        q, qt = self._query(queries)
        ids, dists, counts = self._out(q.shape[0], k, q.device)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(self._lib.ngtgpu_shard_linear_search_device(self.ix._h, self._h, q.data_ptr(), qt, q.shape[0], int(k), float(radius),
                                                               self.id_offset, ids.data_ptr(), dists.data_ptr(), counts.data_ptr(),
                                                               stream))
        return ids, dists, counts

    def close(self):
        if self._h:
            self._lib.ngtgpu_comm_destroy(self._h)
            self._h = None


class ShardedIndex:
    """One process, several GPUs (ngtgpu_sharded_*): host rows in, host results out. What ngt_open_index serves when
    NGTGPU_DEVICES lists several devices."""

    def __init__(self, object_type, distance_type, dimension, devices):
        self._lib = lib = _lib.load()
        self.object_type, self.distance_type, self.dimension = object_type, distance_type, int(dimension)
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        _lib.check(lib.ngtgpu_sharded_create(C.byref(h), arr, len(self.devices), object_type, distance_type, self.dimension))
        self._h = h

    def set_objects(self, rows, normalize=False):
        dt = np.uint8 if self.object_type == _lib.OBJECT_UINT8 else np.float32
        rows = np.ascontiguousarray(rows, dt)
        _lib.check(self._lib.ngtgpu_sharded_set_objects(self._h, rows.ctypes.data, rows.shape[0], int(normalize)))
        self.size = rows.shape[0]

    def build_onng(self, knn=64, outgoing=10, incoming=64, shortcut_reduction=True, edge_size_for_search=40, n_pivots=256):
        _lib.check(self._lib.ngtgpu_sharded_build_onng(self._h, int(knn), int(outgoing), int(incoming), int(shortcut_reduction),
                                                       int(edge_size_for_search), int(n_pivots)))

    def shard(self, g):
        """-> (raw ngtgpu_index handle, id offset, object count) of shard g"""
        ix, off, cnt = C.c_void_p(), C.c_uint64(0), C.c_uint64(0)
        _lib.check(self._lib.ngtgpu_sharded_shard(self._h, int(g), C.byref(ix), C.byref(off), C.byref(cnt)))
        return ix, int(off.value), int(cnt.value)

    def _queries(self, queries):
        q = np.asarray(queries)
        qt = _lib.OBJECT_UINT8 if q.dtype == np.uint8 else _lib.OBJECT_FLOAT
        return np.ascontiguousarray(q, np.uint8 if qt == _lib.OBJECT_UINT8 else np.float32), qt

    def search(self, queries, k, epsilon=0.1, radius=-1.0, edge_size=-1, n_seeds=10):
        q, qt = self._queries(queries)
        nq = q.shape[0]
        ids, dists, counts = np.zeros((nq, k), np.uint32), np.zeros((nq, k), np.float32), np.zeros(nq, np.uint32)
        p = _lib.SearchParams(int(k), float(epsilon), float(radius), int(edge_size))
        _lib.check(self._lib.ngtgpu_sharded_search(self._h, q.ctypes.data, qt, nq, C.byref(p), int(n_seeds), ids.ctypes.data,
                                                   dists.ctypes.data, counts.ctypes.data))
        return ids, dists, counts

    def linear_search(self, queries, k, radius=-1.0):
        q, qt = self._queries(queries)
        nq = q.shape[0]
        ids, dists, counts = np.zeros((nq, k), np.uint32), np.zeros((nq, k), np.float32), np.zeros(nq, np.uint32)
        _lib.check(self._lib.ngtgpu_sharded_linear_search(self._h, q.ctypes.data, qt, nq, int(k), float(radius), ids.ctypes.data,
                                                          dists.ctypes.data, counts.ctypes.data))
        return ids, dists, counts

    def last_timing(self):
        ms = (C.c_double * 5)()
        _lib.check(self._lib.ngtgpu_sharded_last_timing(self._h, ms))
        return dict(zip(("upload_broadcast_ms", "search_ms", "allgather_ms", "merge_ms", "download_ms"), [float(v) for v in ms]))

    def close(self):
        if self._h:
            self._lib.ngtgpu_sharded_destroy(self._h)
            self._h = None

"""Row-sharded search over the GPUs of one box.

Rank g owns the objects with global ids g*n_local+1 .. (g+1)*n_local as its own index (own graph, own seed
table; no cross-shard edges). Every rank gets the full query batch, searches its shard with the same
epsilon / k, and the per-shard top-k lists are exchanged with ONE all-gather (NCCL over NVLink) and merged per
query by (distance, id) -- the order of ObjectDistance, lib/NGT/Common.h:1946-1952 -- which is what a single
priority queue over the union keeps. The reference has no distribution; this is the
B200-native addition.

The exchange is written against torch.distributed only (backend nccl on GPUs, gloo in the CPU tests): lists
travel as int64 keys = ordered distance bits << 32 | global id, so one flat tensor per rank is gathered.
Packing and merging run in libngtgpu.so on the device; `merge_keys_host` is the same merge stated in numpy
for the CPU (gloo) tests of the plumbing.
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

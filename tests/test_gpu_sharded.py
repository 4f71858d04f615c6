"""Row-sharded search inside libngtgpu.so (csrc/shard.cu): keys written by the traversal kernel,
ONE ncclAllGather, device merge. Needs two visible GPUs (skipped otherwise): run with `gpurun --gpus 2`.
The bar: merged == search of the union, bit for bit (ids, float bits, counts) -- for the exhaustive scan against one index
holding every row, for the graph search against the host-side merge of the per-shard answers."""
import ctypes as C
import json
import os
import socket

import numpy as np
import pytest

import capi
from conftest import GOLDEN, ROOT
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


needs2 = pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")


def _rows(kind, n, seed):
    from ngt_b200 import synth
    x = synth.make("sift", n, seed)
    if kind == "f32":
        return x, po.FLOAT, po.L2
    if kind == "u8":
        return x.astype(np.uint8), po.UINT8, po.L2
    return np.packbits(x > 64.0, axis=1, bitorder="little").astype(np.uint8), po.UINT8, po.HAMMING


@needs2
@pytest.mark.parametrize("kind", ["u8", "ham", "f32"])
def test_one_process_sharded_handle_equals_union(eng, kind):
    from ngt_b200 import _lib, sharded
    n, nq, k = 30001, 257, 10          # an uneven split
    base, otype, dtype = _rows(kind, n, 31)
    qs = _rows(kind, nq, 32)[0]
    sh = sharded.ShardedIndex(otype, dtype, base.shape[1], [0, 1])
    sh.set_objects(base)
    sh.build_onng(knn=24, outgoing=8, incoming=24, shortcut_reduction=True, edge_size_for_search=40, n_pivots=128)
    whole = eng.GpuIndex(otype, dtype, base.shape[1])
    whole.set_objects(base)
    # exhaustive: merged shards == the one index over every row, bit for bit
    for radius in (-1.0, float(np.median(whole.linear_search(qs[:8], 3)[1][:, 2]))):
        ids, dists, counts = sh.linear_search(qs, k, radius)
        rids, rd, rc = whole.linear_search(qs, k, radius)
        assert (counts == rc).all() and (ids == rids).all() and (dists.view(np.uint32) == rd.view(np.uint32)).all()
    # graph search: merged == host merge of what each shard answers on its own (same graphs, same seeds)
    lib = _lib.load()
    per = []
    for g in range(2):
        h, off, cnt = sh.shard(g)
        assert off == (g * n) // 2 and cnt == ((g + 1) * n) // 2 - off
        i, d, c = np.zeros((nq, k), np.uint32), np.zeros((nq, k), np.float32), np.zeros(nq, np.uint32)
        p = _lib.SearchParams(k, 0.1, -1.0, -1)
        _lib.check(lib.ngtgpu_search(h, qs.ctypes.data, _lib.OBJECT_UINT8 if qs.dtype == np.uint8 else _lib.OBJECT_FLOAT, nq,
                                     C.byref(p), None, 10, i.ctypes.data, d.ctypes.data, c.ctypes.data, None))
        per.append(sharded.pack_keys_host(i, d, c, off))
    mids, md, mc = sharded.merge_keys_host(np.stack(per), k)
    ids, dists, counts = sh.search(qs, k, 0.1, n_seeds=10)
    assert (counts == mc).all() and (ids == mids).all() and (dists.view(np.uint32) == md.view(np.uint32)).all()
    t = sh.last_timing()
    assert t["search_ms"] > 0 and t["allgather_ms"] > 0
    # and it is a good search: recall against the exhaustive answer
    # (recall as lib/NGT/Optimizer.h:496-507 counts it: an id of the ground truth, or a distance within its k-th)
    gt, gt_d, _ = whole.linear_search(qs, k)
    hit = np.mean([np.mean([(ids[q, r] in gt[q]) or dists[q, r] <= gt_d[q, -1] for r in range(k)]) for q in range(nq)])
    assert hit >= 0.9, hit
    assert sh.search(qs, 0, 0.1)[2].sum() == 0
    whole.close()
    sh.close()


@needs2
def test_c_api_sharded_open(sift5k):
    """ngt_open_index_sharded / NGTGPU_DEVICES: the same C-API calls, answered by two GPUs. The exhaustive answers equal
    the reference's on the same index (tests/golden/idx200_answers.json) bit for bit; mutation is refused."""
    from ngt_b200 import _lib
    lib = capi.bind(_lib.SO_PATH)
    lib.ngt_open_index_sharded.restype = C.c_void_p
    lib.ngt_open_index_sharded.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_void_p]
    ans = json.load(open(os.path.join(GOLDEN, "idx200_answers.json")))
    err = lib.ngt_create_error_object()
    devs = (C.c_int * 2)(0, 1)
    path = os.path.join(GOLDEN, "idx200").encode()
    for how in ("call", "env"):
        if how == "call":
            ix = lib.ngt_open_index_sharded(path, devs, 2, err)
        else:
            os.environ["NGTGPU_DEVICES"] = "0,1"
            try:
                ix = lib.ngt_open_index(path, err)
            finally:
                del os.environ["NGTGPU_DEVICES"]
        assert ix, lib.ngt_get_error_string(err)
        qs = np.ascontiguousarray(sift5k["queries"].astype(np.float32))
        r = lib.ngt_create_empty_results(err)
        for qi in range(3):
            q = np.ascontiguousarray(qs[qi])
            assert lib.ngt_linear_search_index_as_float(ix, capi.fptr(q), 128, 5, r, err), lib.ngt_get_error_string(err)
            assert capi.results_of(lib, r, err) == [(a[0], float(np.float32(a[1]))) for a in ans["linear"][qi]]
            assert lib.ngt_search_index_as_float(ix, capi.fptr(q), 128, 5, 0.3, -1.0, r, err), lib.ngt_get_error_string(err)
            assert capi.results_of(lib, r, err) == [(a[0], float(np.float32(a[1]))) for a in ans["graph"][qi]]
        ids, ds, cnt = np.zeros((3, 5), np.uint32), np.zeros((3, 5), np.float32), np.zeros(3, np.uint32)
        u32 = C.POINTER(C.c_uint32)
        assert lib.ngt_batch_search_index_as_float(ix, capi.fptr(qs), 3, 128, 5, 0.3, -1.0, -1, ids.ctypes.data_as(u32), capi.fptr(ds),
                                                   cnt.ctypes.data_as(u32), err)
        assert ids.tolist() == [[a[0] for a in ans["graph"][i]] for i in range(3)]
        assert lib.ngt_batch_append_index(ix, capi.fptr(qs), 3, err) is False and b"read-only" in lib.ngt_get_error_string(err)
        assert lib.ngt_remove_index(ix, 3, err) is False
        lib.ngt_destroy_results(r)
        lib.ngt_close_index(ix)
    lib.ngt_destroy_error_object(err)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ngt_b200 import _lib, build, engine, sharded
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n, nq, k = 24000, 300, 10
    base = _rows("u8", n, 41)[0]
    qs = torch.from_numpy(_rows("u8", nq, 42)[0]).to(dev)
    n_local = n // world
    mine = base[rank * n_local:(rank + 1) * n_local]
    ix = engine.GpuIndex(po.UINT8, po.L2, 128, device=rank)
    ix.set_objects(mine)
    ids, dists, counts = build.knn_graph(ix, 16)
    rp, col, _ = build.reconstruct_graph(ids, dists, counts, 8, 16)
    ix.set_graph(rp, col)
    ix.set_search_property(0, 30, 20)
    ix.build_seed_table(128, 1)
    S = sharded.LibShardedSearcher(ix, rank, world, rank * n_local)
    S.set_timing(True)
    lin = [t.cpu().numpy() for t in S.linear_search(qs, k)]
    gs = [t.cpu().numpy() for t in S.search(qs, k, 0.1, edge_size=-1, n_seeds=10)]
    own = ix.search(qs, k, 0.1, edge_size=-1, n_seeds=10)
    ms, calls = S.pop_timing()
    assert calls == 2 and ms["search_ms"] > 0
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lin_ids=lin[0], lin_d=lin[1], lin_c=lin[2], g_ids=gs[0], g_d=gs[1], g_c=gs[2],
             own_ids=own[0].cpu().numpy(), own_d=own[1].cpu().numpy(), own_c=own[2].cpu().numpy())
    S.close()
    ix.close()
    dist.barrier()
    dist.destroy_process_group()


@needs2
def test_one_process_per_gpu_nccl_allgather_in_the_library(eng, tmp_path):
    """Two ranks (torch.distributed only carries NCCL's id): ngtgpu_shard_*_device on each; every rank holds the merged
    answer; exhaustive == union bit for bit, graph search == host merge of the two shards' own answers."""
    import torch.multiprocessing as mp
    from ngt_b200 import sharded
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    n, nq, k = 24000, 300, 10
    base, qs = _rows("u8", n, 41)[0], _rows("u8", nq, 42)[0]
    d2 = ((base[None, :, :].astype(np.int32) - qs[:, None, :].astype(np.int32)) ** 2).sum(-1)
    order = np.lexsort((np.broadcast_to(np.arange(n), d2.shape), d2), axis=1)[:, :k]
    exact_d = np.sqrt(np.take_along_axis(d2, order, 1).astype(np.float64)).astype(np.float32)
    for r in range(world):
        assert (z[r]["lin_c"] == k).all() and (z[r]["lin_ids"] == order + 1).all()
        assert (z[r]["lin_d"].view(np.uint32) == exact_d.view(np.uint32)).all()
    keys = np.stack([sharded.pack_keys_host(z[r]["own_ids"], z[r]["own_d"], z[r]["own_c"], r * (n // world)) for r in range(world)])
    mids, md, mc = sharded.merge_keys_host(keys, k)
    for r in range(world):
        assert (z[r]["g_c"] == mc).all() and (z[r]["g_ids"].astype(np.uint32) == mids).all()
        assert (z[r]["g_d"].view(np.uint32) == md.view(np.uint32)).all()

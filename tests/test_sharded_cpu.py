"""World-size-2 test of the row-sharded search plumbing on CPU (gloo): per-shard exhaustive top-k from the
oracle, key packing, ONE all-gather, merge by (distance, id) == the oracle on the union, bit for bit.
The device kernels that do packing/merging on the GPU are checked against the same host functions in
tests/test_gpu_parity.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ngt_b200 import sharded
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = dict(np.load(os.path.join(GOLDEN, "sift5k.npz")))
    data, qs = z["data"], z["queries"]
    n_local = data.shape[0] // world
    mine = data[rank * n_local:(rank + 1) * n_local]
    port_o = po.Port()
    k = 20
    ids, dists, counts = port_o.linear_search(po.L2, po.UINT8, po.pad_objects(mine, po.UINT8), po.pad_queries(qs, po.UINT8), k)
    keys = sharded.pack_keys_host(ids, dists, counts, rank * n_local)
    gathered = sharded.all_gather_keys(torch.from_numpy(keys.view(np.int64)), world)
    mids, mdists, mcounts = sharded.merge_keys_host(gathered.numpy().view(np.uint64), k)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ids=mids, dists=mdists, counts=mcounts)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_merge_equals_union_gloo(tmp_path, port, sift5k):
    from oracle import pyoracle as po
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    k = 20
    ref_ids, ref_d, ref_c = port.linear_search(po.L2, po.UINT8, po.pad_objects(sift5k["data"], po.UINT8),
                                               po.pad_queries(sift5k["queries"], po.UINT8), k)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert (z["counts"] == ref_c).all()
        assert (z["ids"] == ref_ids).all()
        assert (z["dists"].view(np.uint32) == ref_d.view(np.uint32)).all()


def test_key_order_is_object_distance_order():
    from ngt_b200 import sharded
    rng = np.random.default_rng(1)
    d = np.concatenate([rng.random(50).astype(np.float32) * 100, np.array([0.0, -0.0, 1.5, 1.5, 3e38], np.float32)])
    ids = rng.permutation(d.size).astype(np.uint32) + 1
    keys = sharded.pack_keys_host(ids[None, :], d[None, :], np.array([d.size]), 0)[0]
    order = np.argsort(keys, kind="stable")
    ref = np.lexsort((ids, d + np.float32(0.0)))        # (distance, id), Common.h:1946-1952
    assert (order == ref).all()
    i2, d2, v = sharded.unpack_keys_host(keys)
    assert v.all() and (i2 == ids).all() and (d2.view(np.uint32) == (d + np.float32(0)).view(np.uint32)).all()

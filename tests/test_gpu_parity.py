"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors produced by the
unmodified reference (tests/golden/*.npz) and the C restatement (oracle/) on seeded inputs.

Bars (tests/parity.py): bit-exact ids + float bits for uint8 L2, Hamming and integer-valued float data;
1e-6 relative on distances and near-tie id swaps only for general float data.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from parity import (EDGE_GRID, EPS_GRID, FLOAT_CASES, INTEGER_EXACT, NORMALIZED, assert_bit_exact,
                    assert_float_parity, grid_key)

pytestmark = pytest.mark.gpu


def _gpu_index(eng, otype, dtype, objects, row_ptr=None, col=None, prop=None, normalize=False):
    ix = eng.GpuIndex(otype, dtype, objects.shape[1])
    ix.set_objects(objects, normalize=normalize)
    if row_ptr is not None:
        ix.set_graph(np.asarray(row_ptr, np.uint64), col)
    if prop is not None:
        ix.set_search_property(*[int(v) for v in prop])
    return ix


# ---------------------------------------------------------------------------------------------------
# reference golden vectors
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,otype", [("u8", po.UINT8), ("f32", po.FLOAT)])
def test_sift5k_golden(eng, sift5k, tag, otype):
    data = sift5k["data"]
    qs = sift5k["queries"].astype(np.float32)   # every ngt_search_index* entry point takes numbers, Capi.cpp:377-406
    ix = _gpu_index(eng, otype, po.L2, data, sift5k["row_ptr"], sift5k["col"], sift5k["prop"])
    ids, dists, counts = ix.linear_search(qs, 20)
    assert_bit_exact(ids, dists, counts, sift5k[tag + "_lin_ids"], sift5k[tag + "_lin_dists"], what="linear")
    for eps in EPS_GRID:
        for e in EDGE_GRID:
            key = grid_key(tag, eps, e)
            ids, dists, counts, stats = ix.search(qs, 20, eps, edge_size=e, seeds=sift5k["seeds"], with_stats=True)
            assert_bit_exact(ids, dists, counts, sift5k[key + "_ids"], sift5k[key + "_dists"],
                             sift5k[key + "_counts"], what=key)
            ref = sift5k[key + "_stats"]   # writable path: counts after the seeds, Graph.cpp:592,604
            assert (stats[:, 0].astype(np.int64) - sift5k["seeds"].shape[1] == ref[:, 0]).all(), key
            assert (stats[:, 1] == ref[:, 1]).all(), key
    ix.close()


def test_readme_known_answer(eng, sift5k):
    """bin/ngt/README.md:254-323 with the seeds the reference's DVP-tree produced."""
    import json
    import os
    from conftest import GOLDEN
    kat = json.load(open(os.path.join(GOLDEN, "readme_kat.json")))
    ix = _gpu_index(eng, po.UINT8, po.L2, sift5k["data"], sift5k["row_ptr"], sift5k["col"], sift5k["prop"])
    ids, dists, counts = ix.search(sift5k["queries"], 20, 0.1, seeds=sift5k["tree_seeds"])
    for q in range(3):
        assert [[int(ids[q, i]), "%g" % float(dists[q, i])] for i in range(20)] == kat[q]
    ix.close()


def _case(z, tag):
    data_tag = "f32l2" if tag == "onng" else tag
    ot, dt, dim, es, base, rate = [int(v) for v in z[tag + "_meta"]]
    return dict(otype=ot, dtype=dt, prop=(es, base, rate), objects=z[data_tag + "_objects"],
                queries=z[data_tag + "_queries"], row_ptr=z[tag + "_row_ptr"], col=z[tag + "_col"], seeds=z[tag + "_seeds"])


@pytest.mark.parametrize("tag", INTEGER_EXACT + FLOAT_CASES)
def test_synth_golden(eng, synth_golden, tag):
    z = synth_golden
    c = _case(z, tag)
    exact = tag in INTEGER_EXACT
    check = assert_bit_exact if exact else assert_float_parity
    # stored rows are what the reference holds (already normalised for Normalized* types)
    ix = _gpu_index(eng, c["otype"], c["dtype"], c["objects"], c["row_ptr"], c["col"], c["prop"], normalize=False)
    qs = c["queries"].astype(np.float32)
    if tag != "onng":
        ids, dists, counts = ix.linear_search(qs, 10)
        check(ids, dists, counts, z[tag + "_lin_ids"], z[tag + "_lin_dists"], what=tag + " linear")
        rad = float(z[tag + "_linr_radius"][0])
        ids, dists, counts = ix.linear_search(qs, 10, radius=rad)
        if exact:
            assert_bit_exact(ids, dists, counts, z[tag + "_linr_ids"], z[tag + "_linr_dists"], z[tag + "_linr_counts"],
                             what=tag + " linear radius")
    for eps in EPS_GRID:
        for e in EDGE_GRID:
            key = grid_key(tag, eps, e)
            ids, dists, counts, stats = ix.search(qs, 10, eps, edge_size=e, seeds=c["seeds"], with_stats=True)
            check(ids, dists, counts, z[key + "_ids"], z[key + "_dists"], z[key + "_counts"], what=key)
            if exact:
                assert (stats[:, 1] == z[key + "_stats"][:, 1]).all(), key
    ix.close()


def test_normalization_matches_reference(eng, synth_golden):
    """ObjectSpace.h:251-266 on the device: raw rows in, the reference's stored rows out (1e-6)."""
    raw = synth_golden["glove_l2_objects"]
    stored = synth_golden["glove_ncos_objects"]
    ix = eng.GpuIndex(po.FLOAT, po.NORMALIZED_COSINE, raw.shape[1])
    ix.set_objects(raw)          # normalises by default for Normalized* types
    got = np.stack([ix.get_object(i + 1) for i in range(0, raw.shape[0], 97)])
    assert np.abs(got - stored[::97]).max() <= 1e-6
    with pytest.raises(eng.NgtGpuError):
        ix.linear_search(np.zeros((1, raw.shape[1]), np.float32), 5)     # zero query vector: the reference throws
    ix.close()
    ix2 = eng.GpuIndex(po.FLOAT, po.NORMALIZED_COSINE, raw.shape[1])
    bad = raw.copy()
    bad[3] = 0
    with pytest.raises(eng.NgtGpuError):
        ix2.set_objects(bad)
    ix2.close()


# ---------------------------------------------------------------------------------------------------
# seeded inputs against the C restatement
# ---------------------------------------------------------------------------------------------------
def _knn_csr(ids, counts):
    """adjacency lists from exhaustive results (already ascending by (distance,id))."""
    n = ids.shape[0]
    row_ptr = np.zeros(n + 2, np.uint64)
    row_ptr[2:] = np.cumsum(counts.astype(np.uint64))
    col = np.concatenate([ids[i, :counts[i]] for i in range(n)]).astype(np.uint32)
    return row_ptr, col


@pytest.mark.parametrize("kind", ["u8l2", "ham", "f32int", "f32cos"])
def test_seeded_against_port(eng, port, kind):
    from ngt_b200 import synth
    rng = np.random.default_rng(7)
    n, nq, k = 6000, 64, 10
    base = synth.make("sift", n, 1)
    qs = synth.make("sift", nq, 2)
    if kind == "u8l2":
        otype, dtype, objs, q = po.UINT8, po.L2, base.astype(np.uint8), qs
    elif kind == "ham":
        otype, dtype = po.UINT8, po.HAMMING
        objs = synth.hamming_from(base, 64.0)
        q = synth.hamming_from(qs, 64.0).astype(np.float32)
    elif kind == "f32int":
        otype, dtype, objs, q = po.FLOAT, po.L2, base, qs
    else:
        otype, dtype = po.FLOAT, po.COSINE
        objs = (base - 64.0).astype(np.float32) / 40.0
        q = (qs - 64.0).astype(np.float32) / 40.0
    ix = eng.GpuIndex(otype, dtype, objs.shape[1])
    ix.set_objects(objs)
    pobj = po.pad_objects(objs, otype)
    pq = po.pad_queries(q, otype)
    check = assert_float_parity if kind == "f32cos" else assert_bit_exact
    # exhaustive search, with and without a radius, k beyond one warp, k > n handled below
    for kk in (1, k, 40):
        ids, dists, counts = ix.linear_search(q, kk)
        rids, rdists, rcounts = port.linear_search(dtype, otype, pobj, pq, kk)
        check(ids, dists, counts, rids, rdists, rcounts, what="%s linear k=%d" % (kind, kk))
    rad = float(rdists[0, 20])
    ids, dists, counts = ix.linear_search(q, k, radius=rad)
    rids, rdists2, rcounts = port.linear_search(dtype, otype, pobj, pq, k, radius=rad)
    check(ids, dists, counts, rids, rdists2, rcounts, what=kind + " linear radius")
    # a kNN graph made by the engine itself (self included: harmless), searched by both
    gids, _, gcounts = ix.linear_search(objs.astype(np.float32), 12)
    row_ptr, col = _knn_csr(gids, gcounts)
    ix.set_graph(row_ptr, col)
    seeds = np.stack([rng.choice(n, 8, replace=False) + 1 for _ in range(nq)]).astype(np.uint32)
    for eps, cap, kk in ((0.1, 2 ** 31 - 1, k), (0.25, 7, k), (0.0, 2 ** 31 - 1, 40), (-0.05, 2 ** 31 - 1, k)):
        es = 0 if cap == 2 ** 31 - 1 else cap
        ids, dists, counts, stats = ix.search(q, kk, eps, edge_size=es, seeds=seeds, with_stats=True)
        rids, rdists, rcounts, rstats = port.graph_search(dtype, otype, pobj, row_ptr, col, pq, seeds, kk, eps,
                                                          edge_size=cap)
        what = "%s graph eps=%g cap=%d k=%d" % (kind, eps, cap, kk)
        check(ids, dists, counts, rids, rdists, rcounts, what=what)
        if kind != "f32cos":
            assert (stats.astype(np.uint64) == rstats).all(), what
    # a tiny on-chip working set forces every query to overflow: the second shared-memory tier, and the
    # HBM tier (exact bitmap + queue in global memory), give the same answers
    rids, rdists, rcounts, _ = port.graph_search(dtype, otype, pobj, row_ptr, col, pq, seeds, k, 0.1)
    for tiers in (2, 1):
        ix.set_search_workspace(hash_bits=8, queue_cap=64, onchip_tiers=tiers)
        ids, dists, counts = ix.search(q, k, 0.1, edge_size=0, seeds=seeds)
        assert ix.last_overflows > 0
        check(ids, dists, counts, rids, rdists, rcounts, what="%s overflow, %d on-chip tiers" % (kind, tiers))
    ix.close()


@pytest.mark.parametrize("kind", ["f32l2_128", "f32cos_100", "f32l2_48", "u8l2_128", "f32ncos_32", "u8l2_256", "u8ham_128"])
def test_fast_kernel_equals_general_kernel(eng, port, kind):
    """The lean first-tier kernel (search_fast.cuh) and the general one (search.cuh) restate the same loop
    (Graph.cpp:398-495): ids, distance bits, counts and work counters are identical, for every row width class
    (1, 2, 4 chunks per lane), with repeated seeds, small and large edge caps, k up to 32, and when queries
    overflow the first tier. Integer-valued kinds are also compared with the C restatement of the reference."""
    from ngt_b200 import synth
    rng = np.random.default_rng(11)
    name, dim = kind.split("_")
    dim = int(dim)
    n, nq = 20000, 300
    base = synth.make("sift", n, 1)[:, :min(dim, 128)]
    qs = synth.make("sift", nq, 2)[:, :min(dim, 128)]
    if dim > 128:
        base, qs = np.concatenate([base, base[:, ::-1]], 1), np.concatenate([qs, qs[:, ::-1]], 1)
    normalize = False
    if name == "f32l2":
        otype, dtype, objs, q = po.FLOAT, po.L2, base, qs
    elif name == "u8l2":
        otype, dtype, objs, q = po.UINT8, po.L2, base.astype(np.uint8), qs
    elif name == "u8ham":
        otype, dtype = po.UINT8, po.HAMMING   # 128 bits = one 16-byte chunk per object
        objs, q = synth.hamming_from(base, 64.0), synth.hamming_from(qs, 64.0).astype(np.float32)
    elif name == "f32cos":
        otype, dtype = po.FLOAT, po.COSINE
        objs, q = (base - 64.0).astype(np.float32) / 40.0, (qs - 64.0).astype(np.float32) / 40.0
    else:
        otype, dtype, normalize = po.FLOAT, po.NORMALIZED_COSINE, True
        objs, q = (base - 64.0).astype(np.float32) / 40.0, (qs - 64.0).astype(np.float32) / 40.0
    ix = eng.GpuIndex(otype, dtype, objs.shape[1])
    ix.set_objects(objs, normalize=normalize)
    gids, _, gcounts = ix.linear_search(objs.astype(np.float32), 25)
    row_ptr, col = _knn_csr(gids, gcounts)
    ix.set_graph(row_ptr, col)
    seeds = np.stack([rng.choice(n, 10, replace=False) + 1 for _ in range(nq)]).astype(np.uint32)
    seeds_rep = seeds.copy()
    seeds_rep[::3, 5] = seeds_rep[::3, 1]   # a repeated seed is evaluated once by both kernels
    exact = name in ("f32l2", "u8l2", "u8ham")
    for eps, cap, kk in ((0.1, 16, 10), (0.3, 100, 10), (0.0, 24, 32), (0.2, 128, 1)):
        out = {}
        for fast in (True, False):
            ix.set_fast_kernel(fast)
            out[fast] = ix.search(q, kk, eps, edge_size=cap, seeds=seeds, with_stats=True)
        what = "%s eps=%g cap=%d k=%d" % (kind, eps, cap, kk)
        for a, b in zip(out[True], out[False]):
            a, b = np.asarray(a), np.asarray(b)
            assert (a.view(np.uint32) == b.view(np.uint32)).all(), what
        if cap <= 64 and objs.shape[1] * objs.dtype.itemsize <= 128:
            # the one-warp-per-query shape of the lean kernel (rounds of <= 64 edges in two filter passes): same everything
            ix.set_fast_kernel(True)
            for w in (1, 4):
                ix.set_fast_shape(w, 0)
                for sd in (seeds, seeds_rep):
                    one = ix.search(q, kk, eps, edge_size=cap, seeds=sd, with_stats=True)
                    two = out[False] if sd is seeds else None
                    if two is None:
                        ix.set_fast_kernel(False)
                        two = ix.search(q, kk, eps, edge_size=cap, seeds=sd, with_stats=True)
                        ix.set_fast_kernel(True)
                    for x, y in zip(one, two):
                        assert (np.asarray(x).view(np.uint32) == np.asarray(y).view(np.uint32)).all(), what + " %d warp(s) per query" % w
            ix.set_fast_shape(0, 0)
        if cap > 64:
            # the two-warp shape asked for on rounds of 65..128 edges (two edges per thread)
            ix.set_fast_kernel(True)
            ix.set_fast_shape(2, 0)
            two = ix.search(q, kk, eps, edge_size=cap, seeds=seeds, with_stats=True)
            ix.set_fast_shape(0, 0)
            for x, y in zip(two, out[False]):
                assert (np.asarray(x).view(np.uint32) == np.asarray(y).view(np.uint32)).all(), what + " two warps, two passes"
        rep = {}
        for fast in (True, False):
            ix.set_fast_kernel(fast)
            rep[fast] = ix.search(q, kk, eps, edge_size=cap, seeds=seeds_rep, with_stats=True)
        for a, b in zip(rep[True], rep[False]):
            assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all(), what + " repeated seeds"
        if exact:
            pobj, pq = po.pad_objects(objs, otype), po.pad_queries(q, otype)
            rids, rdists, rcounts, rstats = port.graph_search(dtype, otype, pobj, row_ptr, col, pq, seeds, kk, eps,
                                                              edge_size=cap)
            assert_bit_exact(out[True][0], out[True][1], out[True][2], rids, rdists, rcounts, what=what)
            assert (out[True][3].astype(np.uint64) == rstats).all(), what
    # queries that outgrow the first tier leave the lean kernel for the general one's later tiers
    ix.set_fast_kernel(False)
    ref = ix.search(q, 10, 0.3, edge_size=100, seeds=seeds)
    ix.set_fast_kernel(True)
    ix.set_search_workspace(hash_bits=9, queue_cap=64)
    got = ix.search(q, 10, 0.3, edge_size=100, seeds=seeds)
    assert ix.last_overflows > 0
    for a, b in zip(got, ref):
        assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all(), kind + " overflow"
    ix.close()


@pytest.mark.parametrize("kind", ["f32l2_128", "f32cos_100", "f32l2_48", "u8l2_128", "f32ncos_32", "u8ham_256"])
def test_fast_kernel_wide_result_lists(eng, port, kind):
    """search_fast_kernel<.., KL = 4>: result lists of 33..128 keys, four per lane of the control warp -- the searches of
    the construction loop and of refineANNG, whose k is the edge count (Index.h:815-837, GraphReconstructor.h:852).
    Ids, distance bits, counts and work counters equal the general kernel's (sorted array in shared memory) for every
    row width class; integer-valued kinds also equal the C restatement of the reference. Covered: k not a multiple of
    four, lists that stay shorter than k (radius), repeated seeds, queries that overflow the first tier."""
    from ngt_b200 import synth
    rng = np.random.default_rng(13)
    name, dim = kind.split("_")
    dim = int(dim)
    n, nq = 20000, 200
    base = synth.make("sift", n, 1)[:, :min(dim, 128)]
    qs = synth.make("sift", nq, 2)[:, :min(dim, 128)]
    normalize = False
    if name == "f32l2":
        otype, dtype, objs, q = po.FLOAT, po.L2, base, qs
    elif name == "u8l2":
        otype, dtype, objs, q = po.UINT8, po.L2, base.astype(np.uint8), qs
    elif name == "u8ham":
        otype, dtype = po.UINT8, po.HAMMING   # 1024 bits = eight 16-byte chunks per object
        wide, wq = np.concatenate([base, base[:, ::-1]], 1), np.concatenate([qs, qs[:, ::-1]], 1)
        objs = np.concatenate([synth.hamming_from(wide, t) for t in (40.0, 56.0, 72.0, 88.0)], 1)
        q = np.concatenate([synth.hamming_from(wq, t) for t in (40.0, 56.0, 72.0, 88.0)], 1).astype(np.float32)
    elif name == "f32cos":
        otype, dtype = po.FLOAT, po.COSINE
        objs, q = (base - 64.0).astype(np.float32) / 40.0, (qs - 64.0).astype(np.float32) / 40.0
    else:
        otype, dtype, normalize = po.FLOAT, po.NORMALIZED_COSINE, True
        objs, q = (base - 64.0).astype(np.float32) / 40.0, (qs - 64.0).astype(np.float32) / 40.0
    ix = eng.GpuIndex(otype, dtype, objs.shape[1])
    ix.set_objects(objs, normalize=normalize)
    gids, gd, gcounts = ix.linear_search(objs.astype(np.float32), 25)
    row_ptr, col = _knn_csr(gids, gcounts)
    ix.set_graph(row_ptr, col)
    seeds = np.stack([rng.choice(n, 10, replace=False) + 1 for _ in range(nq)]).astype(np.uint32)
    seeds_rep = seeds.copy()
    seeds_rep[::3, 5] = seeds_rep[::3, 1]
    exact = name in ("f32l2", "u8l2", "u8ham")
    rad = float(np.median(gd[:, 20]))   # lists that end below k: about twenty objects lie this close to a base row
    for eps, cap, kk, radius in ((0.1, 16, 33, -1.0), (0.1, 100, 64, -1.0), (0.0, 24, 100, -1.0), (0.2, 128, 128, -1.0),
                                 (0.1, 40, 50, rad), (0.1, 30, 127, -1.0)):
        what = "%s eps=%g cap=%d k=%d radius=%g" % (kind, eps, cap, kk, radius)
        for sd in (seeds_rep, seeds):   # (a repeated seed is evaluated once by both kernels)
            out = {}
            for fast in (True, False):
                ix.set_fast_kernel(fast)
                out[fast] = ix.search(q, kk, eps, radius=radius, edge_size=cap, seeds=sd, with_stats=True)
            for a, b in zip(out[True], out[False]):
                assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all(), what
        if radius >= 0:
            assert (np.asarray(out[True][2]) < kk).any(), what + ": no short list in this case"
        if exact:
            pobj, pq = po.pad_objects(objs, otype), po.pad_queries(q, otype)
            rids, rdists, rcounts, rstats = port.graph_search(dtype, otype, pobj, row_ptr, col, pq, seeds, kk, eps,
                                                              edge_size=cap, radius=radius)
            assert_bit_exact(out[True][0], out[True][1], out[True][2], rids, rdists, rcounts, what=what)
            assert (out[True][3].astype(np.uint64) == rstats).all(), what
    # queries that outgrow the first tier: the second tier is the same kernel with larger slabs, then the general one
    ix.set_fast_kernel(False)
    ref = ix.search(q, 64, 0.3, edge_size=100, seeds=seeds)
    ix.set_fast_kernel(True)
    for tiers in (2, 1):
        ix.set_search_workspace(hash_bits=9, queue_cap=64, onchip_tiers=tiers)
        got = ix.search(q, 64, 0.3, edge_size=100, seeds=seeds)
        assert ix.last_overflows > 0
        for a, b in zip(got, ref):
            assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all(), kind + " overflow, %d tiers" % tiers
    ix.close()


@pytest.mark.parametrize("kind", ["u8l2_128", "f32l2_32", "u8l2_128_wide"])
def test_one_warp_per_query_shape(eng, port, kind):
    """search_fast_kernel<.., W = 1>: rounds of 33..64 edges and seed lists of 40 (two filter passes per round, a seed id
    repeated across the passes), the back of the unchecked set in its global slab: identical to the general kernel and,
    being integer-valued, to the C restatement of the reference -- ids, distance bits, counts, work counters."""
    from ngt_b200 import synth
    rng = np.random.default_rng(5)
    # "wide": rounds of 65..100 edges and 80 seeds -- the two-warp shape with two edges per thread (asked for), against four warps
    wide = kind.endswith("_wide")
    name, dim = kind.split("_")[:2]
    dim = int(dim)
    n, nq = 20000, 400
    base, qs = synth.make("sift", n, 1)[:, :dim], synth.make("sift", nq, 2)[:, :dim]
    otype = po.UINT8 if name == "u8l2" else po.FLOAT
    objs = base.astype(np.uint8) if name == "u8l2" else base
    ix = eng.GpuIndex(otype, po.L2, dim)
    ix.set_objects(objs)
    gids, _, gcounts = ix.linear_search(objs.astype(np.float32), 100 if wide else 60)
    row_ptr, col = _knn_csr(gids, gcounts)
    ix.set_graph(row_ptr, col)
    ns = 80 if wide else 40
    seeds = np.stack([rng.choice(n, ns, replace=False) + 1 for _ in range(nq)]).astype(np.uint32)
    seeds_rep = seeds.copy()
    seeds_rep[::3, ns - 5] = seeds_rep[::3, 1]   # evaluated once (the engine's contract for repeated seeds)
    pobj, pq = po.pad_objects(objs, otype), po.pad_queries(qs, otype)
    for eps, cap, kk in (((0.1, 100, 10), (0.2, 96, 32), (0.0, 65, 1)) if wide else ((0.1, 64, 10), (0.25, 48, 32), (0.0, 33, 1))):
        what = "%s eps=%g cap=%d k=%d" % (kind, eps, cap, kk)
        for sd in (seeds_rep, seeds):
            ix.set_fast_kernel(False)
            ref = ix.search(qs, kk, eps, edge_size=cap, seeds=sd, with_stats=True)
            ix.set_fast_kernel(True)
            for w in ((4, 2) if wide else (2, 1)):
                ix.set_fast_shape(w, 0)
                got = ix.search(qs, kk, eps, edge_size=cap, seeds=sd, with_stats=True)
                for x, y in zip(got, ref):
                    assert (np.asarray(x).view(np.uint32) == np.asarray(y).view(np.uint32)).all(), what + " %d warp(s)" % w
        rids, rdists, rcounts, rstats = port.graph_search(po.L2, otype, pobj, row_ptr, col, pq, seeds, kk, eps, edge_size=cap)
        assert_bit_exact(got[0], got[1], got[2], rids, rdists, rcounts, what=what)
        assert (got[3].astype(np.uint64) == rstats).all(), what
    # a small slab and queue: the queries that outgrow the one-warp tier are finished by the later tiers
    ix.set_fast_shape(2 if wide else 1, 0)
    ix.set_search_workspace(hash_bits=9, queue_cap=64)
    got = ix.search(qs, 10, 0.25, edge_size=48, seeds=seeds)
    assert ix.last_overflows > 0
    ix.set_fast_kernel(False)
    ix.set_search_workspace(hash_bits=14, queue_cap=512)
    ref = ix.search(qs, 10, 0.25, edge_size=48, seeds=seeds)
    for x, y in zip(got, ref):
        assert (np.asarray(x).view(np.uint32) == np.asarray(y).view(np.uint32)).all(), kind + " overflow"
    ix.close()


def test_edge_cases(eng, port):
    rng = np.random.default_rng(3)
    objs = rng.integers(0, 256, (50, 24)).astype(np.uint8)       # ragged dimension: padded to 32
    q = rng.integers(0, 256, (5, 24)).astype(np.uint8)
    ix = eng.GpuIndex(po.UINT8, po.L2, 24)
    ix.set_objects(objs)
    assert ix.padded_dimension == 32 and ix.size == 50
    pobj, pq = po.pad_objects(objs, po.UINT8), po.pad_queries(q, po.UINT8)
    # k larger than the repository: everything comes back, ascending
    ids, dists, counts = ix.linear_search(q, 64)
    rids, rdists, rcounts = port.linear_search(po.L2, po.UINT8, pobj, pq, 64)
    assert (counts == 50).all()
    assert_bit_exact(ids, dists, counts, rids, rdists, rcounts, what="k > n")
    # size 0 returns nothing (Index.h:1141-1144); an empty batch is a no-op
    ids, dists, counts = ix.linear_search(q, 0)
    assert ids.shape == (5, 0) and (counts == 0).all()
    ids, dists, counts = ix.linear_search(np.zeros((0, 24), np.uint8), 3)
    assert ids.shape[0] == 0
    # removed objects are skipped (ObjectSpaceRepository.h:485)
    removed = np.array([1, 7, 50], np.uint32)
    ix.set_removed(removed)
    valid = np.ones(51, np.uint8)
    valid[0] = 0
    valid[removed] = 0
    ids, dists, counts = ix.linear_search(q, 10)
    rids, rdists, rcounts = port.linear_search(po.L2, po.UINT8, pobj, pq, 10, valid=valid)
    assert_bit_exact(ids, dists, counts, rids, rdists, rcounts, what="removed")
    assert not np.isin(ids, removed).any()
    # graph search without a graph / without seeds fails loudly
    with pytest.raises(eng.NgtGpuError):
        ix.search(q, 5, 0.1, seeds=np.ones((5, 2), np.uint32))
    # bad edge-size mode: the reference's message (Graph.h:687-689)
    ix.set_removed(np.zeros(0, np.uint32))
    row_ptr = np.zeros(52, np.uint64)
    ix.set_graph(row_ptr, np.zeros(0, np.uint32))
    with pytest.raises(eng.NgtGpuError, match="Invalid edge size parameters"):
        ix.search(q, 5, 0.1, edge_size=-3, seeds=np.ones((5, 2), np.uint32))
    # a graph with no edges returns the seeds that fit
    ids, dists, counts = ix.search(q, 5, 0.1, seeds=np.tile(np.array([[3, 9]], np.uint32), (5, 1)))
    rids, rdists, rcounts, _ = port.graph_search(po.L2, po.UINT8, pobj, row_ptr, np.zeros(1, np.uint32), pq,
                                                 np.tile(np.array([[3, 9]], np.uint32), (5, 1)), 5, 0.1)
    assert_bit_exact(ids, dists, counts, rids, rdists, rcounts, what="edgeless graph")
    # dimension mismatch
    with pytest.raises(eng.NgtGpuError):
        ix.linear_search(np.zeros((1, 25), np.uint8), 3)
    ix.close()
    with pytest.raises(eng.NgtGpuError):
        eng.GpuIndex(po.UINT8, po.COSINE, 16)                    # unsupported pair


def test_long_rows_and_wide_results(eng, port):
    """dimension > 1024 floats takes the shared-memory-query instantiation; k = 300 the smem result list."""
    rng = np.random.default_rng(11)
    n, d = 700, 1100
    objs = rng.integers(0, 16, (n, d)).astype(np.float32)
    q = rng.integers(0, 16, (6, d)).astype(np.float32)
    ix = eng.GpuIndex(po.FLOAT, po.L2, d)
    ix.set_objects(objs)
    pobj, pq = po.pad_objects(objs, po.FLOAT), po.pad_queries(q, po.FLOAT)
    ids, dists, counts = ix.linear_search(q, 300)
    rids, rdists, rcounts = port.linear_search(po.L2, po.FLOAT, pobj, pq, 300)
    assert_bit_exact(ids, dists, counts, rids, rdists, rcounts, what="long rows linear")
    gids, _, gcounts = ix.linear_search(objs, 10)
    row_ptr, col = _knn_csr(gids, gcounts)
    ix.set_graph(row_ptr, col)
    seeds = np.tile(np.arange(1, 11, dtype=np.uint32), (6, 1))
    ids, dists, counts = ix.search(q, 100, 0.2, edge_size=0, seeds=seeds)
    rids, rdists, rcounts, _ = port.graph_search(po.L2, po.FLOAT, pobj, row_ptr, col, pq, seeds, 100, 0.2)
    assert_bit_exact(ids, dists, counts, rids, rdists, rcounts, what="long rows graph")
    ix.close()


@pytest.mark.parametrize("kind", ["f32l2_128", "u8l2_128", "f32l2_48", "u8ham_128", "f32cos_100"])
def test_seed_selection_is_the_exact_nearest_pivots(eng, port, kind):
    """ngtgpu_select_seeds returns the n_seeds nearest pivots by (distance, id). With every object a pivot that is the
    exhaustive search itself (same ids, same order), for every row-width class of the one-warp-per-query kernel."""
    import ctypes as C
    from ngt_b200 import _lib, synth
    name, dim = kind.split("_")
    dim = int(dim)
    n, nq = 3000, 257
    base, qs = synth.make("sift", n, 1)[:, :dim], synth.make("sift", nq, 2)[:, :dim]
    if name == "u8l2":
        otype, dtype, objs, q = po.UINT8, po.L2, base.astype(np.uint8), qs
    elif name == "u8ham":
        otype, dtype = po.UINT8, po.HAMMING
        objs, q = synth.hamming_from(base, 64.0), synth.hamming_from(qs, 64.0).astype(np.float32)
    elif name == "f32cos":
        otype, dtype = po.FLOAT, po.COSINE
        objs, q = (base - 64.0).astype(np.float32) / 40.0, (qs - 64.0).astype(np.float32) / 40.0
    else:
        otype, dtype, objs, q = po.FLOAT, po.L2, base, qs
    ix = eng.GpuIndex(otype, dtype, objs.shape[1])
    ix.set_objects(objs)
    ix.build_seed_table(n, 1)   # one pivot per stride of one id: every object
    lib = _lib.load()
    lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    q = np.ascontiguousarray(q, np.float32)
    for ns in (1, 10, 32):
        seeds = np.zeros((nq, ns), np.uint32)
        _lib.check(lib.ngtgpu_select_seeds(ix._h, q.ctypes.data, _lib.OBJECT_FLOAT, nq, ns, seeds.ctypes.data))
        ids, dists, counts = ix.linear_search(q, ns)
        if name == "f32cos":   # general float data: equal up to swaps among (near-)ties
            same = (np.sort(seeds, 1) == np.sort(np.asarray(ids), 1)).all(1).mean()
            assert same > 0.98, (kind, ns, same)
        else:
            assert (seeds == np.asarray(ids)).all(), (kind, ns)
    ix.close()


@pytest.mark.parametrize("kind", ["f32l2_128", "u8l2_128", "f32cos_100", "f32l2_48"])
def test_seeds_selected_inside_the_traversal_equal_explicit_seeds(eng, kind):
    """Without a seed list the nearest pivots are selected by a pass of their own or (seed fusion) by the lean kernel
    itself in the first tier, which hands them to the later tiers: every output equals a search from the seeds
    ngtgpu_select_seeds returns, also when queries overflow."""
    import ctypes as C
    from ngt_b200 import _lib, synth
    name, dim = kind.split("_")
    dim = int(dim)
    n, nq = 6000, 200
    base, qs = synth.make("sift", n, 1)[:, :dim], synth.make("sift", nq, 2)[:, :dim]
    if name == "u8l2":
        otype, dtype, objs, q = po.UINT8, po.L2, base.astype(np.uint8), qs
    elif name == "f32cos":
        otype, dtype = po.FLOAT, po.COSINE
        objs, q = (base - 64.0).astype(np.float32) / 40.0, (qs - 64.0).astype(np.float32) / 40.0
    else:
        otype, dtype, objs, q = po.FLOAT, po.L2, base, qs
    q = np.ascontiguousarray(q, np.float32)
    ix = eng.GpuIndex(otype, dtype, objs.shape[1])
    ix.set_objects(objs)
    gids, _, gcounts = ix.linear_search(objs.astype(np.float32), 13)
    row_ptr, col = _knn_csr(gids, gcounts)
    ix.set_graph(row_ptr, col)
    ix.build_seed_table(200, 3)
    lib = _lib.load()
    lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    seeds = np.zeros((nq, 10), np.uint32)
    _lib.check(lib.ngtgpu_select_seeds(ix._h, q.ctypes.data, _lib.OBJECT_FLOAT, nq, 10, seeds.ctypes.data))
    for hb, qc in ((14, 512), (9, 64)):
        ix.set_search_workspace(hash_bits=hb, queue_cap=qc)
        b = ix.search(q, 10, 0.2, edge_size=12, seeds=seeds, with_stats=True)
        for fused in (True, False):
            ix.set_seed_fusion(fused)
            a = ix.search(q, 10, 0.2, edge_size=12, n_seeds=10, with_stats=True)
            for x, y in zip(a, b):
                assert (np.asarray(x).view(np.uint32) == np.asarray(y).view(np.uint32)).all(), (kind, hb, fused)
    assert ix.last_overflows > 0
    ix.close()


def test_device_seed_table_recall(eng, port):
    """Seeds from the device pivot table (stand-in for the DVP-tree leaf): recall at the reference's epsilon
    is at least what the restated search reaches from the same seeds, and >= 0.9 on this set."""
    from ngt_b200 import synth
    n, nq, k = 20000, 200, 10
    base, qs = synth.make("sift", n, 1), synth.make("sift", nq, 2)
    ix = eng.GpuIndex(po.FLOAT, po.L2, 128)
    ix.set_objects(base)
    gids, _, gcounts = ix.linear_search(base, 17)
    # symmetrised kNN graph (what ANNG construction approximates), lists ordered by (distance,id)
    pobj, pq = po.pad_objects(base, po.FLOAT), po.pad_queries(qs, po.FLOAT)
    src = np.repeat(np.arange(1, n + 1, dtype=np.uint32), 16)
    dst = gids[:, 1:17].reshape(-1)
    e = np.unique(np.concatenate([np.stack([src, dst], 1), np.stack([dst, src], 1)]), axis=0)
    d = np.linalg.norm(base[e[:, 0] - 1] - base[e[:, 1] - 1], axis=1)
    order = np.lexsort((e[:, 1], d, e[:, 0]))
    e = e[order]
    row_ptr = np.zeros(n + 2, np.uint64)
    np.add.at(row_ptr, e[:, 0].astype(np.int64) + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.uint64)
    col = e[:, 1].astype(np.uint32)
    ix.set_graph(row_ptr, col)
    ix.build_seed_table(1024, 5)
    gt_ids, gt_d, _ = ix.linear_search(qs, k)
    ids, dists, counts = ix.search(qs, k, 0.1, edge_size=0, n_seeds=10)
    rec = port.mean_recall(ids, dists, counts, gt_ids, gt_d)
    assert rec >= 0.9, rec
    # torch (already-in-HBM) entry point returns the same thing
    import torch
    tq = torch.from_numpy(qs).cuda()
    tids, tdists, tcounts = ix.search(tq, k, 0.1, edge_size=0, n_seeds=10)
    torch.cuda.synchronize()
    assert (tids.cpu().numpy().astype(np.uint32) == ids).all()
    assert (tdists.cpu().numpy().view(np.uint32) == dists.view(np.uint32)).all()
    ix.close()


def test_pack_and_merge_kernels_match_host(eng):
    """ngtgpu_pack_keys / ngtgpu_merge_keys (the multi-GPU exchange format and merge) against numpy."""
    import torch
    from ngt_b200 import sharded
    rng = np.random.default_rng(9)
    world, nq, k = 3, 70, 10
    lists = []
    for r in range(world):
        d = np.sort(rng.integers(0, 60, (nq, k)).astype(np.float32), axis=1)     # many cross-shard ties
        ids = rng.integers(1, 1000, (nq, k)).astype(np.uint32)
        counts = rng.integers(0, k + 1, nq).astype(np.uint32)
        order = np.lexsort((ids, d), axis=1)
        d, ids = np.take_along_axis(d, order, 1), np.take_along_axis(ids, order, 1)
        lists.append((ids, d, counts))
    dev = torch.device("cuda", 0)
    gathered = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
    host_keys = []
    for r, (ids, d, counts) in enumerate(lists):
        s = sharded.ShardedSearcher(None, r, world, 1000)
        lib = sharded._fn()
        ti, td, tc = (torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).to(dev) for x in (ids, d, counts))
        from ngt_b200 import _lib
        _lib.check(lib.ngtgpu_pack_keys(ti.data_ptr(), td.data_ptr(), tc.data_ptr(), nq, k, s.id_offset,
                                        gathered[r].data_ptr(), torch.cuda.current_stream().cuda_stream))
        host_keys.append(sharded.pack_keys_host(ids, d, counts, r * 1000))
    torch.cuda.synchronize()
    assert (gathered.cpu().numpy().view(np.uint64) == np.stack(host_keys)).all()
    oi = torch.empty((nq, k), dtype=torch.int32, device=dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int32, device=dev)
    _lib.check(lib.ngtgpu_merge_keys(gathered.data_ptr(), world, nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    hi, hd, hc = sharded.merge_keys_host(np.stack(host_keys), k)
    assert (oc.cpu().numpy().astype(np.uint32) == hc).all()
    assert_bit_exact(oi.cpu().numpy().astype(np.uint32), od.cpu().numpy(), hc, hi, hd, hc, what="merge kernel")


@pytest.mark.parametrize("case", ["sift_l2_exact_bf16", "glove_l2_split", "glove_ncos_split", "glove_cos_split", "sift_u8_l2",
                                  "sift_hamming", "gist_l2_split_streamed", "gist_cos_split_streamed", "sift_l2_forced_stream"])
def test_tensor_core_knn_matches_cuda_core_scan(eng, port, case, monkeypatch):
    """knn_tc.cu (tcgen05 filter + exact re-evaluation) returns exactly what the CUDA-core scan returns: ids,
    distance bits and counts, for batches, kNN-graph construction (self excluded) and with removed slots."""
    from ngt_b200 import build, synth
    n, nq = 40000, 1500
    ot = po.FLOAT
    if case == "sift_l2_forced_stream":   # the streamed query operand on a K axis that would fit: same answers
        monkeypatch.setenv("NGTGPU_TC_STREAM", "1")
        base, qs, dt = synth.make("sift", n, 1), synth.make("sift", nq, 2), po.L2
    elif case.startswith("gist"):         # 960-d split floats: 46 k-chunks, query operand streamed through the ring
        n = 33000
        base, qs = synth.make("gist", n, 1), synth.make("gist", nq, 2)
        dt = po.L2 if case == "gist_l2_split_streamed" else po.COSINE
    elif case == "sift_l2_exact_bf16":
        base, qs, dt = synth.make("sift", n, 1), synth.make("sift", nq, 2), po.L2
    elif case == "sift_u8_l2":      # uint8 values are bf16 numbers: one exact segment, integer re-evaluation (dp4a)
        base, qs, dt, ot = synth.make("sift", n, 1).astype(np.uint8), synth.make("sift", nq, 2).astype(np.uint8), po.L2, po.UINT8
    elif case == "sift_hamming":    # Hamming = squared L2 of the bits: one K element per bit
        base, qs = synth.hamming_from(synth.make("sift", n, 1), 64.0), synth.hamming_from(synth.make("sift", nq, 2), 64.0)
        dt, ot = po.HAMMING, po.UINT8
    else:
        base, qs = synth.make("glove", n, 1), synth.make("glove", nq, 2)
        dt = {"glove_l2_split": po.L2, "glove_ncos_split": po.NORMALIZED_COSINE, "glove_cos_split": po.COSINE}[case]
    ix = eng.GpuIndex(ot, dt, base.shape[1])
    ix.set_objects(base)
    for k in (10, 64, 120):
        ix.set_tensor_core(True)
        before = ix.tensor_core_batches
        ids, dists, counts = ix.linear_search(qs, k)
        assert ix.tensor_core_batches == before + 1, "the tensor-core path did not run"
        ix.set_tensor_core(False)
        rids, rdists, rcounts = ix.linear_search(qs, k)
        assert ix.tensor_core_batches == before + 1
        assert_bit_exact(ids, dists, counts, rids, rdists, rcounts, what="%s k=%d" % (case, k))
    # ... and what the C restatement of the reference's linearSearch returns, directly (so that a bug shared by both device
    # paths cannot hide): bit for bit where the arithmetic is exact, within the float tolerance otherwise
    ix.set_tensor_core(True)
    before = ix.tensor_core_batches
    ids, dists, counts = ix.linear_search(qs, 10)
    assert ix.tensor_core_batches == before + 1
    sub = 0 if case == "glove_ncos_split" else 40   # (normalised kinds: the engine prepares queries itself; their parity is test_synth_golden's)
    pobj = po.pad_objects(base, ot)
    pq = po.pad_queries(np.asarray(qs[:sub], np.float32 if ot == po.FLOAT else np.uint8), ot)
    rids, rdists, rcounts = port.linear_search(dt, ot, pobj, pq, 10) if sub else (ids[:0], dists[:0], counts[:0])
    if case in ("sift_l2_exact_bf16", "sift_u8_l2", "sift_hamming", "sift_l2_forced_stream"):
        assert_bit_exact(ids[:sub], dists[:sub], counts[:sub], rids, rdists, rcounts, what=case + " vs the C restatement")
    elif case == "gist_cos_split_streamed":
        # 1 - cos of all-positive 960-d rows is ~0.02: the subtraction cancels five digits, so the tolerance is on that scale
        assert_float_parity(ids[:sub], dists[:sub], counts[:sub], rids, rdists, rcounts, what=case + " vs the C restatement",
                            rtol=1e-4, tie=1e-4)
    else:
        assert_float_parity(ids[:sub], dists[:sub], counts[:sub], rids, rdists, rcounts, what=case + " vs the C restatement")
    # kNN-graph construction: stored rows as queries, the row itself dropped
    ix.set_tensor_core(True)
    gi, gd, gc = build.knn_graph(ix, 16, batch=8192)
    ix.set_tensor_core(False)
    ri, rd, rc = build.knn_graph(ix, 16, batch=8192)
    import torch
    torch.cuda.synchronize()
    assert (gc == rc).all() and (gi == ri).all() and (gd.view(torch.int32) == rd.view(torch.int32)).all()
    assert not (gi.cpu().numpy() == np.arange(1, n + 1)[:, None]).any()
    # removed slots are never returned
    removed = np.arange(5, n, 7, dtype=np.uint32)
    ix.set_removed(removed)
    ix.set_tensor_core(True)
    ids, dists, counts = ix.linear_search(qs, 10)
    ix.set_tensor_core(False)
    rids, rdists, rcounts = ix.linear_search(qs, 10)
    assert_bit_exact(ids, dists, counts, rids, rdists, rcounts, what=case + " removed")
    assert not np.isin(ids, removed).any()
    ix.close()

"""Pins the C restatement (oracle/ngt_oracle.c) to the reference: README known-answer listing and
reference-generated golden vectors (tests/golden/make_golden.py). CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from parity import (EDGE_GRID, EPS_GRID, FLOAT_CASES, INTEGER_EXACT, NORMALIZED, assert_bit_exact,
                    assert_float_parity, grid_key)
from conftest import GOLDEN


def test_readme_known_answer(port, sift5k):
    """bin/ngt/README.md:254-323: tree seeds (from the reference) -> restated search == the listing."""
    kat = json.load(open(os.path.join(GOLDEN, "readme_kat.json")))
    objs = po.pad_objects(sift5k["data"], po.UINT8)
    qs = po.pad_queries(sift5k["queries"], po.UINT8)
    es, base, rate = [int(v) for v in sift5k["prop"]]
    cap = port.edge_size(-1, es, np.float32(1.1), base, rate)
    ids, dists, counts, _ = port.graph_search(po.L2, po.UINT8, objs, sift5k["row_ptr"], sift5k["col"], qs,
                                              sift5k["tree_seeds"], 20, 0.1, edge_size=cap)
    for q in range(3):
        got = [[int(ids[q, i]), "%g" % float(dists[q, i])] for i in range(20)]
        assert got == kat[q]
    # the listing is also the exact top-20, so it pins linearSearch + the uint8 L2 kernel too
    lids, ldists, _ = port.linear_search(po.L2, po.UINT8, objs, qs, 20)
    for q in range(3):
        assert [[int(lids[q, i]), "%g" % float(ldists[q, i])] for i in range(20)] == kat[q]


@pytest.mark.parametrize("tag,otype", [("u8", po.UINT8), ("f32", po.FLOAT)])
def test_sift5k_bit_exact(port, sift5k, tag, otype):
    objs = po.pad_objects(sift5k["data"], otype)
    qs = po.pad_queries(sift5k["queries"], otype)
    lids, ldists, lc = port.linear_search(po.L2, otype, objs, qs, 20)
    assert_bit_exact(lids, ldists, lc, sift5k[tag + "_lin_ids"], sift5k[tag + "_lin_dists"], what="linear")
    es, base, rate = [int(v) for v in sift5k["prop"]]
    for eps in EPS_GRID:
        for e in EDGE_GRID:
            cap = port.edge_size(e, es, np.float32(eps + 1.0), base, rate)
            ids, dists, counts, stats = port.graph_search(po.L2, otype, objs, sift5k["row_ptr"], sift5k["col"], qs,
                                                          sift5k["seeds"], 20, eps, edge_size=cap)
            key = grid_key(tag, eps, e)
            assert_bit_exact(ids, dists, counts, sift5k[key + "_ids"], sift5k[key + "_dists"],
                             sift5k[key + "_counts"], what=key)
            ref = sift5k[key + "_stats"]
            # the writable path counts distance computations after the seeds (Graph.cpp:604) and every
            # adjacency entry it examines (Graph.cpp:590)
            assert (stats[:, 0] - sift5k["seeds"].shape[1] == ref[:, 0]).all(), key
            assert (stats[:, 1] == ref[:, 1]).all(), key


def _case(z, tag):
    graph_tag = tag
    data_tag = "f32l2" if tag == "onng" else tag
    ot, dt, dim, es, base, rate = [int(v) for v in z[graph_tag + "_meta"]]
    return dict(otype=ot, dtype=dt, dim=dim, es=es, base=base, rate=rate,
                objects=z[data_tag + "_objects"], queries=z[data_tag + "_queries"],
                row_ptr=z[graph_tag + "_row_ptr"], col=z[graph_tag + "_col"], seeds=z[graph_tag + "_seeds"])


def _prep(port, c):
    objs = po.pad_objects(c["objects"], c["otype"])
    q = c["queries"]
    if c["dtype"] in NORMALIZED:
        q = port.normalize(q)
    return objs, po.pad_queries(q, c["otype"])


@pytest.mark.parametrize("tag", INTEGER_EXACT)
def test_synth_integer_bit_exact(port, synth_golden, tag):
    z = synth_golden
    c = _case(z, tag)
    objs, qs = _prep(port, c)
    if tag != "onng":
        lids, ldists, lc = port.linear_search(c["dtype"], c["otype"], objs, qs, 10)
        assert_bit_exact(lids, ldists, lc, z[tag + "_lin_ids"], z[tag + "_lin_dists"], what=tag + " linear")
        rad = float(z[tag + "_linr_radius"][0])
        lids, ldists, lc = port.linear_search(c["dtype"], c["otype"], objs, qs, 10, radius=rad)
        assert_bit_exact(lids, ldists, lc, z[tag + "_linr_ids"], z[tag + "_linr_dists"], z[tag + "_linr_counts"],
                         what=tag + " linear radius")
    for eps in EPS_GRID:
        for e in EDGE_GRID:
            cap = port.edge_size(e, c["es"], np.float32(eps + 1.0), c["base"], c["rate"])
            ids, dists, counts, stats = port.graph_search(c["dtype"], c["otype"], objs, c["row_ptr"], c["col"], qs,
                                                          c["seeds"], 10, eps, edge_size=cap)
            key = grid_key(tag, eps, e)
            assert_bit_exact(ids, dists, counts, z[key + "_ids"], z[key + "_dists"], z[key + "_counts"], what=key)
            assert (stats[:, 1] == z[key + "_stats"][:, 1]).all(), key


@pytest.mark.parametrize("tag", FLOAT_CASES)
def test_synth_float_tolerance(port, synth_golden, tag):
    z = synth_golden
    c = _case(z, tag)
    objs, qs = _prep(port, c)
    lids, ldists, lc = port.linear_search(c["dtype"], c["otype"], objs, qs, 10)
    assert_float_parity(lids, ldists, lc, z[tag + "_lin_ids"], z[tag + "_lin_dists"], what=tag + " linear")
    for eps in EPS_GRID:
        for e in EDGE_GRID:
            cap = port.edge_size(e, c["es"], np.float32(eps + 1.0), c["base"], c["rate"])
            ids, dists, counts, _ = port.graph_search(c["dtype"], c["otype"], objs, c["row_ptr"], c["col"], qs,
                                                      c["seeds"], 10, eps, edge_size=cap)
            key = grid_key(tag, eps, e)
            assert_float_parity(ids, dists, counts, z[key + "_ids"], z[key + "_dists"], z[key + "_counts"], what=key)


def test_normalize_matches_reference_within_tolerance(port, synth_golden):
    """ObjectSpace.h:251-266: the reference's stored (normalised) rows have unit norm; ours agree to 1e-6."""
    stored = synth_golden["glove_ncos_objects"]
    assert np.allclose(np.linalg.norm(stored.astype(np.float64), axis=1), 1.0, atol=1e-6)
    raw = synth_golden["glove_l2_objects"]
    mine = port.normalize(raw)
    assert np.abs(mine - stored).max() <= 1e-6
    with pytest.raises(ValueError):
        port.normalize(np.zeros((1, 8), np.float32))


def test_edge_size_modes(port):
    """Graph.h:675-692."""
    assert port.edge_size(-1, 40, 1.1, 30, 20) == 40
    assert port.edge_size(0, 40, 1.1, 30, 20) == 2 ** 31 - 1
    assert port.edge_size(7, 40, 1.1, 30, 20) == 7
    assert port.edge_size(-2, 40, np.float32(1.1), 30, 20) == 30 + 100
    assert port.edge_size(-1, -2, np.float32(1.1), 32, 8) == 32 + int(10 ** (float(np.float32(1.1) - 1.0) * 8))
    assert port.edge_size(-3, 40, 1.1, 30, 20) == -1


def test_recall_definition(port):
    """Optimizer.h:496-507: id hit, or distance <= farthest ground-truth distance."""
    gt_ids = np.array([1, 2, 3, 4], np.uint32)
    gt_d = np.array([1.0, 2.0, 3.0, 4.0], np.float32)
    assert port.recall(np.array([1, 2, 9, 8], np.uint32), np.array([1, 2, 4.0, 5.0], np.float32), gt_ids, gt_d) == 0.75
    assert port.recall(np.array([7], np.uint32), np.array([9.0], np.float32), gt_ids, gt_d) == 0.0


def test_adjust_paths_restatement_matches_the_reference():
    """GraphReconstructor::adjustPathsEffectively (GraphReconstructor.h:197-386): the sequential restatement in
    oracle/pyoracle.py against the graphs the reference's GraphOptimizer::execute wrote with shortcut reduction on
    (tests/golden/adjust_paths.npz, made by tests/golden/make_golden_adjust_paths.py)."""
    z = np.load(os.path.join(GOLDEN, "adjust_paths.npz"))
    for tag in ("sift", "glove"):
        for o, i in ((5, 20), (10, 40)):
            k = "%s_o%d_i%d" % (tag, o, i)
            out = po.adjust_paths_loop(z[k + "_in_row_ptr"], z[k + "_in_col"], z[k + "_in_dist"], 0)
            arp, acol, adist = z[k + "_adj_row_ptr"], z[k + "_adj_col"], z[k + "_adj_dist"]
            assert len(acol) < len(z[k + "_in_col"])
            for nid in range(len(arp) - 1):
                ref = [(float(adist[e]), int(acol[e])) for e in range(int(arp[nid]), int(arp[nid + 1]))]
                assert ref == out[nid], (k, nid)


# ---- construction path: the sequential restatements in oracle/pyoracle.py against graphs
# the UNMODIFIED reference built (tests/golden/make_golden_anng.py: createIndex's batched loop and refineANNG on a
# graph-only index with SeedType FixedNodes, the reference's own deterministic seed mode)
def _lists(rp, col, dist):
    rp = np.asarray(rp, np.int64)
    return [[(float(dist[e]), int(col[e])) for e in range(int(rp[i]), int(rp[i + 1]))] for i in range(len(rp) - 1)]


def anng_case(z, tag):
    from ngt_b200 import synth
    objtype, n, n_first, seed, e, es, ss, bs = [int(v) for v in z[tag + "_meta"]]
    base = synth.make("sift", n, seed)
    otype = po.UINT8 if chr(objtype) == "c" else po.FLOAT
    return dict(base=base, otype=otype, n=n, n_first=n_first, e=e, es=es, ss=ss, bs=bs,
                lists=_lists(z[tag + "_row_ptr"], z[tag + "_col"], z[tag + "_dist"])[:n + 1])


@pytest.mark.parametrize("tag", ["f_b200", "f_b64_all", "f_b1000_s5", "u8_b200"])
def test_anng_build_loop_restatement_equals_the_reference(port, tag):
    """Index.cpp:631-719,721-792 + Index.h:815-837 + Graph.h:611-626,845-886: build, then insertion into the built
    graph, edge for edge (ids and float bits) against NGT::Index::createIndex of the reference."""
    c = anng_case(np.load(os.path.join(GOLDEN, "anng_build.npz")), tag)
    pobj = po.pad_objects(c["base"], c["otype"])
    seeds = po.fixed_node_seeds(c["n"], c["ss"])
    rows_int = np.vstack([np.zeros((1, c["base"].shape[1]), np.int64), c["base"].astype(np.int64)])
    # the uint8 case takes the in-batch distances from the C restatement of the comparator instead
    ri = rows_int if c["otype"] == po.FLOAT else None
    got = po.build_anng_loop(port, pobj, ri, seeds, 1, c["n_first"], None, c["e"], 0.1, c["es"], c["bs"], po.L2, c["otype"])
    if c["n_first"] < c["n"]:
        got = po.build_anng_loop(port, pobj, ri, seeds, c["n_first"] + 1, c["n"] - c["n_first"], got, c["e"], 0.1, c["es"],
                                 c["bs"], po.L2, c["otype"])
    assert got == c["lists"]


@pytest.mark.parametrize("tag", ["r0_all", "r0_b400", "r12_b500", "rm6_all", "u8_r0_b500"])
def test_refine_anng_restatement_equals_the_reference(port, tag):
    """GraphReconstructor.h:814-924 against GraphReconstructor::refineANNG of the reference (all three noOfEdges modes,
    several batch sizes)."""
    zb = np.load(os.path.join(GOLDEN, "anng_build.npz"))
    zr = np.load(os.path.join(GOLDEN, "refine_anng.npz"))
    src = str(zr[tag + "_src"][0])
    c = anng_case(zb, src)
    noe, explore, bs = [int(v) for v in zr[tag + "_meta"]]
    eps = float(zr[tag + "_eps"][0])
    pobj = po.pad_objects(c["base"], c["otype"])
    seeds = po.fixed_node_seeds(c["n"], c["ss"])
    n = c["n"]
    got = po.refine_anng_loop(port, po.L2, c["otype"], pobj, zb[src + "_row_ptr"][:n + 2].astype(np.uint64), zb[src + "_col"],
                              zb[src + "_dist"], seeds, eps, noe, c["es"] if c["es"] else 2 ** 31 - 1, bs, c["e"])
    assert got == _lists(zr[tag + "_row_ptr"], zr[tag + "_col"], zr[tag + "_dist"])[:n + 1]


def test_remove_edges_reliably_restatement_equals_the_reference(port):
    """Graph.cpp:641-864 (NGT::Index::remove): nine removals from the reference-built ANNG of anng_build.npz, edge for edge
    against the graph the reference was left with (tests/golden/remove.npz)."""
    zb = np.load(os.path.join(GOLDEN, "anng_build.npz"))
    zr = np.load(os.path.join(GOLDEN, "remove.npz"))
    c = anng_case(zb, "f_b200")
    pobj = po.pad_objects(c["base"], po.FLOAT)
    lists = [list(l) for l in c["lists"]]
    for rid in zr["removed"]:
        po.remove_edges_reliably_loop(port, po.L2, po.FLOAT, pobj, lists, int(rid))
    assert lists == _lists(zr["row_ptr"], zr["col"], zr["dist"])[:c["n"] + 1]

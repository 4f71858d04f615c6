"""Graph surgery of ONNG construction on the device: reconstructGraph + adjustPathsEffectively
(lib/NGT/GraphReconstructor.h:425-561, 197-386) against the graphs the unmodified reference wrote
(tests/golden/adjust_paths.npz) and against the sequential restatement in oracle/pyoracle.py on larger graphs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _lists(rp, col, dist):
    rp, col, dist = np.asarray(rp), np.asarray(col), np.asarray(dist)
    return [[(float(dist[e]), int(col[e])) for e in range(int(rp[i]), int(rp[i + 1]))] for i in range(len(rp) - 1)]


def _adjust_on_device(rp, col, dist, min_edges=0):
    import torch
    from ngt_b200 import build
    dev = torch.device("cuda", 0)
    out = build.adjust_paths(torch.from_numpy(np.asarray(rp).astype(np.int64)).to(dev),
                             torch.from_numpy(np.asarray(col).astype(np.int32)).to(dev),
                             torch.from_numpy(np.asarray(dist).astype(np.float32)).to(dev), min_edges, with_stats=True)
    return out[0].cpu().numpy(), out[1].cpu().numpy().astype(np.uint32), out[2].cpu().numpy(), out[3]


@pytest.mark.parametrize("key", ["sift_o5_i20", "sift_o10_i40", "glove_o5_i20", "glove_o10_i40"])
def test_adjust_paths_equals_the_reference(eng, key):
    z = np.load(os.path.join(GOLDEN, "adjust_paths.npz"))
    rp, col, dist, st = _adjust_on_device(z[key + "_in_row_ptr"], z[key + "_in_col"], z[key + "_in_dist"])
    assert (rp == z[key + "_adj_row_ptr"].astype(np.int64)).all()
    assert (col == z[key + "_adj_col"]).all()
    assert (dist.view(np.uint32) == z[key + "_adj_dist"].view(np.uint32)).all()
    assert st["removed"] == len(z[key + "_in_col"]) - len(z[key + "_adj_col"])


@pytest.mark.parametrize("min_edges", [0, 3, 12])
def test_adjust_paths_min_edges_and_long_lists(eng, min_edges):
    """A device-built ONNG-style graph with long reverse lists (hub nodes beyond the shared-memory staging limit of
    the candidate kernel are made by a duplicated cluster centre) against the sequential restatement."""
    import torch
    from ngt_b200 import build, synth
    n = 4000
    base = synth.make("sift", n, 7)
    base[1:2600] = base[0] + np.random.RandomState(3).randint(0, 3, size=(2599, base.shape[1]))   # one dense blob
    ix = eng.GpuIndex(po.FLOAT, po.L2, base.shape[1])
    ix.set_objects(base)
    ids, dists, counts = build.knn_graph(ix, 24)
    rp, col, dist = build.reconstruct_graph(ids, dists, counts, 6, 24)
    deg = (rp[2:] - rp[1:-1])
    out_rp, out_col, out_dist, st = build.adjust_paths(rp, col, dist, min_edges, with_stats=True)
    ref = po.adjust_paths_loop(rp.cpu().numpy(), col.cpu().numpy().astype(np.uint32), dist.cpu().numpy(), min_edges)
    got = _lists(out_rp.cpu().numpy(), out_col.cpu().numpy().astype(np.uint32), out_dist.cpu().numpy())
    assert got == ref
    assert st["removed"] > 0 and int(deg.max()) > 0
    kept_deg = out_rp[2:] - out_rp[1:-1]
    assert bool(((kept_deg >= torch.clamp(deg, max=min_edges)) | (deg == 0)).all())
    ix.close()


def test_adjust_paths_empty_and_edgeless(eng):
    import torch
    from ngt_b200 import build
    dev = torch.device("cuda", 0)
    rp = torch.zeros(12, dtype=torch.int64, device=dev)
    out = build.adjust_paths(rp, torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, device=dev))
    assert out[1].numel() == 0 and int(out[0][-1]) == 0
    with pytest.raises(eng.NgtGpuError):
        build.adjust_paths(rp.cpu(), torch.zeros(0, dtype=torch.int32), torch.zeros(0))


@pytest.mark.parametrize("o,i", [(5, 20), (10, 40), (0, 15)])
def test_reconstruct_graph_on_device_equals_the_reference(eng, o, i):
    """ngtgpu_graph_reconstruct against the graph the reference's GraphOptimizer::execute wrote from its own ANNG
    (path adjustment off): tests/golden/reconstruct.npz."""
    import torch
    from ngt_b200 import build
    z = np.load(os.path.join(GOLDEN, "reconstruct.npz"))
    dev = torch.device("cuda", 0)
    rp, col, dist = build.reconstruct_graph_device(torch.from_numpy(z["anng_row_ptr"].astype(np.int64)).to(dev),
                                                   torch.from_numpy(z["anng_col"].astype(np.int32)).to(dev),
                                                   torch.from_numpy(z["anng_dist"]).to(dev), o, i)
    key = "o%d_i%d" % (o, i)
    assert (rp.cpu().numpy() == z[key + "_row_ptr"].astype(np.int64)).all()
    assert (col.cpu().numpy().astype(np.uint32) == z[key + "_col"]).all()
    assert (dist.cpu().numpy().view(np.uint32) == z[key + "_dist"].view(np.uint32)).all()


def test_onng_recipe_end_to_end_equals_the_reference(eng):
    """reconstruct + adjust on the device == `ngt reconstruct-graph -o 5 -i 20` of the reference on its own ANNG."""
    import torch
    from ngt_b200 import build
    z = np.load(os.path.join(GOLDEN, "reconstruct.npz"))
    a = np.load(os.path.join(GOLDEN, "adjust_paths.npz"))
    dev = torch.device("cuda", 0)
    # (adjust_paths.npz's sift ANNG is reconstruct.npz's: same data, same build parameters)
    assert (a["sift_o5_i20_in_col"] == z["o5_i20_col"]).all()
    g = build.reconstruct_graph_device(torch.from_numpy(z["anng_row_ptr"].astype(np.int64)).to(dev),
                                       torch.from_numpy(z["anng_col"].astype(np.int32)).to(dev),
                                       torch.from_numpy(z["anng_dist"]).to(dev), 5, 20)
    rp, col, dist = build.adjust_paths(*g)
    assert (rp.cpu().numpy() == a["sift_o5_i20_adj_row_ptr"].astype(np.int64)).all()
    assert (col.cpu().numpy().astype(np.uint32) == a["sift_o5_i20_adj_col"]).all()


@pytest.mark.parametrize("no_of_edges,batch", [(0, 700), (0, 10000), (12, 900), (-6, 10000)])
def test_refine_anng_equals_the_sequential_restatement(eng, port, no_of_edges, batch):
    """GraphReconstructor::refineANNG on the device (batched self-search + merges) against the sequential restatement
    over the C oracle's search, same seeds: integer-valued L2 data, so ids and distance bits must be identical."""
    import ctypes as C
    import torch
    from ngt_b200 import _lib, build, synth
    n, ec = 2500, 8
    base = synth.make("sift", n, 11)
    ix = eng.GpuIndex(po.FLOAT, po.L2, base.shape[1])
    ix.set_objects(base)
    ids, dists, counts = build.knn_graph(ix, 6)
    rp, col, dist = build.reconstruct_graph(ids, dists, counts, 6, 6)      # a sparse ANNG-like start
    ix.set_graph(rp, col)
    ix.build_seed_table(256, 5)
    lib = _lib.load()
    lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    seeds = np.zeros((n, 10), np.uint32)
    _lib.check(lib.ngtgpu_select_seeds(ix._h, base.ctypes.data, _lib.OBJECT_FLOAT, n, 10, seeds.ctypes.data))
    out_rp, out_col, out_dist = build.refine_anng(ix, rp, col, dist, epsilon=0.1, no_of_edges=no_of_edges, edge_size=-1,
                                                   batch_size=batch, edge_size_for_creation=ec, n_seeds=10)
    got = _lists(out_rp.cpu().numpy(), out_col.cpu().numpy().astype(np.uint32), out_dist.cpu().numpy())
    pobj = po.pad_objects(base, po.FLOAT)
    ref = po.refine_anng_loop(port, po.L2, po.FLOAT, pobj, rp.cpu().numpy(), col.cpu().numpy().astype(np.uint32),
                              dist.cpu().numpy(), seeds, 0.1, no_of_edges, 40, batch, ec)
    assert got == ref
    assert no_of_edges < 0 or out_col.numel() != col.numel()
    # the index now searches the refined graph
    if no_of_edges == 0:
        r = ix.search(base[:50], 5, 0.1)
        assert (r[0][:, 0] == np.arange(1, 51)).all()
    ix.close()


@pytest.mark.parametrize("batch,edge_size", [(200, 40), (64, 0), (1000, 5)])
def test_anng_construction_loop_equals_the_sequential_restatement(eng, port, batch, edge_size):
    """NGT::Index::createIndex as the reference runs it (batches searched on the frozen graph, in-batch distances,
    insertion with reverse edges) on the device, then insertion of more objects into the finished index: the same
    graph as the sequential restatement over the C oracle's search, edge for edge (integer-valued L2 data)."""
    import ctypes as C
    from ngt_b200 import _lib, build, synth
    n, n_first, e = 2600, 2000, 8
    base = synth.make("sift", n, 21)
    ix = eng.GpuIndex(po.FLOAT, po.L2, base.shape[1])
    ix.set_objects(base)
    ix.set_search_property(edge_size, 30, 20)
    lib = _lib.load()
    lib.ngtgpu_select_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    # the seeds every batch will use: nearest pivots among the ids inserted before the batch
    starts = [(s, min(batch, n_first + 1 - s)) for s in range(1, n_first + 1, batch)] + \
             [(s, min(batch, n + 1 - s)) for s in range(n_first + 1, n + 1, batch)]
    seeds = np.zeros((n, 10), np.uint32)
    for s, m in starts[1:]:
        _lib.check(lib.ngtgpu_index_build_seed_table_range(ix._h, 128, 9, s - 1))
        tmp = np.zeros((m, 10), np.uint32)
        _lib.check(lib.ngtgpu_select_seeds(ix._h, np.ascontiguousarray(base[s - 1:s - 1 + m]).ctypes.data, _lib.OBJECT_FLOAT, m, 10,
                                           tmp.ctypes.data))
        seeds[s - 1:s - 1 + m] = tmp
    g = build.insert_objects(ix, 1, n_first, None, e, 0.1, -1, batch, 10, 128, 9)
    g2 = build.insert_objects(ix, n_first + 1, n - n_first, g, e, 0.1, -1, batch, 10, 128, 9)
    pobj = po.pad_objects(base, po.FLOAT)
    rows_int = np.vstack([np.zeros((1, base.shape[1]), np.int64), base.astype(np.int64)])
    ref1 = po.build_anng_loop(port, pobj, rows_int, seeds, 1, n_first, None, e, 0.1, edge_size, batch)
    ref = po.build_anng_loop(port, pobj, rows_int, seeds, n_first + 1, n - n_first, ref1, e, 0.1, edge_size, batch)
    got = _lists(g2[0].cpu().numpy(), g2[1].cpu().numpy().astype(np.uint32), g2[2].cpu().numpy())
    assert got == ref
    # the graph is navigable: (nearly) every object finds itself
    ix.build_seed_table(128, 1)
    r = ix.search(base[:200], 3, 0.1, edge_size=0)
    assert (r[0][:, 0] == np.arange(1, 201)).mean() >= 0.9
    ix.close()


# ---- the construction path against graphs the UNMODIFIED reference built (tests/golden/anng_build.npz,
# refine_anng.npz; generator make_golden_anng.py): NGT::Index::createIndex's batched loop and refineANNG with the
# reference's SeedTypeFixedNodes, i.e. every search starts from ids 1..seedSize
def _set_fixed_seeds(ix, seed_size):
    import ctypes as C
    from ngt_b200 import _lib
    lib = _lib.load()
    lib.ngtgpu_index_set_seed_table_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    ids = np.arange(1, seed_size + 1, dtype=np.uint32)
    _lib.check(lib.ngtgpu_index_set_seed_table_ids(ix._h, ids.ctypes.data, seed_size))


@pytest.mark.parametrize("tag", ["f_b200", "f_b64_all", "f_b1000_s5", "u8_b200"])
def test_anng_construction_loop_equals_the_reference(eng, tag):
    """ngtgpu_index_insert_batch == createIndex of the reference (Index.cpp:631-719,721-792; Index.h:815-837;
    Graph.h:611-626,845-886), edge for edge with distance bits: build, then insertion into the built graph."""
    from test_oracle_pin import anng_case
    from ngt_b200 import build
    c = anng_case(np.load(os.path.join(GOLDEN, "anng_build.npz")), tag)
    base = c["base"].astype(np.uint8) if c["otype"] == po.UINT8 else c["base"]
    ix = eng.GpuIndex(c["otype"], po.L2, base.shape[1])
    ix.set_objects(base)
    ix.set_search_property(c["es"], 30, 20)
    g = build.insert_objects(ix, 1, c["n_first"], None, c["e"], 0.1, -1, c["bs"], c["ss"], 0, 0)
    if c["n_first"] < c["n"]:
        g = build.insert_objects(ix, c["n_first"] + 1, c["n"] - c["n_first"], g, c["e"], 0.1, -1, c["bs"], c["ss"], 0, 0)
    got = _lists(g[0].cpu().numpy(), g[1].cpu().numpy().astype(np.uint32), g[2].cpu().numpy())
    assert got == c["lists"]
    ix.close()


@pytest.mark.parametrize("tag", ["r0_all", "r0_b400", "r12_b500", "rm6_all", "u8_r0_b500"])
def test_refine_anng_equals_the_reference(eng, tag):
    """ngtgpu_index_refine_anng == GraphReconstructor::refineANNG of the reference (GraphReconstructor.h:814-924)."""
    import torch
    from test_oracle_pin import anng_case
    from ngt_b200 import build
    zb = np.load(os.path.join(GOLDEN, "anng_build.npz"))
    zr = np.load(os.path.join(GOLDEN, "refine_anng.npz"))
    src = str(zr[tag + "_src"][0])
    c = anng_case(zb, src)
    noe, explore, bs = [int(v) for v in zr[tag + "_meta"]]
    n = c["n"]
    base = c["base"].astype(np.uint8) if c["otype"] == po.UINT8 else c["base"]
    ix = eng.GpuIndex(c["otype"], po.L2, base.shape[1])
    ix.set_objects(base)
    ix.set_search_property(c["es"], 30, 20)
    dev = torch.device("cuda", 0)
    rp = torch.from_numpy(zb[src + "_row_ptr"][:n + 2].astype(np.int64)).to(dev)
    col = torch.from_numpy(zb[src + "_col"].astype(np.int32)).to(dev)
    dist = torch.from_numpy(zb[src + "_dist"]).to(dev)
    ix.set_graph(rp, col)
    _set_fixed_seeds(ix, c["ss"])
    out = build.refine_anng(ix, rp, col, dist, epsilon=float(zr[tag + "_eps"][0]), no_of_edges=noe, edge_size=-1, batch_size=bs,
                            edge_size_for_creation=c["e"], n_seeds=c["ss"])
    got = _lists(out[0].cpu().numpy(), out[1].cpu().numpy().astype(np.uint32), out[2].cpu().numpy())
    assert got == _lists(zr[tag + "_row_ptr"], zr[tag + "_col"], zr[tag + "_dist"])[:n + 1]
    ix.close()


def test_refine_anng_is_reproducible_at_size(eng):
    """100k objects, result lists of 40 keys (four per lane of the lean kernel's control warp): refineANNG gives the same
    graph call after call and on both traversal kernels. Guards the ordering of ngtgpu_index_set_graph's device-to-device
    copies with its head-table kernel: on the legacy stream they were not ordered with the index's stream, and at this
    size the head table of one batch could hold edges of the batch before (one call in three differed)."""
    import torch
    from ngt_b200 import build, synth
    dev = torch.device("cuda", 0)
    base = synth.make_device("sift", 100000, 1, dev)
    ix = eng.GpuIndex(po.FLOAT, po.L2, base.shape[1])
    ix.set_objects(base)
    g = ix.build_onng(64, 10, 64, True, want_graph=True)
    ix.build_seed_table(256, 1)
    rp, col, dist = g["graph"]
    runs = []
    for fast in (True, True, True, False):
        ix.set_fast_kernel(fast)
        ix.set_graph(rp, col)
        runs.append(build.refine_anng(ix, rp, col, dist, 0.1, 0, -1, 10000, 40, 10))
    for r in runs[1:]:
        for x, y in zip(r, runs[0]):
            assert x.shape == y.shape and torch.equal(x, y)
    assert runs[0][1].numel() > col.numel()
    ix.set_fast_kernel(True)
    ix.close()


def test_insert_batch_merge_equals_the_full_sort(eng):
    """ngtgpu_index_insert_batch merges a batch's sorted edge triples into the lists of the graph; lists that do not come
    in (distance, target) order send the batch through the full sort of every edge instead, which repairs them. Both
    routes must leave the same graph: the same second half inserted into a graph given sorted (merge from the first
    batch on) and given with every list reversed (full sort first; no edge cap, so the searches see the same sets)."""
    import torch
    from ngt_b200 import build, synth
    n, half = 3000, 2000
    base = synth.make("sift", n, 21)
    ix = eng.GpuIndex(po.FLOAT, po.L2, base.shape[1])
    ix.set_objects(base)
    g1 = build.insert_objects(ix, 1, half, None, 8, 0.1, 0, 200, 10, 0)
    ref = build.insert_objects(ix, half + 1, n - half, g1, 8, 0.1, 0, 200, 10, 0)
    rp, col, dist = [t.cpu().numpy() for t in g1]
    rcol, rdist = col.copy(), dist.copy()
    for i in range(len(rp) - 1):
        b, e = int(rp[i]), int(rp[i + 1])
        rcol[b:e], rdist[b:e] = col[b:e][::-1], dist[b:e][::-1]
    assert (rcol != col).any()
    dev = g1[1].device
    g1r = (g1[0], torch.from_numpy(rcol).to(dev), torch.from_numpy(rdist).to(dev))
    got = build.insert_objects(ix, half + 1, n - half, g1r, 8, 0.1, 0, 200, 10, 0)
    for x, y in zip(got, ref):
        assert x.shape == y.shape and torch.equal(x, y)
    assert ref[1].numel() > col.size
    ix.close()

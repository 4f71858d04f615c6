"""NGT's C API (lib/NGT/Capi.h) served by libngtgpu.so, driven the way python/ngt/base.py drives libngt:
open an index the REFERENCE built (tests/golden/idx200: prf/obj/grp/tre written by `ngt create`), search it, and
the create -> append -> build -> save -> reopen cycle; answers compared with the reference's own C API on the same
index (tests/golden/idx200_answers.json)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import capi
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from ngt_b200 import _lib
    return capi.bind(_lib.SO_PATH)


def test_open_reference_built_index_and_search(lib, sift5k):
    ans = json.load(open(os.path.join(GOLDEN, "idx200_answers.json")))
    err = lib.ngt_create_error_object()
    ix = lib.ngt_open_index(os.path.join(GOLDEN, "idx200").encode(), err)
    assert ix, lib.ngt_get_error_string(err)
    prop = lib.ngt_create_property(err)
    assert lib.ngt_get_property(ix, prop, err)
    assert lib.ngt_get_property_dimension(prop, err) == 128 and lib.ngt_get_property_distance_type(prop, err) == 1
    assert lib.ngt_is_property_object_type_float(lib.ngt_get_property_object_type(prop, err))
    assert lib.ngt_get_object_repository_size(ix, err) == 201
    qs = sift5k["queries"].astype(np.float32)
    for qi, q in enumerate(qs):
        q = np.ascontiguousarray(q)
        r = lib.ngt_create_empty_results(err)
        assert lib.ngt_linear_search_index_as_float(ix, capi.fptr(q), 128, 5, r, err), lib.ngt_get_error_string(err)
        got = capi.results_of(lib, r, err)
        assert [g[0] for g in got] == [a[0] for a in ans["linear"][qi]]
        assert [np.float32(g[1]) for g in got] == [np.float32(a[1]) for a in ans["linear"][qi]]     # bit-exact: integer-valued data
        # graph search (double* entry point, as base.py calls it): on 200 objects at epsilon 0.3 both find the exact answer
        qd = q.astype(np.float64)
        assert lib.ngt_search_index(ix, qd.ctypes.data_as(C.POINTER(C.c_double)), 128, 5, 0.3, -1.0, r, err)
        assert capi.results_of(lib, r, err) == [(a[0], float(np.float32(a[1]))) for a in ans["graph"][qi]]
        # NGTQuery form
        nq = capi.Query(capi.fptr(q), 5, 0.3, -1.0, -1.0, C.c_size_t(-2**31 & (2**64 - 1)))
        assert lib.ngt_search_index_with_query(ix, nq, r, err), lib.ngt_get_error_string(err)
        assert [g[0] for g in capi.results_of(lib, r, err)] == [a[0] for a in ans["graph"][qi]]
        lib.ngt_destroy_results(r)
    # stored objects and edges are what the reference wrote
    sp = lib.ngt_get_object_space(ix, err)
    row = np.ctypeslib.as_array(lib.ngt_get_object_as_float(sp, 7, err), shape=(128,))
    assert (row == sift5k["data"][6].astype(np.float32)).all()
    e = lib.ngt_create_empty_results(err)
    assert lib.ngt_get_edges(ix, 7, e, err)
    edges = capi.results_of(lib, e, err)
    assert len(edges) >= 10 and all(edges[i][1] <= edges[i + 1][1] for i in range(len(edges) - 1))
    # misuse: wrong dimension -> false + message, nothing crashes
    r = lib.ngt_create_empty_results(err)
    assert lib.ngt_search_index_as_float(ix, capi.fptr(qs[0]), 64, 5, 0.1, -1.0, r, err) is False
    assert lib.ngt_get_error_string(err).decode().startswith("Capi : ngt_search_index_as_float() : Error: ")
    # batch entry point == the single-query answers
    ids = np.zeros((3, 5), np.uint32)
    ds = np.zeros((3, 5), np.float32)
    cnt = np.zeros(3, np.uint32)
    q3 = np.ascontiguousarray(qs)
    assert lib.ngt_batch_linear_search_index_as_float(ix, capi.fptr(q3), 3, 128, 5, -1.0, ids.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                      capi.fptr(ds), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), err)
    assert (cnt == 5).all() and ids.tolist() == [[a[0] for a in ans["linear"][i]] for i in range(3)]
    assert lib.ngt_batch_search_index_as_float(ix, capi.fptr(q3), 3, 128, 5, 0.3, -1.0, -1, ids.ctypes.data_as(C.POINTER(C.c_uint32)),
                                               capi.fptr(ds), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), err)
    assert ids.tolist() == [[a[0] for a in ans["graph"][i]] for i in range(3)]
    lib.ngt_close_index(ix)
    lib.ngt_destroy_property(prop)
    lib.ngt_destroy_error_object(err)


def test_create_append_build_save_reopen(lib, tmp_path):
    from ngt_b200 import synth
    err = lib.ngt_create_error_object()
    prop = lib.ngt_create_property(err)
    assert lib.ngt_set_property_dimension(prop, 128, err) and lib.ngt_set_property_edge_size_for_creation(prop, 12, err)
    assert lib.ngt_set_property_object_type_float(prop, err) and lib.ngt_set_property_distance_type_l2(prop, err)
    path = str(tmp_path / "idx").encode()
    ix = lib.ngt_create_graph_and_tree(path, prop, err)
    assert ix, lib.ngt_get_error_string(err)
    base = synth.make("sift", 3000, 1)
    assert lib.ngt_batch_append_index(ix, capi.fptr(base[:2999]), 2999, err)
    one = base[2999].astype(np.float64)
    assert lib.ngt_insert_index(ix, one.ctypes.data_as(C.POINTER(C.c_double)), 128, err) == 3000
    r = lib.ngt_create_empty_results(err)
    q = synth.make("sift", 2, 2)
    assert lib.ngt_search_index_as_float(ix, capi.fptr(q[0]), 128, 5, 0.1, -1.0, r, err) is False     # not built yet
    assert b"ngt_create_index" in lib.ngt_get_error_string(err)
    assert lib.ngt_create_index(ix, 8, err), lib.ngt_get_error_string(err)
    d = np.linalg.norm(base - q[0], axis=1)
    exact = (np.argsort(d, kind="stable")[:5] + 1).tolist()
    assert lib.ngt_linear_search_index_as_float(ix, capi.fptr(q[0]), 128, 5, r, err)
    assert [g[0] for g in capi.results_of(lib, r, err)] == exact
    assert lib.ngt_search_index_as_float(ix, capi.fptr(q[0]), 128, 5, 0.3, -1.0, r, err)
    assert [g[0] for g in capi.results_of(lib, r, err)] == exact
    assert lib.ngt_remove_index(ix, exact[0], err)
    assert lib.ngt_search_index_as_float(ix, capi.fptr(q[0]), 128, 5, 0.3, -1.0, r, err)
    assert exact[0] not in [g[0] for g in capi.results_of(lib, r, err)]
    assert lib.ngt_save_index(ix, path, err), lib.ngt_get_error_string(err)
    lib.ngt_close_index(ix)
    again = lib.ngt_open_index(path, err)
    assert again, lib.ngt_get_error_string(err)
    assert lib.ngt_search_index_as_float(again, capi.fptr(q[0]), 128, 5, 0.3, -1.0, r, err)
    assert [g[0] for g in capi.results_of(lib, r, err)] == [i for i in (np.argsort(d, kind="stable")[:6] + 1).tolist() if i != exact[0]]
    lib.ngt_close_index(again)
    # the files are NGT's own format: the reference CLI opens them when it is available on this box
    ref_ngt = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "ngt")
    if os.path.exists(ref_ngt):
        import subprocess
        qf = str(tmp_path / "q.tsv")
        np.savetxt(qf, q[:1], fmt="%g", delimiter="\t")
        out = subprocess.run([ref_ngt, "search", "-i", "s", "-n", "5", path.decode(), qf], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    lib.ngt_destroy_results(r)
    lib.ngt_destroy_property(prop)
    lib.ngt_destroy_error_object(err)


def _edges(lib, ix, n, err):
    r = lib.ngt_create_empty_results(err)
    out = []
    for i in range(1, n + 1):
        assert lib.ngt_get_edges(ix, i, r, err), lib.ngt_get_error_string(err)
        out.append(capi.results_of(lib, r, err))
    lib.ngt_destroy_results(r)
    return out


def test_optimizer_execute_equals_the_reference_onng(lib, tmp_path):
    """ngt_optimizer_execute (GraphOptimizer::execute: reconstructGraph + path adjustment) on an ANNG written in NGT's
    own format from the reference's adjacency lists (tests/golden/reconstruct.npz) == the ONNG the reference wrote
    (tests/golden/adjust_paths.npz), edge for edge."""
    from ngt_b200 import index_io, synth
    z = np.load(os.path.join(GOLDEN, "reconstruct.npz"))
    a = np.load(os.path.join(GOLDEN, "adjust_paths.npz"))
    base = synth.make("sift", 1500, 1)
    src = str(tmp_path / "anng")
    os.makedirs(src)
    prf = dict(index_io.DEFAULT_PRF, Dimension="128", EdgeSizeForCreation="20", EdgeSizeForSearch="0")
    index_io.write_prf(src, prf)
    index_io.write_objects(src, base)
    index_io.write_graph(src, z["anng_row_ptr"].astype(np.uint64), z["anng_col"], z["anng_dist"])
    err = lib.ngt_create_error_object()
    for o, i, shortcut, key in ((5, 20, True, "sift_o5_i20_adj"), (10, 40, True, "sift_o10_i40_adj")):
        dst = str(tmp_path / ("onng_%d_%d" % (o, i)))
        opt = lib.ngt_create_optimizer(True, err)
        assert lib.ngt_optimizer_set_minimum(opt, o, i, -1, -1, err)
        assert lib.ngt_optimizer_set_processing_modes(opt, False, False, False, err)
        assert lib.ngt_optimizer_execute(opt, src.encode(), dst.encode(), err), lib.ngt_get_error_string(err)
        lib.ngt_destroy_optimizer(opt)
        ix = lib.ngt_open_index(dst.encode(), err)
        assert ix, lib.ngt_get_error_string(err)
        got = _edges(lib, ix, 1500, err)
        rp, col, dist = a[key + "_row_ptr"], a[key + "_col"], a[key + "_dist"]
        for nid in range(1, 1501):
            ref = [(int(col[e]), float(dist[e])) for e in range(int(rp[nid]), int(rp[nid + 1]))]
            assert got[nid - 1] == ref, (key, nid)
        lib.ngt_close_index(ix)
        assert "ONNG" in open(os.path.join(dst, "prf")).read()
    lib.ngt_destroy_error_object(err)


def test_refine_anng_through_the_c_api(lib, tmp_path):
    from ngt_b200 import synth
    err = lib.ngt_create_error_object()
    prop = lib.ngt_create_property(err)
    assert lib.ngt_set_property_dimension(prop, 128, err) and lib.ngt_set_property_edge_size_for_creation(prop, 6, err)
    ix = lib.ngt_create_graph_and_tree_in_memory(prop, err)
    base = synth.make("sift", 2000, 4)
    assert lib.ngt_batch_append_index(ix, capi.fptr(base), 2000, err) and lib.ngt_create_index(ix, 4, err)
    before = sum(len(e) for e in _edges(lib, ix, 2000, err))
    assert lib.ngt_refine_anng(ix, 0.1, 0.0, 0, -2147483648, 500, err), lib.ngt_get_error_string(err)
    assert sum(len(e) for e in _edges(lib, ix, 2000, err)) >= before     # the exact 6-NN closure has little left to find
    assert lib.ngt_refine_anng(ix, 0.1, 0.0, -12, -2147483648, 800, err), lib.ngt_get_error_string(err)   # search 12 edges
    lists = _edges(lib, ix, 2000, err)
    assert sum(len(e) for e in lists) > before and min(len(e) for e in lists) >= 11     # 12 results, the object itself among them
    for nid, l in enumerate(lists, 1):       # sorted by (distance, id), no self loops, no repeated ids
        assert all(t != nid for t, _ in l) and len({t for t, _ in l}) == len(l)
        assert [(d, t) for t, d in l] == sorted((d, t) for t, d in l)
    assert lib.ngt_refine_anng(ix, 0.1, 0.5, 0, 0, 500, err) is False and b"accuracy table" in lib.ngt_get_error_string(err)
    assert lib.ngt_refine_anng(ix, 0.1, 0.0, 4, -2147483648, 500, err), lib.ngt_get_error_string(err)     # prune to a 4-NN graph
    assert max(len(e) for e in _edges(lib, ix, 2000, err)) == 4
    lib.ngt_close_index(ix)
    lib.ngt_destroy_property(prop)
    lib.ngt_destroy_error_object(err)


def test_insert_into_a_built_index_is_incremental(lib):
    """ngt_insert_index* + ngt_create_index on an index that already has its graph inserts the new objects the way the
    reference's construction loop does (search on the frozen graph, reverse edges): old edges stay, new nodes are linked
    and found."""
    from ngt_b200 import synth
    err = lib.ngt_create_error_object()
    prop = lib.ngt_create_property(err)
    assert lib.ngt_set_property_dimension(prop, 128, err) and lib.ngt_set_property_edge_size_for_creation(prop, 10, err)
    ix = lib.ngt_create_graph_and_tree_in_memory(prop, err)
    base = synth.make("sift", 2400, 6)
    assert lib.ngt_batch_append_index(ix, capi.fptr(base[:2000]), 2000, err) and lib.ngt_create_index(ix, 4, err)
    before = _edges(lib, ix, 2000, err)
    ids = np.zeros(400, np.uint32)
    assert lib.ngt_batch_insert_index(ix, capi.fptr(base[2000:]), 400, ids.ctypes.data_as(C.POINTER(C.c_uint32)), err)
    assert ids[0] == 2001 and ids[-1] == 2400
    assert lib.ngt_create_index(ix, 4, err), lib.ngt_get_error_string(err)
    after = _edges(lib, ix, 2400, err)
    for nid in range(2000):
        assert set(before[nid]) <= set(after[nid])                       # nothing was rebuilt
        assert all(t > 2000 for t, _ in set(after[nid]) - set(before[nid]))   # only reverse edges of new nodes were added
    assert all(1 <= len(after[nid]) for nid in range(2000, 2400))
    assert all(len([t for t, _ in after[nid] if t <= 2000]) + len([t for t, _ in after[nid] if t > 2000]) >= 10
               for nid in range(2000, 2400))
    r = lib.ngt_create_empty_results(err)
    found = 0
    for q in range(2000, 2400, 7):
        assert lib.ngt_search_index_as_float(ix, capi.fptr(base[q]), 128, 3, 0.1, -1.0, r, err)
        found += capi.results_of(lib, r, err)[0][0] == q + 1
    assert found >= 0.9 * len(range(2000, 2400, 7))
    lib.ngt_destroy_results(r)
    lib.ngt_close_index(ix)
    lib.ngt_destroy_property(prop)
    lib.ngt_destroy_error_object(err)


def test_reference_base_py_runs_on_libngtgpu(tmp_path):
    """The reference's ctypes binding (python/ngt/base.py, UNMODIFIED) with libngtgpu.so standing where libngt.so would:
    create -> insert_blob -> build -> search -> get_object -> remove -> save -> reopen, answers checked against numpy."""
    from ngt_b200 import _lib, synth
    mod = capi.load_reference_base_py(_lib.SO_PATH)
    if mod is None:
        pytest.skip("no copy of the reference's python/ngt/base.py on this box")
    path = str(tmp_path / "idx").encode()
    base = synth.make("sift", 1200, 3)
    index = mod.Index.create(path, 128, edge_size_for_creation=12)
    index.insert_blob(base.tolist())
    q = synth.make("sift", 3, 4)
    d = np.linalg.norm(base.astype(np.float64) - q[0].astype(np.float64), axis=1)
    exact = (np.argsort(d, kind="stable")[:5] + 1).tolist()
    res = index.search(q[0].tolist(), 5, 0.3)
    assert [o.id for o in res] == exact
    assert [np.float32(o.distance) for o in res] == [np.float32(x) for x in np.sqrt((d[np.array(exact) - 1] ** 2))]
    assert index.get_object(17) == base[16].tolist()
    new_id = index.insert_object(q[1].tolist())            # ngt_insert_index (double*), then ngt_create_index
    assert new_id == 1201
    index.build_index()
    assert index.search(q[1].tolist(), 1, 0.1)[0].id == 1201
    index.remove(exact[0])
    assert exact[0] not in [o.id for o in index.search(q[0].tolist(), 5, 0.3)]
    with pytest.raises(mod.NativeError):
        index.get_object(exact[0])
    index.save()
    again = mod.Index(path)
    assert [o.id for o in again.search(q[0].tolist(), 4, 0.3)] == exact[1:]
    del again, index


def test_expected_accuracy_search_uses_the_accuracy_table(lib, sift5k, tmp_path):
    """NGTQuery.accuracy > 0 (Capi.cpp:346-375 -> Index.h:1156-1158): epsilon comes from the AccuracyTable of `prf`; the
    answers equal a search with the epsilon the REFERENCE derives from the same table (tests/golden/accuracy_table.json).
    An index without a table fails with the reference's message."""
    import shutil
    tab = json.load(open(os.path.join(GOLDEN, "accuracy_table.json")))
    src = os.path.join(GOLDEN, "idx200")
    dst = str(tmp_path / "idx200_acc")
    shutil.copytree(src, dst)
    lines = open(os.path.join(dst, "prf")).read().splitlines()
    open(os.path.join(dst, "prf"), "w").write("\n".join(("AccuracyTable\t" + tab["table"]) if l.startswith("AccuracyTable") else l
                                                       for l in lines) + "\n")
    err = lib.ngt_create_error_object()
    ix = lib.ngt_open_index(dst.encode(), err)
    assert ix, lib.ngt_get_error_string(err)
    q = np.ascontiguousarray(sift5k["queries"][0].astype(np.float32))
    r = lib.ngt_create_empty_results(err)
    for acc, eps in tab["cases"]:
        if acc < 0.5:
            continue
        nq = capi.Query(capi.fptr(q), 10, 9.0, acc, -1.0, C.c_size_t(-2**31 & (2**64 - 1)))    # epsilon 9.0 must be ignored
        assert lib.ngt_search_index_with_query(ix, nq, r, err), lib.ngt_get_error_string(err)
        with_acc = capi.results_of(lib, r, err)
        ne = capi.Query(capi.fptr(q), 10, eps, 0.0, -1.0, C.c_size_t(-2**31 & (2**64 - 1)))
        assert lib.ngt_search_index_with_query(ix, ne, r, err)
        assert with_acc == capi.results_of(lib, r, err) and len(with_acc) == 10, acc
    lib.ngt_close_index(ix)
    plain = lib.ngt_open_index(src.encode(), err)
    nq = capi.Query(capi.fptr(q), 10, 0.1, 0.9, -1.0, C.c_size_t(-2**31 & (2**64 - 1)))
    assert lib.ngt_search_index_with_query(plain, nq, r, err) is False
    assert lib.ngt_get_error_string(err).decode() == \
        "Capi : ngt_search_index_with_query() : Error: AccuracyTable: The accuracy table is not set yet. The table size=0"
    lib.ngt_close_index(plain)
    lib.ngt_destroy_results(r)
    lib.ngt_destroy_error_object(err)


def test_uint8_batch_entry_points(lib, sift5k, tmp_path):
    """ngt_batch_search_index_as_uint8 / ngt_batch_linear_search_index_as_uint8: byte queries give the
    answers of the float entry points on an Integer-1 index; on a Float-4 index they are refused."""
    err = lib.ngt_create_error_object()
    prop = lib.ngt_create_property(err)
    assert lib.ngt_set_property_dimension(prop, 128, err) and lib.ngt_set_property_object_type_integer(prop, err)
    ix = lib.ngt_create_graph_and_tree_in_memory(prop, err)
    base = np.ascontiguousarray(sift5k["data"][:3000].astype(np.float32))
    assert lib.ngt_batch_append_index(ix, capi.fptr(base), 3000, err) and lib.ngt_create_index(ix, 4, err)
    qf = np.ascontiguousarray(sift5k["queries"].astype(np.float32))
    qb = np.ascontiguousarray(qf.astype(np.uint8))
    u32, u8 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    out = {}
    for tag in ("f", "b"):
        ids, ds, cnt = np.zeros((3, 7), np.uint32), np.zeros((3, 7), np.float32), np.zeros(3, np.uint32)
        lids, lds, lcnt = np.zeros((3, 7), np.uint32), np.zeros((3, 7), np.float32), np.zeros(3, np.uint32)
        if tag == "f":
            assert lib.ngt_batch_search_index_as_float(ix, capi.fptr(qf), 3, 128, 7, 0.2, -1.0, -1, ids.ctypes.data_as(u32),
                                                       capi.fptr(ds), cnt.ctypes.data_as(u32), err)
            assert lib.ngt_batch_linear_search_index_as_float(ix, capi.fptr(qf), 3, 128, 7, -1.0, lids.ctypes.data_as(u32),
                                                              capi.fptr(lds), lcnt.ctypes.data_as(u32), err)
        else:
            assert lib.ngt_batch_search_index_as_uint8(ix, qb.ctypes.data_as(u8), 3, 128, 7, 0.2, -1.0, -1, ids.ctypes.data_as(u32),
                                                       capi.fptr(ds), cnt.ctypes.data_as(u32), err), lib.ngt_get_error_string(err)
            assert lib.ngt_batch_linear_search_index_as_uint8(ix, qb.ctypes.data_as(u8), 3, 128, 7, -1.0, lids.ctypes.data_as(u32),
                                                              capi.fptr(lds), lcnt.ctypes.data_as(u32), err)
        out[tag] = (ids, ds, cnt, lids, lds, lcnt)
    for a, b in zip(out["f"], out["b"]):
        assert (a.view(np.uint32) == b.view(np.uint32)).all()
    # exhaustive answers == the integer formula
    d2 = ((base[None, :, :].astype(np.int64) - qb[:, None, :].astype(np.int64)) ** 2).sum(-1)
    assert (out["b"][3] == np.argsort(d2, axis=1, kind="stable")[:, :7] + 1).all()
    lib.ngt_close_index(ix)
    fprop = lib.ngt_create_property(err)
    assert lib.ngt_set_property_dimension(fprop, 128, err)
    fx = lib.ngt_create_graph_and_tree_in_memory(fprop, err)
    assert lib.ngt_batch_append_index(fx, capi.fptr(base), 100, err) and lib.ngt_create_index(fx, 4, err)
    ids, ds, cnt = np.zeros((3, 7), np.uint32), np.zeros((3, 7), np.float32), np.zeros(3, np.uint32)
    assert lib.ngt_batch_search_index_as_uint8(fx, qb.ctypes.data_as(u8), 3, 128, 7, 0.2, -1.0, -1, ids.ctypes.data_as(u32),
                                               capi.fptr(ds), cnt.ctypes.data_as(u32), err) is False
    assert b"not integer" in lib.ngt_get_error_string(err)
    lib.ngt_close_index(fx)
    lib.ngt_destroy_property(prop)
    lib.ngt_destroy_property(fprop)
    lib.ngt_destroy_error_object(err)


def test_create_index_with_nothing_queued_keeps_the_graph(lib, tmp_path):
    """ngt_create_index on an opened ONNG with nothing appended is a no-op (createIndex only indexes objects that are
    not in the graph, Index.cpp:645-648): the optimised graph and GraphType stay. Removal while objects are queued
    strips the id from the old graph, and the queued objects are then inserted incrementally."""
    from ngt_b200 import index_io, synth
    z = np.load(os.path.join(GOLDEN, "reconstruct.npz"))
    base = synth.make("sift", 1500, 1)
    src = str(tmp_path / "onng")
    os.makedirs(src)
    index_io.write_prf(src, dict(index_io.DEFAULT_PRF, Dimension="128", EdgeSizeForCreation="20", EdgeSizeForSearch="0", GraphType="ONNG"))
    index_io.write_objects(src, base)
    index_io.write_graph(src, z["o10_i40_row_ptr"].astype(np.uint64), z["o10_i40_col"], z["o10_i40_dist"])
    err = lib.ngt_create_error_object()
    ix = lib.ngt_open_index(src.encode(), err)
    assert ix, lib.ngt_get_error_string(err)
    before = _edges(lib, ix, 1500, err)
    assert lib.ngt_create_index(ix, 4, err)
    assert _edges(lib, ix, 1500, err) == before
    more = synth.make("sift", 300, 9)
    assert lib.ngt_batch_append_index(ix, capi.fptr(more), 300, err)
    assert lib.ngt_remove_index(ix, 7, err), lib.ngt_get_error_string(err)          # while 300 objects are queued
    assert lib.ngt_create_index(ix, 4, err), lib.ngt_get_error_string(err)
    after = _edges(lib, ix, 1800, err)
    assert after[6] == [] and all(7 not in [t for t, _ in l] for l in after)
    for nid in range(1500):
        if nid != 6:
            assert set(e for e in before[nid] if e[0] != 7) <= set(after[nid])     # the old graph was kept, not rebuilt
    assert all(len(after[nid]) >= 1 for nid in range(1500, 1800))
    out = str(tmp_path / "saved")
    assert lib.ngt_save_index(ix, out.encode(), err), lib.ngt_get_error_string(err)
    assert "ONNG" in open(os.path.join(out, "prf")).read()
    lib.ngt_close_index(ix)
    lib.ngt_destroy_error_object(err)


def test_concurrent_searches_on_one_handle(lib, sift5k):
    """The reference's threading contract: many threads search one read-only index at once. Every thread
    gets the answers of the same calls made one after another (each concurrent call runs on a lane of its own)."""
    import threading
    err = lib.ngt_create_error_object()
    prop = lib.ngt_create_property(err)
    assert lib.ngt_set_property_dimension(prop, 128, err)
    ix = lib.ngt_create_graph_and_tree_in_memory(prop, err)
    base = np.ascontiguousarray(sift5k["data"].astype(np.float32))
    assert lib.ngt_batch_append_index(ix, capi.fptr(base), base.shape[0], err) and lib.ngt_create_index(ix, 4, err)
    u32 = C.POINTER(C.c_uint32)
    rng = np.random.default_rng(5)
    batches = [np.ascontiguousarray(base[rng.integers(0, base.shape[0], 700)] + rng.integers(0, 3, (700, 128)).astype(np.float32))
               for _ in range(6)]

    def run(b, e):
        ids, ds, cnt = np.zeros((700, 8), np.uint32), np.zeros((700, 8), np.float32), np.zeros(700, np.uint32)
        ok = lib.ngt_batch_search_index_as_float(ix, capi.fptr(b), 700, 128, 8, 0.1, -1.0, -1, ids.ctypes.data_as(u32), capi.fptr(ds),
                                                 cnt.ctypes.data_as(u32), e)
        lids, lds, lcnt = np.zeros((700, 8), np.uint32), np.zeros((700, 8), np.float32), np.zeros(700, np.uint32)
        ok = ok and lib.ngt_batch_linear_search_index_as_float(ix, capi.fptr(b), 700, 128, 8, -1.0, lids.ctypes.data_as(u32),
                                                               capi.fptr(lds), lcnt.ctypes.data_as(u32), e)
        return ok, ids, ds, cnt, lids, lds

    serial = [run(b, err) for b in batches]
    assert all(s[0] for s in serial)
    out = [None] * 24

    def worker(t):
        e = lib.ngt_create_error_object()
        for rep in range(4):
            j = (t + rep) % len(batches)
            out[t * 4 + rep] = (j, run(batches[j], e))
        lib.ngt_destroy_error_object(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for j, got in out:
        assert got[0]
        for a, b in zip(got[1:], serial[j][1:]):
            assert (a.view(np.uint32) == b.view(np.uint32)).all()
    lib.ngt_close_index(ix)
    lib.ngt_destroy_property(prop)
    lib.ngt_destroy_error_object(err)


def test_remove_repairs_edges_like_the_reference(lib, tmp_path):
    """ngt_remove_index == NGT::Index::remove of the reference (removeEdgesReliably, Graph.cpp:641-864): the ANNG the
    reference built (tests/golden/anng_build.npz), written in NGT's file format, opened through the C API, the same nine
    removals: every list equals the reference's (tests/golden/remove.npz). The Python mirror does the same."""
    from ngt_b200 import index as ngt
    from ngt_b200 import index_io, synth
    zb = np.load(os.path.join(GOLDEN, "anng_build.npz"))
    zr = np.load(os.path.join(GOLDEN, "remove.npz"))
    objtype, n, n_first, seed, e, es, ss, bs = [int(v) for v in zb["f_b200_meta"]]
    base = synth.make("sift", n, seed)
    src = str(tmp_path / "anng")
    os.makedirs(src)
    index_io.write_prf(src, dict(index_io.DEFAULT_PRF, Dimension="128", EdgeSizeForCreation=str(e), EdgeSizeForSearch=str(es)))
    index_io.write_objects(src, base)
    index_io.write_graph(src, zb["f_b200_row_ptr"][:n + 2].astype(np.uint64), zb["f_b200_col"], zb["f_b200_dist"])
    rp, col, dist = zr["row_ptr"], zr["col"], zr["dist"]
    want = [[(int(col[x]), float(dist[x])) for x in range(int(rp[i]), int(rp[i + 1]))] for i in range(1, n + 1)]
    err = lib.ngt_create_error_object()
    ix = lib.ngt_open_index(src.encode(), err)
    assert ix, lib.ngt_get_error_string(err)
    for rid in zr["removed"]:
        assert lib.ngt_remove_index(ix, int(rid), err), lib.ngt_get_error_string(err)
    assert _edges(lib, ix, n, err) == want
    assert lib.ngt_remove_index(ix, int(zr["removed"][0]), err) is False        # already gone
    lib.ngt_close_index(ix)
    lib.ngt_destroy_error_object(err)
    m = ngt.Index(src, zero_based_numbering=False)
    for rid in zr["removed"]:
        m.remove(int(rid))
    mrp, mcol, mdist = m._graph
    got = [[(int(mcol[x]), float(mdist[x])) for x in range(int(mrp[i]), int(mrp[i + 1]))] for i in range(1, n + 1)]
    assert got == want
    m.close()


def test_optimizer_execute_generates_the_accuracy_table(lib, tmp_path):
    """ngt_optimizer_execute with the accuracy-table step on (GraphOptimizer.h:355-368 -> Optimizer::generateAccuracyTable,
    Optimizer.h:1494-1573: every search of its loops is one batch call here) on the ANNG the reference built
    (anng_build.npz, f_b64_all), against the table the reference generated from the same ANNG with the same settings
    (tests/golden/accuracy_table.json "generated"). The graphs are identical; the seeds differ (device pivots vs the
    reference's fixed nodes), so the tables agree closely where accuracy is high and only in shape below."""
    from ngt_b200 import index_io, synth
    from ngt_b200.index import epsilon_from_accuracy_table
    gen = json.load(open(os.path.join(GOLDEN, "accuracy_table.json")))["generated"]
    zb = np.load(os.path.join(GOLDEN, "anng_build.npz"))
    objtype, n, n_first, seed, e, es, ss, bs = [int(v) for v in zb[gen["case"] + "_meta"]]
    src, dst = str(tmp_path / "anng"), str(tmp_path / "onng")
    os.makedirs(src)
    index_io.write_prf(src, dict(index_io.DEFAULT_PRF, Dimension="128", EdgeSizeForCreation=str(e), EdgeSizeForSearch=str(es)))
    index_io.write_objects(src, synth.make("sift", n, seed))
    index_io.write_graph(src, zb[gen["case"] + "_row_ptr"][:n + 2].astype(np.uint64), zb[gen["case"] + "_col"], zb[gen["case"] + "_dist"])
    err = lib.ngt_create_error_object()
    opt = lib.ngt_create_optimizer(True, err)
    assert lib.ngt_optimizer_set_minimum(opt, gen["outgoing"], gen["incoming"], gen["queries"], gen["results"], err)
    assert lib.ngt_optimizer_set_processing_modes(opt, False, False, True, err)
    assert lib.ngt_optimizer_execute(opt, src.encode(), dst.encode(), err), lib.ngt_get_error_string(err)
    lib.ngt_destroy_optimizer(opt)
    table = index_io.read_prf(dst)["AccuracyTable"]
    pts = [(float(t.split(":")[0]), float(t.split(":")[1])) for t in table.split(",")]
    assert len(pts) >= 8 and abs(pts[0][0] + 0.6) < 1e-6                  # starts at epsilon -0.6 like the reference's sweep
    assert all(b[0] > a[0] and b[1] > a[1] for a, b in zip(pts, pts[1:]))      # strictly increasing in both
    assert pts[-1][1] > 0.98
    for acc in (0.9, 0.95, 0.98):
        assert abs(epsilon_from_accuracy_table(table, acc) - epsilon_from_accuracy_table(gen["table"], acc)) <= 0.05, acc
    # the table is used: an expected-accuracy search through the C API on the written index
    ix = lib.ngt_open_index(dst.encode(), err)
    assert ix, lib.ngt_get_error_string(err)
    q = np.ascontiguousarray(synth.make("sift", 1, 3)[0])
    r = lib.ngt_create_empty_results(err)
    nq = capi.Query(capi.fptr(q), 10, 0.0, 0.95, -1.0, C.c_size_t(-2**31 & (2**64 - 1)))
    assert lib.ngt_search_index_with_query(ix, nq, r, err), lib.ngt_get_error_string(err)
    assert len(capi.results_of(lib, r, err)) == 10
    lib.ngt_destroy_results(r)
    lib.ngt_close_index(ix)
    # an index on a fixed edge cap keeps its empty table (the reference gets to -2 through its timed tuning, not run here)
    index_io.write_prf(src, dict(index_io.DEFAULT_PRF, Dimension="128", EdgeSizeForCreation=str(e), EdgeSizeForSearch="40"))
    dst2 = str(tmp_path / "onng40")
    opt = lib.ngt_create_optimizer(True, err)
    assert lib.ngt_optimizer_execute(opt, src.encode(), dst2.encode(), err), lib.ngt_get_error_string(err)
    lib.ngt_destroy_optimizer(opt)
    assert index_io.read_prf(dst2)["AccuracyTable"] == ""
    lib.ngt_destroy_error_object(err)

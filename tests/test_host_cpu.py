"""CPU-only checks of the host side: the C-ABI library loads and exports what include/ngtgpu.h declares,
index files round-trip and interoperate with the reference's own files, graph reconstruction follows
GraphReconstructor::reconstructGraph. No device compute happens here."""
import ctypes as C
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT

REF_NGT = os.path.join(ROOT, "oracle", "_ref", "ngt")


def test_c_abi_exports_every_declared_symbol():
    from ngt_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ngtgpu.h")).read()
    declared = set(re.findall(r"\b(ngtgpu_[a-z0-9_]+)\s*\(", header))
    declared -= {"ngtgpu_index", "ngtgpu_search_params"}
    assert len(declared) >= 20
    lib = C.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libngtgpu.so does not export %s" % name
    # the ctypes table covers the hot-path entry points
    for name in ("ngtgpu_search", "ngtgpu_search_device", "ngtgpu_linear_search", "ngtgpu_linear_search_device",
                 "ngtgpu_index_create", "ngtgpu_index_set_objects", "ngtgpu_index_set_graph"):
        assert name in _lib.SYMBOLS and name in declared


def test_public_headers_are_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/*.h compile as pedantic C99 and as C++11 (plain pointers and sizes, no
    C++ or torch types in any signature), so a cgo / JNI / ctypes / C++ host can bind them."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "ngtgpu.h"\n#include "ngt_capi_ext.h"\nint main(void) { return 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I" + inc, str(src)],
                ["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I" + inc, "-x", "c++", str(src)]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, " ".join(cmd) + "\n" + r.stderr


def test_additive_capi_header_symbols_are_exported():
    """include/ngt_capi_ext.h: the `ngt_*` entry points the library adds to lib/NGT/Capi.h's 67."""
    from ngt_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ngt_capi_ext.h")).read()
    declared = set(re.findall(r"^(?:bool|NGTIndex|uint64_t)\s+(ngt_[a-z0-9_]+)\(", header, re.M))
    assert len(declared) == 8, sorted(declared)
    lib = C.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libngtgpu.so does not export %s" % name


def test_ngtpy_module_has_the_reference_surface():
    """ngt_b200/ngtpy.*.so (csrc/ngtpy.cpp): every function and method of the reference's pybind11 module
    (python/src/ngtpy.cpp:500-607) is there with the same signature line -- names, keyword arguments, defaults --
    recorded from the reference's own build (tests/golden/ngtpy_surface.json, make_golden_ngtpy.py). QuantizedIndex
    (NGTQG, outside the hot path) exists and refuses loudly. No compute here: this container has no GPU."""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "ngt_b200"))
    try:
        import ngtpy
        from make_golden_ngtpy import surface
    finally:
        del sys.path[:2]
    assert ngtpy.__file__.startswith(os.path.join(ROOT, "ngt_b200"))
    ours = surface(ngtpy)
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ngtpy_surface.json")))
    checked = 0
    for name, line in ref.items():
        if name.startswith("QuantizedIndex.") and name != "QuantizedIndex.__init__":
            continue
        assert ours.get(name) == line, name
        checked += 1
    assert checked == 24
    assert "Index.batch_search" in ours and "Index.batch_linear_search" in ours          # additive
    with pytest.raises(RuntimeError, match="PropertySet::load: Cannot load the property file /nonexistent/idx/prf."):
        ngtpy.Index("/nonexistent/idx")
    with pytest.raises(RuntimeError, match="outside the hot path"):
        ngtpy.QuantizedIndex("/nonexistent/idx")
    with pytest.raises(RuntimeError, match="invalid distance type"):
        ngtpy.create("/tmp/ngtpy-never-made", 8, distance_type="Chebyshev")
    o = ngtpy.Optimizer(log_disabled=True)
    o.set(num_of_outgoings=5, num_of_incomings=20)
    o.set_processing_modes(shortcut_reduction=False)
    with pytest.raises(RuntimeError, match="outside the hot path"):
        o.optimize_search_parameters("/nonexistent/idx")


def test_no_device_is_an_error_not_a_fallback():
    """Without a CUDA device every entry point fails loudly (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ngt_b200 import _lib, engine
    with pytest.raises(_lib.NgtGpuError) as e:
        engine.GpuIndex(_lib.OBJECT_FLOAT, _lib.DISTANCE_L2, 16)
    assert e.value.code == _lib.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ngt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "ngt_oracle" not in src and "libngt_ref" not in src, f


def _loop_reconstruct(ids, dists, o, i):
    """GraphReconstructor.h:425-561 as plain loops."""
    n, k = ids.shape
    lists = [[] for _ in range(n + 1)]
    for a in range(1, n + 1):
        lists[a] = [(float(dists[a - 1, r]), int(ids[a - 1, r])) for r in range(min(o, k))]
    for a in range(1, n + 1):
        for r in range(min(i, k)):
            lists[int(ids[a - 1, r])].append((float(dists[a - 1, r]), a))
    out = []
    for a in range(1, n + 1):
        s = sorted(lists[a])
        dedup, prev = [], 0
        for d, t in s:
            if t == prev:
                continue
            prev = t
            dedup.append((d, t))
        out.append(dedup)
    return out


def test_reconstruct_graph_matches_reference_semantics():
    import torch
    from ngt_b200 import build
    rng = np.random.default_rng(0)
    n, k = 300, 12
    x = rng.integers(0, 50, (n, 8)).astype(np.float32)
    d = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)).astype(np.float32)
    np.fill_diagonal(d, np.inf)
    order = np.lexsort((np.broadcast_to(np.arange(n), (n, n)), d), axis=1)[:, :k]
    ids = (order + 1).astype(np.int32)
    dist = np.take_along_axis(d, order, 1)
    for o, i in ((4, 12), (12, 12), (3, 5)):
        rp, col, dd = build.reconstruct_graph(torch.from_numpy(ids), torch.from_numpy(dist),
                                              torch.full((n,), k, dtype=torch.int32), o, i)
        ref = _loop_reconstruct(ids, dist, o, i)
        rp, col, dd = rp.numpy(), col.numpy(), dd.numpy()
        assert rp[0] == 0 and rp[1] == 0
        for a in range(1, n + 1):
            got = list(zip(dd[rp[a]:rp[a + 1]].tolist(), col[rp[a]:rp[a + 1]].tolist()))
            assert got == ref[a - 1], (o, i, a)


def test_index_files_round_trip_and_interoperate_with_the_reference(sift5k):
    from ngt_b200 import index_io
    tmp = tempfile.mkdtemp(prefix="ngt-io-")
    try:
        # (1) what we write, we read back
        n, dim = 200, 128
        rows = sift5k["data"][:n].astype(np.float32)
        rp = sift5k["row_ptr"].astype(np.uint64)
        # sub-graph over the first 200 ids
        lists = []
        for a in range(1, n + 1):
            nb = sift5k["col"][int(rp[a]):int(rp[a + 1])]
            nb = nb[nb <= n]
            dd = np.linalg.norm(rows[nb - 1] - rows[a - 1], axis=1).astype(np.float32)
            o = np.lexsort((nb, dd))
            lists.append((nb[o], dd[o]))
        row_ptr = np.zeros(n + 2, np.uint64)
        row_ptr[2:] = np.cumsum([len(l[0]) for l in lists])
        col = np.concatenate([l[0] for l in lists]).astype(np.uint32)
        dist = np.concatenate([l[1] for l in lists]).astype(np.float32)
        mine = os.path.join(tmp, "mine")
        os.makedirs(mine)
        prop = dict(index_io.DEFAULT_PRF)
        prop.update({"Dimension": str(dim), "ObjectType": "Float-4", "DistanceType": "L2"})
        index_io.write_prf(mine, prop)
        present = np.ones(n + 1, np.uint8)
        present[0] = 0
        present[17] = 0           # a removed slot
        index_io.write_objects(mine, rows, present)
        index_io.write_graph(mine, row_ptr, col, dist, present)
        p2 = index_io.read_prf(mine)
        assert p2 == prop
        r2, pres2 = index_io.read_objects(mine, p2)
        assert (pres2 == present).all()
        keep = present[1:] == 1
        assert (r2[keep] == rows[keep]).all() and (r2[~keep] == 0).all()
        rp2, col2, dist2, gp2 = index_io.read_graph(mine)
        assert (gp2 == present).all()
        deg = np.diff(row_ptr.astype(np.int64))[1:]
        deg2 = np.diff(rp2.astype(np.int64))[1:]
        assert (deg2[keep] == deg[keep]).all() and (deg2[~keep] == 0).all()
        # (2) the reference's own files parse, and ours open in the reference
        if not os.path.exists(REF_NGT):
            pytest.skip("oracle/_ref/ngt not built here")
        tsv = os.path.join(tmp, "d.tsv")
        np.savetxt(tsv, rows, fmt="%g", delimiter="\t")
        theirs = os.path.join(tmp, "theirs")
        subprocess.check_call([REF_NGT, "create", "-d", str(dim), "-o", "f", "-D", "2", theirs, tsv],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        tp = index_io.read_prf(theirs)
        assert tp["ObjectType"] == "Float-4" and tp["DistanceType"] == "L2" and int(tp["Dimension"]) == dim
        tr, tpres = index_io.read_objects(theirs, tp)
        assert (tr == rows).all() and tpres[0] == 0 and (tpres[1:] == 1).all()
        trp, tcol, tdist, _ = index_io.read_graph(theirs)
        assert trp[-1] == tcol.size and tcol.size > n and tcol.min() >= 1 and tcol.max() <= n
        for a in (1, 57, n):       # lists ascending by (distance, id), distances are the true ones
            nb, dd = tcol[int(trp[a]):int(trp[a + 1])], tdist[int(trp[a]):int(trp[a + 1])]
            assert (np.diff(dd) >= 0).all()
            assert np.allclose(dd, np.linalg.norm(rows[nb - 1] - rows[a - 1], axis=1), rtol=1e-6)
        # our files (IndexType Graph, full present set) opened by the reference CLI: exhaustive search agrees
        full = os.path.join(tmp, "full")
        os.makedirs(full)
        index_io.write_prf(full, prop)
        index_io.write_objects(full, rows)
        index_io.write_graph(full, row_ptr, col, dist)
        q = os.path.join(tmp, "q.tsv")
        np.savetxt(q, sift5k["queries"][:1].astype(np.float32), fmt="%g", delimiter="\t")
        out = subprocess.run([REF_NGT, "search", "-i", "s", "-n", "5", full, q], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = [int(m.group(2)) for m in re.finditer(r"^(\d+)\t(\d+)\t", out.stdout, re.M)]
        d = np.linalg.norm(rows - sift5k["queries"][0].astype(np.float32), axis=1)
        assert got == (np.lexsort((np.arange(n), d))[:5] + 1).tolist()
        out = subprocess.run([REF_NGT, "search", "-i", "g", "-n", "5", "-e", "0.3", full, q], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = [int(m.group(2)) for m in re.finditer(r"^(\d+)\t(\d+)\t", out.stdout, re.M)]
        assert len(got) == 5
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_reconstruct_graph_matches_the_reference():
    """GraphReconstructor::reconstructGraph: ANNG built by the reference in, the graph the reference's
    GraphOptimizer::execute writes (path adjustment off) out -- tests/golden/reconstruct.npz."""
    import torch
    from conftest import GOLDEN
    from ngt_b200 import build
    z = np.load(os.path.join(GOLDEN, "reconstruct.npz"))
    rp = torch.from_numpy(z["anng_row_ptr"].astype(np.int64))
    col = torch.from_numpy(z["anng_col"].astype(np.int32))
    dist = torch.from_numpy(z["anng_dist"])
    for o, i in ((5, 20), (10, 40), (0, 15)):
        key = "o%d_i%d" % (o, i)
        orp, ocol, odist = build.reconstruct_graph_csr(rp, col, dist, o, i)
        assert (orp.numpy() == z[key + "_row_ptr"].astype(np.int64)).all(), key
        assert (ocol.numpy().astype(np.uint32) == z[key + "_col"]).all(), key
        assert (odist.numpy().view(np.uint32) == z[key + "_dist"].view(np.uint32)).all(), key


def test_ngt_c_api_symbols_and_error_convention():
    """lib/NGT/Capi.h entry points of the hot path are exported under their own names, and the error convention
    is the reference's (lib/NGT/Capi.cpp:25-38): message into the NGTError string, sentinel return value.
    The expected strings for misuse come from the reference itself (tests/golden/idx200_answers.json)."""
    import json
    import capi
    from conftest import GOLDEN
    from ngt_b200 import _lib
    lib = capi.bind(_lib.SO_PATH)              # AttributeError here = a Capi.h function is missing
    assert len(lib._signatures) >= 50
    err = lib.ngt_create_error_object()
    assert err and lib.ngt_get_error_string(err) == b""
    ref = json.load(open(os.path.join(GOLDEN, "idx200_answers.json")))["errors"]
    # a NULL index: same "parametor error" format as the reference
    res = lib.ngt_create_empty_results(err)
    q = np.zeros(128, np.float32)
    assert lib.ngt_search_index_as_float(None, capi.fptr(q), 128, 5, 0.1, -1.0, res, err) is False
    msg = lib.ngt_get_error_string(err).decode()
    assert msg.startswith("Capi : ngt_search_index_as_float() : parametor error: index = 0 query = ")
    assert ref["null_index"][1].startswith("Capi : ngt_search_index_as_float() : parametor error: index = 0 query = ")
    assert msg.endswith("query_dim = 128") and ref["null_index"][1].endswith("query_dim = 128")
    # a path that does not exist: NULL + "Capi : ngt_open_index() : Error: ... Cannot load the property file <path>/prf."
    assert lib.ngt_open_index(b"/nonexistent/idx", err) is None
    msg = lib.ngt_get_error_string(err).decode()
    assert msg.startswith("Capi : ngt_open_index() : Error: ") and msg.endswith("PropertySet::load: Cannot load the property file /nonexistent/idx/prf.")
    assert ref["bad_path"][1].endswith("PropertySet::load: Cannot load the property file /nonexistent/idx/prf.")
    lib.ngt_clear_error_string(err)
    assert lib.ngt_get_error_string(err) == b""
    # results containers and properties work without a device
    assert lib.ngt_get_result_size(res, err) == 0 and lib.ngt_get_size(res, err) == 0
    o = lib.ngt_get_result(res, 3, err)
    assert (o.id, o.distance) == (0, 0.0) and lib.ngt_get_error_string(err).decode().startswith("Capi : ngt_get_result() : Error: ")
    prop = lib.ngt_create_property(err)
    assert lib.ngt_set_property_dimension(prop, 96, err) and lib.ngt_get_property_dimension(prop, err) == 96
    assert lib.ngt_set_property_edge_size_for_creation(prop, 12, err) and lib.ngt_get_property_edge_size_for_creation(prop, err) == 12
    assert lib.ngt_get_property_edge_size_for_search(prop, err) == 40        # Graph.h:401 default
    assert lib.ngt_set_property_object_type_integer(prop, err)
    assert lib.ngt_is_property_object_type_integer(lib.ngt_get_property_object_type(prop, err))
    assert lib.ngt_set_property_distance_type_hamming(prop, err) and lib.ngt_get_property_distance_type(prop, err) == 2
    assert lib.ngt_get_property_dimension(None, err) == -1
    # refine / optimizer: same parameter-error convention as the reference (Capi.cpp:889-895, 976-983)
    assert lib.ngt_refine_anng(None, 0.1, 0.0, 0, 0, 100, err) is False
    assert lib.ngt_get_error_string(err).decode() == "Capi : ngt_refine_anng() : parametor error: index = 0"
    opt = lib.ngt_create_optimizer(True, err)
    assert opt and lib.ngt_optimizer_set_minimum(opt, 5, 20, -1, -1, err)
    assert lib.ngt_optimizer_execute(None, b"a", b"b", err) is False
    assert lib.ngt_get_error_string(err).decode().startswith("Capi : ngt_optimizer_execute() : parametor error: optimizer = ")
    assert lib.ngt_optimizer_execute(opt, b"/nonexistent/in", b"/tmp", err) is False        # the output exists
    assert lib.ngt_get_error_string(err).decode() == \
        "Capi : ngt_optimizer_execute() : Error: Optimizer::execute: The specified index exists. /tmp"
    # search-parameter tuning is outside the hot path: refused, loudly
    assert lib.ngt_optimizer_adjust_search_coefficients(opt, b"x", err) is False
    assert "not provided by the B200 engine" in lib.ngt_get_error_string(err).decode()
    lib.ngt_destroy_optimizer(opt)
    import torch
    if not torch.cuda.is_available():
        # no device: opening a real index fails with a message, it does not fall back to the CPU
        assert lib.ngt_open_index(os.path.join(GOLDEN, "idx200").encode(), err) is None
        assert "no CUDA device" in lib.ngt_get_error_string(err).decode()
    lib.ngt_destroy_property(prop)
    lib.ngt_destroy_results(res)
    lib.ngt_destroy_error_object(err)


def test_every_capi_h_function_is_exported():
    """All 67 functions lib/NGT/Capi.h:60-212 declares resolve in libngtgpu.so (plus the additive batch entry points),
    and the one struct-returning call works without a device (Capi.cpp:1043-1058 defaults)."""
    import capi
    from ngt_b200 import _lib
    lib = capi.bind(_lib.SO_PATH)
    assert len(capi.CAPI_H_FUNCTIONS) == 67
    for name in capi.CAPI_H_FUNCTIONS:
        assert hasattr(lib, name), name
    for name in ("ngt_batch_search_index_as_float", "ngt_batch_search_index_as_uint8", "ngt_batch_linear_search_index_as_float",
                 "ngt_batch_linear_search_index_as_uint8"):
        assert hasattr(lib, name), name
    p = lib.ngt_get_anng_edge_optimization_parameter()
    assert (p.no_of_queries, p.no_of_results, p.no_of_threads, p.target_no_of_objects, p.no_of_sample_objects,
            p.max_of_no_of_edges, p.log) == (200, 50, 16, 0, 100000, 100, False)
    assert abs(p.target_accuracy - 0.9) < 1e-7
    err = lib.ngt_create_error_object()
    assert lib.ngt_optimize_number_of_edges(b"/nonexistent", p, err) is False        # a tuner: outside the hot path, refused loudly
    assert "not provided by the B200 engine" in lib.ngt_get_error_string(err).decode()
    lib.ngt_destroy_error_object(err)


def test_reference_base_py_binds_libngtgpu_unmodified():
    """python/ngt/base.py of the reference, unmodified, imports against libngtgpu.so: every ngt_* symbol its class body
    touches resolves (the functional run of the same module is in tests/test_gpu_capi.py)."""
    import capi
    from ngt_b200 import _lib
    mod = capi.load_reference_base_py(_lib.SO_PATH)
    if mod is None:
        pytest.skip("no copy of the reference's python/ngt/base.py on this box")
    assert hasattr(mod.Index, "search") and hasattr(mod.Index, "insert_blob") and hasattr(mod, "NativeError")
    with pytest.raises(mod.NativeError) as ei:          # error convention: NULL + message -> NativeError
        mod.Index(b"/nonexistent/idx")
    assert b"Cannot load the property file /nonexistent/idx/prf." in ei.value.args[0]


def test_accuracy_table_equals_the_reference():
    """Index::AccuracyTable::getEpsilon (Index.h:293-360): the C ABI function and the Python mirror against values the
    reference returned (tests/golden/accuracy_table.json), bit for bit, and the reference's messages for bad tables."""
    import ctypes as C
    import json
    from conftest import GOLDEN
    from ngt_b200 import _lib
    from ngt_b200.index import epsilon_from_accuracy_table
    lib = _lib.load()
    z = json.load(open(os.path.join(GOLDEN, "accuracy_table.json")))
    for acc, eps in z["cases"]:
        e = C.c_float(0)
        assert lib.ngtgpu_epsilon_from_accuracy_table(z["table"].encode(), acc, C.byref(e)) == 0
        assert np.float32(e.value) == np.float32(eps), (acc, e.value, eps)
        assert np.float32(epsilon_from_accuracy_table(z["table"], acc)) == np.float32(eps)
    for name, (table, msg) in z["errors"].items():
        e = C.c_float(0)
        assert lib.ngtgpu_epsilon_from_accuracy_table(table.encode(), 0.9, C.byref(e)) != 0
        assert lib.ngtgpu_last_error().decode() == msg, name
        with pytest.raises(_lib.NgtGpuError) as ei:
            epsilon_from_accuracy_table(table, 0.9)
        assert str(ei.value) == msg

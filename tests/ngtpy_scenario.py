"""One script written against the `ngtpy` module (python/src/ngtpy.cpp:500-639; the API of python/README-ngtpy.md and
python/sample/sample.py), run UNCHANGED on two modules of that name: the reference's (oracle/_ref/ngtpy, by
tests/golden/make_golden_ngtpy.py, which stores what it returns in tests/golden/ngtpy_scenario.json) and the repo's
(ngt_b200/ngtpy.*.so, by tests/test_gpu_ngtpy.py on the B200). Only the ngtpy API is used on the index; the ONNG file the
optimizer writes is read back with the repo's `grp` reader on both sides."""
import os
import shutil

import numpy as np


def _pairs(res):
    return [[int(i), float(d)] for (i, d) in res]


def run(ngtpy, golden_dir, tmp, read_graph):
    z = np.load(os.path.join(golden_dir, "sift5k.npz"))
    data, queries = z["data"].astype(np.float32), z["queries"].astype(np.float32)
    idx200 = os.path.join(golden_dir, "idx200")
    out = {}

    # ---- A: an index the reference's `ngt create` built, read-only, both numberings ---------------------------------
    for name, zero in (("one_based", False), ("zero_based", True)):
        ix = ngtpy.Index(idx200, read_only=True, zero_based_numbering=zero, log_disabled=True)
        rec = {"linear": [], "linear_ids": [], "graph": [], "graph_ids": []}
        for q in queries:
            rec["linear"].append(_pairs(ix.linear_search(q, size=5)))
            rec["linear_ids"].append([int(i) for i in ix.linear_search(q, size=5, with_distance=False)])
            rec["graph"].append(_pairs(ix.search(q, size=5, epsilon=0.3)))
            rec["graph_ids"].append([int(i) for i in ix.search(q, size=5, epsilon=0.3, with_distance=False)])
        rec["object"] = [float(v) for v in ix.get_object(7)]
        rec["default_size"] = len(ix.linear_search(queries[0]))                        # 20
        ix.set(num_of_search_objects=3)
        rec["set_size"] = _pairs(ix.linear_search(queries[0]))
        d = rec["linear"][0]
        ix.set(search_radius=(d[1][1] + d[2][1]) / 2)                                  # between the 2nd and the 3rd
        rec["set_radius"] = _pairs(ix.linear_search(queries[0], size=5))
        rec["set_radius_graph"] = _pairs(ix.search(queries[0], size=5, epsilon=0.3))
        rec["wrong_dimension"] = len(ix.search(queries[0][:64], size=5))               # message on stderr, empty result
        rec["wrong_dimension_ids"] = len(ix.linear_search(queries[0][:64], size=5, with_distance=False))
        ix.close()
        out[name] = rec

    # ---- B: create -> batch_insert -> insert -> build_index -> remove -> save -> reopen (float L2) -------------------
    a = os.path.join(tmp, "a")
    ngtpy.create(a, 128, edge_size_for_creation=10, edge_size_for_search=40, distance_type="L2", object_type="Float")
    ix = ngtpy.Index(a, log_disabled=True)
    ix.batch_insert(data[:300].astype(np.float64), num_threads=4)
    rec = {"inserted_id": int(ix.insert(data[300].astype(np.float64)))}
    ix.build_index(4)
    rec["linear"] = [_pairs(ix.linear_search(q, size=5)) for q in queries]
    rec["graph"] = [_pairs(ix.search(q, size=5, epsilon=0.3)) for q in queries]
    rec["self"] = _pairs(ix.linear_search(data[300], size=2))
    ix.remove(12)
    rec["after_remove"] = _pairs(ix.linear_search(data[12], size=3))
    rec["after_remove_graph"] = _pairs(ix.search(data[12], size=3, epsilon=0.3))
    ix.save()
    ix.close()
    ix = ngtpy.Index(a, read_only=True, log_disabled=True)
    rec["reopened"] = [_pairs(ix.linear_search(q, size=5)) for q in queries]
    rec["reopened_object"] = [float(v) for v in ix.get_object(300)]
    ix.close()
    out["float_l2"] = rec

    # ---- C: byte objects (uint8 L2) and Hamming ------------------------------------------------------------------------
    b = os.path.join(tmp, "b")
    ngtpy.create(b, 128, object_type="Byte")
    ix = ngtpy.Index(b, log_disabled=True)
    ix.batch_insert(data[:300].astype(np.float64))
    out["byte_l2"] = {"linear": [_pairs(ix.linear_search(q, size=5)) for q in queries],
                      "object": [float(v) for v in ix.get_object(5)]}
    ix.close()
    h = os.path.join(tmp, "h")
    ngtpy.create(h, 16, distance_type="Hamming", object_type="Byte")
    ix = ngtpy.Index(h, log_disabled=True)
    bits = (data[:300, :16] // 4).astype(np.float64)
    ix.batch_insert(bits)
    out["hamming"] = {"linear": [_pairs(ix.linear_search(q[:16] // 4, size=5)) for q in queries]}
    ix.close()

    # ---- D: a normalising space ------------------------------------------------------------------------------------------
    c = os.path.join(tmp, "c")
    ngtpy.create(c, 128, distance_type="Normalized Cosine")
    ix = ngtpy.Index(c, log_disabled=True)
    ix.batch_insert((data[:300] + 1.0).astype(np.float64))
    out["normalized_cosine"] = {"linear": [_pairs(ix.linear_search(q + 1.0, size=5)) for q in queries],
                                "object": [float(v) for v in ix.get_object(0)]}
    ix.close()

    # ---- E: Optimizer.execute on the reference-built ANNG (reconstruction + shortcut reduction, no timed tuning) ------
    o = ngtpy.Optimizer(log_disabled=True)
    o.set(num_of_outgoings=5, num_of_incomings=20)
    o.set_processing_modes(shortcut_reduction=True, search_parameter_optimization=False,
                           prefetch_parameter_optimization=False, accuracy_table_generation=False)
    onng = os.path.join(tmp, "onng200")
    o.execute(idx200, onng)
    row_ptr, col, dist, _ = read_graph(onng)
    out["onng"] = {"row_ptr": [int(v) for v in row_ptr], "col": [int(v) for v in col], "dist": [float(v) for v in dist]}
    ix = ngtpy.Index(onng, read_only=True, zero_based_numbering=False, log_disabled=True)
    out["onng"]["graph"] = [_pairs(ix.search(q, size=5, epsilon=0.3)) for q in queries]
    ix.close()
    o2 = ngtpy.Optimizer(num_of_outgoings=3, num_of_incomings=8, log_disabled=True)
    o2.set_processing_modes(shortcut_reduction=False, search_parameter_optimization=False,
                            prefetch_parameter_optimization=False, accuracy_table_generation=False)
    onng2 = os.path.join(tmp, "onng200b")
    o2.execute(idx200, onng2)
    row_ptr, col, dist, _ = read_graph(onng2)
    out["onng_no_shortcut"] = {"row_ptr": [int(v) for v in row_ptr], "col": [int(v) for v in col]}

    # ---- F: refine_anng on a writable copy ---------------------------------------------------------------------------------
    r = os.path.join(tmp, "refine")
    shutil.copytree(idx200, r)
    ix = ngtpy.Index(r, zero_based_numbering=False, log_disabled=True)
    ix.refine_anng(epsilon=0.1, num_of_edges=0, batch_size=100)
    out["refined"] = {"graph": [_pairs(ix.search(q, size=5, epsilon=0.3)) for q in queries]}
    ix.close()
    return out


def run_reference_sample(module_dir, golden_dir, tmp, sample_py):
    """python/sample/sample.py of the reference, UNMODIFIED (`sample_py`: the copy oracle/Makefile puts under oracle/_ref/),
    as a separate interpreter with `module_dir` first on the module path so that its `import ngtpy` finds the module under
    test; the two data files it reads (../../data/sift-*.tsv) are written from tests/golden/sift5k.npz. -> CompletedProcess"""
    import subprocess
    import sys
    z = np.load(os.path.join(golden_dir, "sift5k.npz"))
    os.makedirs(os.path.join(tmp, "data"))
    cwd = os.path.join(tmp, "python", "sample")
    os.makedirs(cwd)
    for name, arr in (("sift-dataset-5k.tsv", z["data"]), ("sift-query-3.tsv", z["queries"])):
        with open(os.path.join(tmp, "data", name), "w") as f:
            for row in arr:
                f.write("\t".join(str(int(v)) for v in row) + "\n")
    env = dict(os.environ, PYTHONPATH=module_dir + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run([sys.executable, sample_py], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)

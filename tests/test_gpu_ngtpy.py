"""The `ngtpy` module of the repo (ngt_b200/csrc/ngtpy.cpp: pybind11 over the C ABI of libngtgpu.so) against the
reference's module of the same name: tests/ngtpy_scenario.py -- one script written against the ngtpy API -- was run on
the UNMODIFIED reference module by tests/golden/make_golden_ngtpy.py; here the same script runs on the B200 engine and
must return the same answers (bit for bit where the arithmetic is integer-valued, 1e-6 for the normalising space)."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ngtpy():
    sys.path.insert(0, os.path.join(ROOT, "ngt_b200"))
    try:
        import ngtpy as mod
    finally:
        sys.path.pop(0)
    assert mod.__file__.startswith(os.path.join(ROOT, "ngt_b200")), mod.__file__     # not the reference's build
    return mod


@pytest.fixture(scope="module")
def got(ngtpy, tmp_path_factory):
    import ngtpy_scenario
    from ngt_b200 import index_io
    return ngtpy_scenario.run(ngtpy, GOLDEN, str(tmp_path_factory.mktemp("ngtpy")), index_io.read_graph)


@pytest.fixture(scope="module")
def ref():
    return json.load(open(os.path.join(GOLDEN, "ngtpy_scenario.json")))


def same_bits(a, b):
    return [[i, np.float32(d).view(np.uint32)] for i, d in a] == [[i, np.float32(d).view(np.uint32)] for i, d in b]


def overlap(a, b):
    return len({i for i, _ in a} & {i for i, _ in b}) / max(len(b), 1)


@pytest.mark.parametrize("numbering", ["one_based", "zero_based"])
def test_reference_built_index_through_ngtpy(got, ref, numbering):
    g, r = got[numbering], ref[numbering]
    for qi in range(3):
        assert same_bits(g["linear"][qi], r["linear"][qi])
        assert g["linear_ids"][qi] == r["linear_ids"][qi]
        assert same_bits(g["graph"][qi], r["graph"][qi])          # 200 objects, epsilon 0.3: both find the exact answer
        assert g["graph_ids"][qi] == r["graph_ids"][qi]
    assert g["object"] == r["object"]
    assert g["default_size"] == r["default_size"] == 20           # python/src/ngtpy.cpp:43
    assert same_bits(g["set_size"], r["set_size"])
    assert same_bits(g["set_radius"], r["set_radius"]) and len(g["set_radius"]) == 2
    assert same_bits(g["set_radius_graph"], r["set_radius_graph"])
    assert g["wrong_dimension"] == r["wrong_dimension"] == 0
    assert g["wrong_dimension_ids"] == r["wrong_dimension_ids"] == 0


def test_create_insert_build_remove_save_reopen(got, ref):
    g, r = got["float_l2"], ref["float_l2"]
    assert g["inserted_id"] == r["inserted_id"] == 300
    for qi in range(3):
        assert same_bits(g["linear"][qi], r["linear"][qi])
        assert same_bits(g["reopened"][qi], r["reopened"][qi])
        assert overlap(g["graph"][qi], r["linear"][qi]) >= 0.8    # the graphs differ (seeds of the build), the answers do not
    assert same_bits(g["self"], r["self"])
    assert same_bits(g["after_remove"], r["after_remove"])
    assert overlap(g["after_remove_graph"], r["after_remove"]) >= 0.66
    assert 12 not in [i for i, _ in g["after_remove_graph"]]
    assert g["reopened_object"] == r["reopened_object"]


def test_byte_hamming_and_normalising_spaces(got, ref):
    for qi in range(3):
        assert same_bits(got["byte_l2"]["linear"][qi], ref["byte_l2"]["linear"][qi])
        assert same_bits(got["hamming"]["linear"][qi], ref["hamming"]["linear"][qi])
        a, b = got["normalized_cosine"]["linear"][qi], ref["normalized_cosine"]["linear"][qi]
        assert [i for i, _ in a] == [i for i, _ in b]
        assert np.allclose([d for _, d in a], [d for _, d in b], rtol=0, atol=1e-6)
    assert got["byte_l2"]["object"] == ref["byte_l2"]["object"]
    assert np.allclose(got["normalized_cosine"]["object"], ref["normalized_cosine"]["object"], rtol=0, atol=1e-6)


def test_optimizer_execute_writes_the_reference_onng(got, ref):
    g, r = got["onng"], ref["onng"]
    assert g["row_ptr"] == r["row_ptr"] and g["col"] == r["col"]
    assert (np.array(g["dist"], np.float32).view(np.uint32) == np.array(r["dist"], np.float32).view(np.uint32)).all()
    for qi in range(3):
        assert same_bits(g["graph"][qi], r["graph"][qi])
    assert got["onng_no_shortcut"] == ref["onng_no_shortcut"]
    for qi in range(3):
        assert overlap(got["refined"]["graph"][qi], ref["refined"]["graph"][qi]) >= 0.8


def test_batch_entry_points_and_counters(ngtpy, sift5k):
    ix = ngtpy.Index(os.path.join(GOLDEN, "idx200"), read_only=True, zero_based_numbering=False, log_disabled=True)
    qs = sift5k["queries"].astype(np.float32)
    assert ix.get_num_of_distance_computations() == 0
    one = [ix.search(q, size=5, epsilon=0.3) for q in qs]
    n1 = ix.get_num_of_distance_computations()
    assert n1 >= 3 * 5                                            # Graph.cpp:592: one per evaluated object
    ids, dists = ix.batch_search(qs, size=5, epsilon=0.3)
    assert ids.dtype == np.int64 and dists.dtype == np.float32 and ids.shape == (3, 5)
    assert ids.tolist() == [[i for i, _ in r] for r in one]
    assert (dists == np.array([[d for _, d in r] for r in one], np.float32)).all()
    lids, ldists = ix.batch_linear_search(qs, size=250)           # more than the index holds: -1 past the end
    assert (lids[:, :200] > 0).all() and (lids[:, 200:] == -1).all()
    assert lids[:, :5].tolist() == [[i for i, _ in ix.linear_search(q, size=5)] for q in qs]
    with pytest.raises(RuntimeError):
        ix.batch_search(qs[:, :64], size=5)
    with pytest.raises(RuntimeError, match="accuracy table"):
        ix.batch_search(qs, size=5, expected_accuracy=0.9)        # idx200 carries no table (Index.h:331-335)
    ix.close()
    with pytest.raises(RuntimeError, match="closed"):
        ix.search(qs[0])


def test_the_reference_sample_script_runs_unmodified(tmp_path):
    """python/sample/sample.py of the reference (create, batch_insert of rows given as strings, save, reopen, search,
    get_object, 5 000 single inserts, build_index, remove, save) executed as it is, `import ngtpy` resolving to the repo's
    module: the result tables it prints are the ones the reference's module printed (tests/golden/ngtpy_sample.txt; top-5/6
    of a 5k / 10k index at the default epsilon 0.1 -- both find the exact neighbours); the distance-computation counters
    differ (other seeds, other graph) and are only required to be there and to grow."""
    import ngtpy_scenario
    sample = os.path.join(ROOT, "oracle", "_ref", "python", "sample", "sample.py")
    if not os.path.exists(sample):
        pytest.skip("no copy of the reference's python/sample/sample.py on this box")
    done = ngtpy_scenario.run_reference_sample(os.path.join(ROOT, "ngt_b200"), GOLDEN, str(tmp_path), sample)
    assert done.returncode == 0, done.stderr[-3000:]
    want = open(os.path.join(GOLDEN, "ngtpy_sample.txt")).read().splitlines()
    have = done.stdout.splitlines()
    key = "# of distance computations="
    assert [l for l in have if not l.startswith(key)] == [l for l in want if not l.startswith(key)]
    counts = [int(l[len(key):]) for l in have if l.startswith(key)]
    assert len(counts) == 4 and counts[0] > 0 and counts[1] > counts[0]           # accumulated until the next insertion
    assert counts[2] > 0 and counts[3] > counts[2]

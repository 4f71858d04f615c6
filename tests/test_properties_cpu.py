"""Size-independent properties of the path, checked on the CPU (no device): they tie the oracle's graph search to its
own exhaustive search, and the host twins of the multi-GPU key format to plain numpy, on randomised inputs.

* exhaustive search (ObjectSpaceRepository::linearSearch, lib/NGT/ObjectSpaceRepository.h:466-502) == numpy's sort of
  the exact integer distances (uint8 L2, Hamming);
* graph search (NeighborhoodGraph::search, lib/NGT/Graph.cpp:398-495) on a COMPLETE graph visits every object whatever
  epsilon is, so it equals the exhaustive search; on any graph its results are a subset of the objects reachable from
  the seeds, ascending, without repeats, and no object's distance is computed twice;
* 64-bit keys (ordered distance bits << 32 | id): unsigned order == (distance, id) order of ObjectDistance::operator<
  (lib/NGT/Common.h:1946-1952); pack -> unpack is the identity; the k-way merge of per-shard lists == the k smallest of
  the union."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pyoracle as po


def _complete_graph(n):
    row_ptr = np.zeros(n + 2, np.uint64)
    row_ptr[2:] = np.cumsum(np.full(n, n - 1, np.uint64))
    col = np.concatenate([np.delete(np.arange(1, n + 1, dtype=np.uint32), i) for i in range(n)])
    return row_ptr, col


def _knn_graph(port, dtype, otype, pobj, k):
    ids, _, counts = port.linear_search(dtype, otype, pobj, pobj[1:], k + 1)
    n = pobj.shape[0] - 1
    lists = [[int(t) for t in ids[i, :counts[i]] if int(t) != i + 1][:k] for i in range(n)]
    row_ptr = np.zeros(n + 2, np.uint64)
    row_ptr[2:] = np.cumsum(np.array([len(x) for x in lists], np.uint64))
    return row_ptr, np.array([t for x in lists for t in x], np.uint32), lists


@settings(max_examples=12, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from(["u8l2", "ham"]), st.integers(1, 40))
def test_linear_search_is_the_sorted_exact_distance(port, seed, kind, k):
    rng = np.random.default_rng(seed)
    n, nq, dim = 300, 8, 32
    objs = rng.integers(0, 256, (n, dim), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, dim), dtype=np.uint8)
    dtype = po.L2 if kind == "u8l2" else po.HAMMING
    pobj, pq = po.pad_objects(objs, po.UINT8), po.pad_queries(q.astype(np.float32), po.UINT8)
    ids, dists, counts = port.linear_search(dtype, po.UINT8, pobj, pq, k)
    for i in range(nq):
        if kind == "u8l2":
            exact = ((objs.astype(np.int64) - q[i].astype(np.int64)) ** 2).sum(1)
            want_d = np.sqrt(exact.astype(np.float64)).astype(np.float32)
        else:
            exact = np.unpackbits(objs ^ q[i], axis=1).sum(1).astype(np.int64)
            want_d = exact.astype(np.float32)
        order = np.lexsort((np.arange(1, n + 1), exact))[:k]     # ascending (distance, id)
        assert counts[i] == min(k, n)
        assert (ids[i, :counts[i]] == order + 1).all()
        assert (dists[i, :counts[i]].view(np.uint32) == want_d[order].view(np.uint32)).all()


@settings(max_examples=8, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([0.0, 0.1, 0.5]), st.integers(1, 20))
def test_graph_search_on_a_complete_graph_is_exhaustive(port, seed, eps, k):
    rng = np.random.default_rng(seed)
    n, nq, dim = 60, 6, 16
    objs = rng.integers(0, 256, (n, dim), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, dim)).astype(np.float32)
    pobj, pq = po.pad_objects(objs, po.UINT8), po.pad_queries(q, po.UINT8)
    row_ptr, col = _complete_graph(n)
    seeds = np.stack([rng.choice(n, 3, replace=False) + 1 for _ in range(nq)]).astype(np.uint32)
    ids, dists, counts, stats = port.graph_search(po.L2, po.UINT8, pobj, row_ptr, col, pq, seeds, k, eps)
    rids, rdists, rcounts = port.linear_search(po.L2, po.UINT8, pobj, pq, k)
    assert (counts == rcounts).all() and (ids == rids).all()
    assert (dists.view(np.uint32) == rdists.view(np.uint32)).all()
    assert (stats[:, 0] == n).all()          # every object's distance is computed exactly once


@settings(max_examples=8, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1))
def test_graph_search_results_are_reachable_sorted_and_unique(port, seed):
    rng = np.random.default_rng(seed)
    n, nq, dim, k = 400, 10, 24, 10
    objs = rng.integers(0, 256, (n, dim), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, dim)).astype(np.float32)
    pobj, pq = po.pad_objects(objs, po.UINT8), po.pad_queries(q, po.UINT8)
    row_ptr, col, lists = _knn_graph(port, po.L2, po.UINT8, pobj, 6)
    seeds = np.stack([rng.choice(n, 4, replace=False) + 1 for _ in range(nq)]).astype(np.uint32)
    for eps in (0.0, 0.1, 0.3, 1.0):
        ids, dists, counts, stats = port.graph_search(po.L2, po.UINT8, pobj, row_ptr, col, pq, seeds, k, eps)
        for i in range(nq):
            c = int(counts[i])
            got = ids[i, :c]
            assert len(set(got.tolist())) == c
            keys = [(float(dists[i, j]), int(ids[i, j])) for j in range(c)]
            assert keys == sorted(keys)
            reach, todo = set(int(s) for s in seeds[i]), [int(s) for s in seeds[i]]
            while todo:
                for t in lists[todo.pop() - 1]:
                    if t not in reach:
                        reach.add(t)
                        todo.append(t)
            assert set(got.tolist()) <= reach
            assert stats[i, 0] <= len(reach)     # distance computations: at most one per reachable object


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 6), st.integers(1, 12))
def test_key_format_round_trip_and_merge_is_the_union(seed, world, k):
    from ngt_b200 import sharded
    rng = np.random.default_rng(seed)
    nq, n_local = 5, 50
    per, union = [], [[] for _ in range(nq)]
    for r in range(world):
        counts = rng.integers(0, k + 1, nq).astype(np.uint32)
        ids = np.zeros((nq, k), np.uint32)
        d = np.zeros((nq, k), np.float32)
        for qi in range(nq):
            c = int(counts[qi])
            # few distinct distance values: ties across shards are the rule, and are broken by the global id
            rows = sorted(zip(rng.integers(0, 4, c).astype(np.float32).tolist(), rng.choice(n_local, c, replace=False) + 1))
            for j, (dd, t) in enumerate(rows):
                ids[qi, j], d[qi, j] = t, dd
                union[qi].append((dd, int(t) + r * n_local))
        keys = sharded.pack_keys_host(ids, d, counts, r * n_local)
        ui, ud, valid = sharded.unpack_keys_host(keys)
        assert (valid.sum(1) == counts).all()
        for qi in range(nq):
            c = int(counts[qi])
            assert (ui[qi, :c] == ids[qi, :c] + r * n_local).all() and (ud[qi, :c] == d[qi, :c]).all()
            assert (keys[qi, 1:c] > keys[qi, :max(c - 1, 0)]).all()   # unsigned key order == (distance, id) order
        per.append(keys)
    mi, md, mc = sharded.merge_keys_host(np.stack(per), k)
    for qi in range(nq):
        want = sorted(union[qi])[:k]
        assert int(mc[qi]) == len(want)
        assert [(float(md[qi, j]), int(mi[qi, j])) for j in range(len(want))] == want


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.integers(0, 40), st.integers(0, 12), st.booleans())
def test_rank_merge_of_the_construction_loop_equals_the_full_sort(seed, n_old, n_new, allow_repeats):
    """ngtgpu_index_insert_batch merges a batch's new edges into a node's sorted list by rank (merge_lists_kernel) and
    takes the full sort when the kernel raises its flag. Restated in oracle/pyoracle.py: whenever the flag stays down the
    rank merge IS the full sort's list (which drops an entry whose target equals the entry before it), and the flag is
    up whenever the full sort would have dropped anything."""
    rng = np.random.default_rng(seed)
    targets = rng.choice(60, n_old + n_new, replace=allow_repeats) + 1      # few distances: ties are the rule
    keys = [(int(rng.integers(0, 5)) << 32) | int(t) for t in targets]
    if allow_repeats and n_old and n_new and rng.random() < 0.5:
        keys[n_old] = keys[int(rng.integers(0, n_old))]                   # the same edge again (an id inserted twice)
    old = sorted(set(keys[:n_old]))
    if not allow_repeats or rng.random() < 0.5:
        seen, uniq = set(), []
        for k in old:                                                     # a list as the engine keeps it: one entry per target
            if k & 0xFFFFFFFF not in seen:
                uniq.append(k)
                seen.add(k & 0xFFFFFFFF)
        old = uniq
    new = sorted(keys[n_old:])
    want = po.full_sort_lists(old, new)
    got, redo = po.merge_lists_by_rank(old, new)
    if len(want) != len(old) + len(new):
        assert redo                      # something was dropped: the batch must take the full sort
    if not redo:
        assert got == want
        assert all(b > a for a, b in zip(got, got[1:]))


KEY_NONE = 0xFFFFFFFFFFFFFFFF


def _res_insert(res, kk, k, kl):
    """search_fast.cuh res_insert<KL>, lane by lane: res[lane][slot], position p on lane p // KL, slot p % KL."""
    pos = sum(1 for lane in range(32) for m in range(kl) if res[lane][m] < kk)            # KL ballots + popc
    up = [res[lane - 1][kl - 1] if lane > 0 else res[0][kl - 1] for lane in range(32)]    # shfl_up of the last slot
    for lane in range(32):
        for m in range(kl - 1, -1, -1):
            idx = lane * kl + m
            prev = res[lane][m - 1] if m > 0 else up[lane]
            if idx == pos:
                res[lane][m] = kk
            elif idx > pos:
                res[lane][m] = prev
            if idx >= k:
                res[lane][m] = KEY_NONE


@settings(max_examples=60, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([1, 4]), st.integers(1, 128), st.integers(0, 300))
def test_result_list_of_the_lean_kernel_is_the_k_smallest(seed, kl, k, n_keys):
    """The control warp's result list (search_fast.cuh: KL keys per lane, k <= 32 * KL) after any sequence of distinct
    keys holds the k smallest in ascending positions, and res_kth reads position k - 1 -- ResultSet + the pop of the
    extras in NeighborhoodGraph::search (lib/NGT/Graph.cpp:467-479), restated lane by lane."""
    k = min(k, 32 * kl)
    rng = np.random.default_rng(seed)
    keys = [(int(d) << 32) | int(t) for d, t in zip(rng.integers(0, 50, n_keys), rng.permutation(n_keys) + 1)]
    res = [[KEY_NONE] * kl for _ in range(32)]
    seen = []
    for kk in keys:
        _res_insert(res, kk, k, kl)
        seen.append(kk)
        want = sorted(seen)[:k]
        flat = [res[p // kl][p % kl] for p in range(32 * kl)]
        assert flat[:len(want)] == want
        assert all(v == KEY_NONE for v in flat[len(want):])
        kth = res[(k - 1) // kl][(k - 1) % kl]                                          # res_kth<KL>(res, k - 1)
        assert kth == (want[k - 1] if len(want) == k else KEY_NONE)

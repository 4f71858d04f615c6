#!/usr/bin/env python
"""Generates tests/golden/*.npz + readme_kat.json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py
The outputs are committed; tests never read /root/reference.

What each fixture pins
  readme_kat.json : the reference's only known-answer listing, bin/ngt/README.md:254-323
                    (`ngt create -d 128 -o c` + `ngt search -n 20`), parsed from the README text.
  sift5k.npz      : data/sift-dataset-5k.tsv (5000x128, uint8-valued) + data/sift-query-3.tsv, the ANNG
                    the reference builds on it (defaults E=10,S=40), the seeds its DVP-tree returns for
                    the 3 queries, and reference outputs of linearSearch / GraphIndex::search for an
                    (epsilon, edge size) grid with explicit seeds, for both `-o c` (uint8) and `-o f`.
  synth.npz       : small synthetic sets of the BASELINE shapes for every distance type on the path
                    (uint8 L2, Hamming, float L2 d=100/960, Cosine, Angle, Normalized*), reference
                    linearSearch and graph search outputs (ANNG and ONNG graphs) with explicit seeds.
"""
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from ngt_b200 import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
EPS_GRID = (0.0, 0.1, 0.3)
EDGE_GRID = (-1, 0, 5, -2)


def readme_kat():
    lines = open(os.path.join(REF, "bin/ngt/README.md")).read().split("\n")[250:330]
    queries, cur = [], None
    for ln in lines:
        ln = ln.strip()
        if ln.startswith("Query No."):
            cur = []
            queries.append(cur)
            continue
        m = re.match(r"^(\d+)\s+(\d+)\s+([0-9.]+)$", ln)
        if m and cur is not None:
            assert int(m.group(1)) == len(cur) + 1
            cur.append([int(m.group(2)), m.group(3)])
    assert len(queries) == 3 and all(len(q) == 20 for q in queries), [len(q) for q in queries]
    return queries


def fmt6(x):
    """The CLI prints distances with the default ostream precision (6 significant digits)."""
    return "%g" % float(x)


def search_grid(R, h, queries, seeds, k, tag, out):
    for eps in EPS_GRID:
        for es in EDGE_GRID:
            ids, dists, counts, stats, _ = R.search(h, queries, k, epsilon=eps, edge_size=es, seeds=seeds)
            key = "%s_e%02d_s%d" % (tag, int(round(eps * 100)), es)
            out[key + "_ids"] = ids
            out[key + "_dists"] = dists
            out[key + "_counts"] = counts
            out[key + "_stats"] = stats


def make_sift5k(R, tmp):
    data = np.ascontiguousarray(np.loadtxt(os.path.join(REF, "data/sift-dataset-5k.tsv"), dtype=np.float32)[:, :128])
    qs = np.loadtxt(os.path.join(REF, "data/sift-query-3.tsv"), dtype=np.float32)
    kat = readme_kat()
    out = {"data": data.astype(np.uint8), "queries": qs.astype(np.uint8)}
    rng = np.random.default_rng(5)
    seeds = np.stack([rng.choice(5000, 10, replace=False) + 1 for _ in range(3)]).astype(np.uint32)
    out["seeds"] = seeds
    for ot, tag in (("c", "u8"), ("f", "f32")):
        path = os.path.join(tmp, "sift5k-" + tag)
        R.build_index(path, data, objtype=ot, disttype=po.L2, edge_creation=10, edge_search=40, threads=4)
        h = R.open(path, readonly=False)
        inf = R.info(h)
        rp, col, dist = R.graph(h)
        if tag == "u8":
            out["row_ptr"], out["col"] = rp, col
            out["prop"] = np.array([inf["edge_size_for_search"], inf["dyn_base"], inf["dyn_rate"]], np.int64)
            # tree-seeded search must reproduce the README listing digit for digit
            ids, dists, counts, _, _ = R.search(h, qs, 20, epsilon=0.1, edge_size=-1, seeds=None)
            for q in range(3):
                got = [[int(ids[q, i]), fmt6(dists[q, i])] for i in range(20)]
                assert got == kat[q], (q, got, kat[q])
            ts, tn = R.tree_seeds(h, qs, 20)
            assert (tn == 10).all()
            out["tree_seeds"] = ts[:, :10].copy()
            out["tree_ids"], out["tree_dists"] = ids, dists
        else:
            assert (out["row_ptr"] == rp).all() and (out["col"] == col).all(), "u8 and f32 ANNGs differ"
        li, ld, lc, _ = R.linear_search(h, qs, 20)
        out[tag + "_lin_ids"], out[tag + "_lin_dists"] = li, ld
        search_grid(R, h, qs, seeds, 20, tag, out)
        R.close(h)
    json.dump(kat, open(os.path.join(OUT, "readme_kat.json"), "w"))
    np.savez_compressed(os.path.join(OUT, "sift5k.npz"), **out)
    print("sift5k.npz", os.path.getsize(os.path.join(OUT, "sift5k.npz")))


SYNTH_CASES = [
    # tag, shape, n, object type, distance, E
    ("u8l2", "sift", 2000, "c", po.L2, 10),
    ("ham", "sift", 2000, "c", po.HAMMING, 10),
    ("f32l2", "sift", 2000, "f", po.L2, 10),
    ("glove_l2", "glove", 1500, "f", po.L2, 10),
    ("glove_cos", "glove", 1500, "f", po.COSINE, 10),
    ("glove_ang", "glove", 1500, "f", po.ANGLE, 10),
    ("glove_ncos", "glove", 1500, "f", po.NORMALIZED_COSINE, 10),
    ("glove_nang", "glove", 1500, "f", po.NORMALIZED_ANGLE, 10),
    ("glove_nl2", "glove", 1500, "f", po.NORMALIZED_L2, 10),
    ("gist_l2", "gist", 400, "f", po.L2, 10),
]


def make_synth(R, tmp):
    out = {}
    raw = {}
    stored_by_tag = {}
    for tag, shape, n, ot, dt, E in SYNTH_CASES:
        if (shape, n) not in raw:
            raw[(shape, n)] = (synth.make(shape, n, 1), synth.make(shape, 16, 2))
        base, qs = raw[(shape, n)]
        if dt == po.HAMMING:
            base = synth.hamming_from(base, 64.0).astype(np.float32)   # 16 bytes/object, fed as numbers
            qs = synth.hamming_from(qs, 64.0).astype(np.float32)
        path = os.path.join(tmp, "synth-" + tag)
        R.build_index(path, base, objtype=ot, disttype=dt, edge_creation=E, edge_search=40, threads=4)
        h = R.open(path, readonly=False)
        inf = R.info(h)
        stored = R.objects(h)            # exactly what the reference holds (normalised when it normalises)
        rp, col, dist = R.graph(h)
        alias = [t for t, a in stored_by_tag.items() if a.shape == stored.shape and a.dtype == stored.dtype
                 and (a.view(np.uint8) == stored.view(np.uint8)).all()]
        if alias:                      # same stored bytes as an earlier case: keep one copy
            out[tag + "_objects_alias"] = np.array(alias[0])
        else:
            out[tag + "_objects"] = stored
            stored_by_tag[tag] = stored
        out[tag + "_queries"] = qs.astype(np.uint8) if ot == "c" else qs
        out[tag + "_row_ptr"], out[tag + "_col"] = rp.astype(np.uint32), col
        out[tag + "_meta"] = np.array([inf["object_type"], inf["distance_type"], inf["dim"],
                                       inf["edge_size_for_search"], inf["dyn_base"], inf["dyn_rate"]], np.int64)
        rng = np.random.default_rng(11)
        seeds = np.stack([rng.choice(n, 10, replace=False) + 1 for _ in range(qs.shape[0])]).astype(np.uint32)
        out[tag + "_seeds"] = seeds
        li, ld, lc, _ = R.linear_search(h, qs, 10)
        out[tag + "_lin_ids"], out[tag + "_lin_dists"] = li, ld
        # radius-limited linear search (ObjectSpaceRepository.h:492): radius = 5th distance of query 0
        rad = float(ld[0, 4])
        li, ld2, lc, _ = R.linear_search(h, qs, 10, radius=rad)
        out[tag + "_linr_ids"], out[tag + "_linr_dists"], out[tag + "_linr_counts"] = li, ld2, lc
        out[tag + "_linr_radius"] = np.array([rad], np.float32)
        search_grid(R, h, qs, seeds, 10, tag, out)
        R.close(h)
    # one ONNG (the metric's graph type): ANNG E=20 -> reconstruct o=5,i=20 with shortcut reduction
    base, qs = raw[("sift", 2000)]
    anng = os.path.join(tmp, "onng-anng")
    onng = os.path.join(tmp, "onng-onng")
    R.build_index(anng, base, objtype="f", disttype=po.L2, edge_creation=20, edge_search=0, threads=4)
    R.build_onng(anng, onng, outgoing=5, incoming=20, shortcut=True)
    h = R.open(onng, readonly=False)
    inf = R.info(h)
    rp, col, dist = R.graph(h)
    out["onng_row_ptr"], out["onng_col"] = rp.astype(np.uint32), col
    out["onng_meta"] = np.array([inf["object_type"], inf["distance_type"], inf["dim"],
                                 inf["edge_size_for_search"], inf["dyn_base"], inf["dyn_rate"]], np.int64)
    out["onng_seeds"] = out["f32l2_seeds"]
    search_grid(R, h, qs, out["onng_seeds"], 10, "onng", out)
    R.close(h)
    # the same ONNG through the read-only path (searchReadOnlyGraph, Graph.cpp:398-495)
    h = R.open(onng, readonly=True)
    ids, dists, counts, _, _ = R.search(h, qs, 10, epsilon=0.1, edge_size=-1, seeds=out["onng_seeds"])
    assert (ids == out["onng_e10_s-1_ids"]).all() and (dists == out["onng_e10_s-1_dists"]).all()
    R.close(h)
    np.savez_compressed(os.path.join(OUT, "synth.npz"), **out)
    print("synth.npz", os.path.getsize(os.path.join(OUT, "synth.npz")))


if __name__ == "__main__":
    po.build(ref=True)
    R = po.Ref()
    tmp = tempfile.mkdtemp(prefix="ngt-golden-")
    try:
        make_sift5k(R, tmp)
        make_synth(R, tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

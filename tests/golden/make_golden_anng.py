#!/usr/bin/env python
"""Golden vectors for the construction path, produced by the UNMODIFIED
reference (oracle/_ref):

  anng_build.npz   NGT::Index::createIndex(threads) -- the batched loop of lib/NGT/Index.cpp:721-792
                   (searchMultipleQueryForCreation on the frozen graph, insertMultipleSearchResults :670-719,
                   insertANNGNode / addEdge, Graph.h:611-626,845-886) -- on a graph-only index whose SeedType
                   property is FixedNodes (every search starts from ids 1..seedSize, Index.h:1122-1127; the
                   reference's own switch, which makes the build independent of rand() and thread scheduling),
                   followed by append + createIndex of more objects into the finished graph.
  refine_anng.npz  GraphReconstructor::refineANNG (GraphReconstructor.h:814-924) on those graphs.

Objects are ngt_b200.synth.make("sift", n, seed) (integer-valued, so distances are exact in every summation
order); only the seeds/parameters and the reference's graphs are stored. Run in the build container:
    python tests/golden/make_golden_anng.py
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from ngt_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# tag: (objtype, n, n_first, data seed, edgeSizeForCreation, edgeSizeForSearch, seedSize, batchSizeForCreation)
BUILD_CASES = {
    "f_b200": ("f", 1300, 1000, 21, 8, 40, 10, 200),     # the reference's default batch size and -S
    "f_b64_all": ("f", 1300, 1000, 21, 8, 0, 10, 64),    # -S 0: all edges explored
    "f_b1000_s5": ("f", 1300, 1000, 21, 8, 5, 10, 1000), # -S 5: short result lists -> the retry of Index.h:826-836
    "u8_b200": ("c", 1200, 1200, 22, 10, 40, 10, 200),   # uint8 objects, one pass
}
# tag: (build case, epsilon, noOfEdges, exploreEdgeSize or None, batchSize)
REFINE_CASES = {
    "r0_all": ("f_b200", 0.1, 0, None, 10000),
    "r0_b400": ("f_b200", 0.1, 0, None, 400),
    "r12_b500": ("f_b200", 0.1, 12, None, 500),
    "rm6_all": ("f_b200", 0.1, -6, None, 10000),
    "u8_r0_b500": ("u8_b200", 0.1, 0, None, 500),
}


def rows_of(case):
    objtype, n, n_first, seed = case[:4]
    return synth.make("sift", n, seed)


if __name__ == "__main__":
    po.build(ref=True)
    R = po.Ref()
    tmp = tempfile.mkdtemp(prefix="ngt-golden-anng-")
    try:
        built, out = {}, {}
        for tag, case in BUILD_CASES.items():
            objtype, n, n_first, seed, e, es, ss, bs = case
            path = os.path.join(tmp, tag)
            R.build_anng_fixed_seeds(path, rows_of(case), n_first, objtype=objtype, disttype=po.L2, edge_creation=e,
                                     edge_search=es, seed_size=ss, batch_size=bs, threads=4)
            h = R.open(path, readonly=False)
            rp, col, dist = R.graph(h)
            R.close(h)
            built[tag] = path
            out[tag + "_meta"] = np.array([ord(objtype), n, n_first, seed, e, es, ss, bs], np.int64)
            out[tag + "_row_ptr"], out[tag + "_col"], out[tag + "_dist"] = rp.astype(np.uint32), col, dist
        np.savez_compressed(os.path.join(OUT, "anng_build.npz"), **out)
        print("anng_build.npz", os.path.getsize(os.path.join(OUT, "anng_build.npz")))
        out = {}
        for tag, (src, eps, noe, explore, bs) in REFINE_CASES.items():
            path = os.path.join(tmp, tag)
            shutil.copytree(built[src], path)
            h = R.open(path, readonly=False)
            R.refine_anng(h, eps, 0.0, noe, -2 ** 31 if explore is None else explore, bs)
            rp, col, dist = R.graph(h)
            R.close(h)
            out[tag + "_meta"] = np.array([noe, -2 ** 31 if explore is None else explore, bs], np.int64)
            out[tag + "_eps"] = np.array([eps], np.float32)
            out[tag + "_src"] = np.array([src])
            out[tag + "_row_ptr"], out[tag + "_col"], out[tag + "_dist"] = rp.astype(np.uint32), col, dist
        np.savez_compressed(os.path.join(OUT, "refine_anng.npz"), **out)
        print("refine_anng.npz", os.path.getsize(os.path.join(OUT, "refine_anng.npz")))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

#!/usr/bin/env python
"""Golden vectors for NGT::Index::remove -> NeighborhoodGraph::removeEdgesReliably (lib/NGT/Graph.cpp:641-864), produced by
the UNMODIFIED reference (oracle/_ref): the ANNG of anng_build.npz case f_b200 (built by the reference), then a sequence of
removals; the graph after all of them is stored.   python tests/golden/make_golden_remove.py"""
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
REMOVED = [5, 700, 1299, 6, 44, 1000, 7, 431, 2]

if __name__ == "__main__":
    po.build(ref=True)
    R = po.Ref()
    zb = np.load(os.path.join(OUT, "anng_build.npz"))
    objtype, n, n_first, seed, e, es, ss, bs = [int(v) for v in zb["f_b200_meta"]]
    tmp = tempfile.mkdtemp(prefix="ngt-golden-remove-")
    try:
        path = os.path.join(tmp, "a")
        R.build_anng_fixed_seeds(path, synth.make("sift", n, seed), n_first, objtype="f", disttype=po.L2, edge_creation=e, edge_search=es,
                                 seed_size=ss, batch_size=bs, threads=4)
        h = R.open(path, readonly=False)
        rp0, col0, _ = R.graph(h)
        assert (col0 == zb["f_b200_col"]).all()          # the same graph the build fixture holds
        for rid in REMOVED:
            R.remove(h, rid)
        rp, col, dist = R.graph(h)
        R.close(h)
        np.savez_compressed(os.path.join(OUT, "remove.npz"), removed=np.array(REMOVED, np.uint32), row_ptr=rp.astype(np.uint32), col=col,
                            dist=dist)
        print("remove.npz", os.path.getsize(os.path.join(OUT, "remove.npz")))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

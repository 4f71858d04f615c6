#!/usr/bin/env python
"""Golden vectors for GraphReconstructor::reconstructGraph (lib/NGT/GraphReconstructor.h:425-561), produced by the
UNMODIFIED reference (oracle/_ref): an ANNG built by the reference and the graph GraphOptimizer::execute writes
from it with path adjustment off (`reconstruct-graph -o 5 -i 20 -s f`). Run in the build container:
    python tests/golden/make_golden_reconstruct.py
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

if __name__ == "__main__":
    po.build(ref=True)
    R = po.Ref()
    tmp = tempfile.mkdtemp(prefix="ngt-golden-rec-")
    try:
        base = synth.make("sift", 1500, 1)
        out = {}
        anng = os.path.join(tmp, "anng")
        R.build_index(anng, base, objtype="f", disttype=po.L2, edge_creation=20, edge_search=0, threads=4)
        h = R.open(anng, readonly=False)
        rp, col, dist = R.graph(h)
        R.close(h)
        out["anng_row_ptr"], out["anng_col"], out["anng_dist"] = rp.astype(np.uint32), col, dist
        for o, i in ((5, 20), (10, 40), (0, 15)):
            onng = os.path.join(tmp, "onng_%d_%d" % (o, i))
            R.build_onng(anng, onng, outgoing=o, incoming=i, shortcut=False)
            h = R.open(onng, readonly=False)
            rp, col, dist = R.graph(h)
            R.close(h)
            key = "o%d_i%d" % (o, i)
            out[key + "_row_ptr"], out[key + "_col"], out[key + "_dist"] = rp.astype(np.uint32), col, dist
        np.savez_compressed(os.path.join(OUT, "reconstruct.npz"), **out)
        print("reconstruct.npz", os.path.getsize(os.path.join(OUT, "reconstruct.npz")))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

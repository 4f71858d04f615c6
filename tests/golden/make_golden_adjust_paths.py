#!/usr/bin/env python
"""Golden vectors for GraphReconstructor::adjustPathsEffectively (lib/NGT/GraphReconstructor.h:197-386), produced by the
UNMODIFIED reference (oracle/_ref): the ONNGs GraphOptimizer::execute writes from a reference-built ANNG with
shortcut reduction ON (`reconstruct-graph -o O -i I -s t`), next to the same graphs with it off (the inputs of the
step; equal to tests/golden/reconstruct.npz). Run in the build container:
    python tests/golden/make_golden_adjust_paths.py
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
    tmp = tempfile.mkdtemp(prefix="ngt-golden-adj-")
    try:
        out = {}
        for tag, shape, n, objtype, disttype, ec in (("sift", "sift", 1500, "f", po.L2, 20),
                                                      ("glove", "glove", 1200, "f", po.COSINE, 15)):
            base = synth.make(shape, n, 1)
            anng = os.path.join(tmp, "anng_" + tag)
            R.build_index(anng, base, objtype=objtype, disttype=disttype, edge_creation=ec, edge_search=0, threads=4)
            for o, i in ((5, 20), (10, 40)):
                for sc in (False, True):
                    onng = os.path.join(tmp, "onng_%s_%d_%d_%d" % (tag, o, i, sc))
                    R.build_onng(anng, onng, outgoing=o, incoming=i, shortcut=sc)
                    h = R.open(onng, readonly=False)
                    rp, col, dist = R.graph(h)
                    R.close(h)
                    key = "%s_o%d_i%d_%s" % (tag, o, i, "adj" if sc else "in")
                    out[key + "_row_ptr"], out[key + "_col"], out[key + "_dist"] = rp.astype(np.uint32), col, dist
        np.savez_compressed(os.path.join(OUT, "adjust_paths.npz"), **out)
        print("adjust_paths.npz", os.path.getsize(os.path.join(OUT, "adjust_paths.npz")))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

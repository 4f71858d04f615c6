#!/usr/bin/env python
"""The accuracy table the UNMODIFIED reference generates (GraphOptimizer::execute with only the accuracy-table step on,
GraphOptimizer.h:355-368 -> Optimizer::generateAccuracyTable, Optimizer.h:1494-1573) for the ANNG of anng_build.npz case
f_b64_all (edgeSizeForSearch 0, built by the reference), after reconstructGraph -o 5 -i 20 + path adjustment.
Stored beside the table of tests/golden/accuracy_table.json.   python tests/golden/make_golden_accuracy_generated.py"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from ngt_b200 import index_io, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    po.build(ref=True)
    R = po.Ref()
    zb = np.load(os.path.join(OUT, "anng_build.npz"))
    objtype, n, n_first, seed, e, es, ss, bs = [int(v) for v in zb["f_b64_all_meta"]]
    tmp = tempfile.mkdtemp(prefix="ngt-golden-acc-")
    try:
        anng, onng = os.path.join(tmp, "a"), os.path.join(tmp, "o")
        R.build_anng_fixed_seeds(anng, synth.make("sift", n, seed), n_first, objtype="f", disttype=po.L2, edge_creation=e, edge_search=es,
                                 seed_size=ss, batch_size=bs, threads=4)
        R.build_onng_with_accuracy_table(anng, onng, 5, 20, True, 100, 20)
        prf = index_io.read_prf(onng)
        p = os.path.join(OUT, "accuracy_table.json")
        z = json.load(open(p))
        z["generated"] = {"case": "f_b64_all", "outgoing": 5, "incoming": 20, "queries": 100, "results": 20, "table": prf["AccuracyTable"]}
        json.dump(z, open(p, "w"), indent=1)
        print(prf["AccuracyTable"])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

#!/usr/bin/env python
"""Golden values for Index::AccuracyTable::getEpsilon (lib/NGT/Index.h:293-360) from the UNMODIFIED reference
(oracle/_ref): a table string in the format GraphOptimizer writes into `prf` (GraphOptimizer.h:355-365), a degenerate
one, and the epsilon the reference returns for a sweep of expected accuracies (inside, below and above the table).
    python tests/golden/make_golden_accuracy_table.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TABLE = ("-0.9:0.05,-0.7:0.21,-0.5:0.43,-0.4:0.57,-0.3:0.69,-0.2:0.8,-0.1:0.875,-0.05:0.905,0:0.93,0.02:0.945,"
         "0.04:0.957,0.06:0.967,0.08:0.975,0.1:0.981,0.15:0.99,0.2:0.9945,0.3:0.998,0.4:0.9993")

if __name__ == "__main__":
    po.build(ref=True)
    R = po.Ref()
    acc = [0.01, 0.05, 0.3, 0.5, 0.57, 0.8, 0.9, 0.93, 0.95, 0.97, 0.99, 0.995, 0.9993, 0.9999, 1.0, 1.7]
    out = {"table": TABLE, "cases": [[a, float(R.epsilon_from_accuracy_table(TABLE, a))] for a in acc], "errors": {}}
    for name, t in (("empty", ""), ("two_points", "0:0.9,0.1:0.95"), ("bad_token", "0:0.9,0.1:0.95:1,0.2:0.99")):
        try:
            R.epsilon_from_accuracy_table(t, 0.9)
            out["errors"][name] = [t, None]
        except RuntimeError as ex:
            out["errors"][name] = [t, str(ex).split(": ", 1)[-1] if ".h:" in str(ex).split(": ", 1)[0] else str(ex)]
    json.dump(out, open(os.path.join(OUT, "accuracy_table.json"), "w"), indent=1)
    print(json.dumps(out)[:600])

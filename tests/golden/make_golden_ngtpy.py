#!/usr/bin/env python
"""Golden records for the `ngtpy` module, produced by the UNMODIFIED reference module (python/src/ngtpy.cpp compiled by
oracle/Makefile into oracle/_ref/ngtpy/):
  ngtpy_surface.json   the pybind11 signature line of every function and method (names, keyword arguments, defaults)
  ngtpy_scenario.json  what tests/ngtpy_scenario.py returns when it runs on the reference's module
  ngtpy_sample.txt     what the reference's own python/sample/sample.py prints on the reference's module
    python tests/golden/make_golden_ngtpy.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def surface(mod):
    """{qualified name: first docstring line} of the module's functions and of its classes' methods."""
    s = {}
    for name in sorted(dir(mod)):
        obj = getattr(mod, name)
        if name.startswith("_"):
            continue
        if isinstance(obj, type):
            for m in sorted(vars(obj)):
                doc = getattr(getattr(obj, m), "__doc__", None)
                if (m == "__init__" or not m.startswith("_")) and doc:
                    s[name + "." + m] = doc.strip().split("\n")[0]
        elif callable(obj):
            s[name] = obj.__doc__.strip().split("\n")[0]
    return s


if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ngtpy"))
    import ngtpy                                   # the reference's
    assert "_ref" in ngtpy.__file__
    from ngt_b200 import index_io
    import ngtpy_scenario
    json.dump(surface(ngtpy), open(os.path.join(OUT, "ngtpy_surface.json"), "w"), indent=1, sort_keys=True)
    tmp = tempfile.mkdtemp(prefix="ngt-golden-ngtpy-")
    try:
        rec = ngtpy_scenario.run(ngtpy, OUT, tmp, index_io.read_graph)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    json.dump(rec, open(os.path.join(OUT, "ngtpy_scenario.json"), "w"))
    tmp = tempfile.mkdtemp(prefix="ngt-golden-ngtpy-")
    try:
        done = ngtpy_scenario.run_reference_sample(os.path.join(ROOT, "oracle", "_ref", "ngtpy"), OUT, tmp,
                                                   os.path.join(ROOT, "oracle", "_ref", "python", "sample", "sample.py"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    assert done.returncode == 0, done.stderr
    open(os.path.join(OUT, "ngtpy_sample.txt"), "w").write(done.stdout)
    print({k: (list(v) if isinstance(v, dict) else v) for k, v in rec.items()})
    print(os.path.getsize(os.path.join(OUT, "ngtpy_scenario.json")), "bytes")

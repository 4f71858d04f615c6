"""Graph surgery of ONNG construction on the device: reconstructGraph + adjustPathsEffectively
(lib/NGT/GraphReconstructor.h:425-561, 197-386) against the graphs the unmodified reference wrote
(tests/golden/adjust_paths.npz) and against the sequential restatement in oracle/pyoracle.py on larger graphs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _lists(rp, col, dist):
    rp, col, dist = np.asarray(rp), np.asarray(col), np.asarray(dist)
    return [[(float(dist[e]), int(col[e])) for e in range(int(rp[i]), int(rp[i + 1]))] for i in range(len(rp) - 1)]


def _adjust_on_device(rp, col, dist, min_edges=0):
    import torch
    from ngt_b200 import build
    dev = torch.device("cuda", 0)
    out = build.adjust_paths(torch.from_numpy(np.asarray(rp).astype(np.int64)).to(dev),
                             torch.from_numpy(np.asarray(col).astype(np.int32)).to(dev),
                             torch.from_numpy(np.asarray(dist).astype(np.float32)).to(dev), min_edges, with_stats=True)
    return out[0].cpu().numpy(), out[1].cpu().numpy().astype(np.uint32), out[2].cpu().numpy(), out[3]


@pytest.mark.parametrize("key", ["sift_o5_i20", "sift_o10_i40", "glove_o5_i20", "glove_o10_i40"])
def test_adjust_paths_equals_the_reference(eng, key):
    z = np.load(os.path.join(GOLDEN, "adjust_paths.npz"))
    rp, col, dist, st = _adjust_on_device(z[key + "_in_row_ptr"], z[key + "_in_col"], z[key + "_in_dist"])
    assert (rp == z[key + "_adj_row_ptr"].astype(np.int64)).all()
    assert (col == z[key + "_adj_col"]).all()
    assert (dist.view(np.uint32) == z[key + "_adj_dist"].view(np.uint32)).all()
    assert st["removed"] == len(z[key + "_in_col"]) - len(z[key + "_adj_col"])


@pytest.mark.parametrize("min_edges", [0, 3, 12])
def test_adjust_paths_min_edges_and_long_lists(eng, min_edges):
    """A device-built ONNG-style graph with long reverse lists (hub nodes beyond the shared-memory staging limit of
    the candidate kernel are made by a duplicated cluster centre) against the sequential restatement."""
    import torch
    from ngt_b200 import build, synth
    n = 4000
    base = synth.make("sift", n, 7)
    base[1:2600] = base[0] + np.random.RandomState(3).randint(0, 3, size=(2599, base.shape[1]))   # one dense blob
    ix = eng.GpuIndex(po.FLOAT, po.L2, base.shape[1])
    ix.set_objects(base)
    ids, dists, counts = build.knn_graph(ix, 24)
    rp, col, dist = build.reconstruct_graph(ids, dists, counts, 6, 24)
    deg = (rp[2:] - rp[1:-1])
    out_rp, out_col, out_dist, st = build.adjust_paths(rp, col, dist, min_edges, with_stats=True)
    ref = po.adjust_paths_loop(rp.cpu().numpy(), col.cpu().numpy().astype(np.uint32), dist.cpu().numpy(), min_edges)
    got = _lists(out_rp.cpu().numpy(), out_col.cpu().numpy().astype(np.uint32), out_dist.cpu().numpy())
    assert got == ref
    assert st["removed"] > 0 and int(deg.max()) > 0
    kept_deg = out_rp[2:] - out_rp[1:-1]
    assert bool(((kept_deg >= torch.clamp(deg, max=min_edges)) | (deg == 0)).all())
    ix.close()


def test_adjust_paths_empty_and_edgeless(eng):
    import torch
    from ngt_b200 import build
    dev = torch.device("cuda", 0)
    rp = torch.zeros(12, dtype=torch.int64, device=dev)
    out = build.adjust_paths(rp, torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, device=dev))
    assert out[1].numel() == 0 and int(out[0][-1]) == 0
    with pytest.raises(eng.NgtGpuError):
        build.adjust_paths(rp.cpu(), torch.zeros(0, dtype=torch.int32), torch.zeros(0))

"""Parity assertions shared by the oracle-pin (CPU) and CUDA (GPU) tests.

Bars: bit-exact ids and float bit patterns for uint8 L2, Hamming and for
integer-valued float data; for general float data
    |d - d_ref| <= RTOL * max(|d_ref|, 1e-6)            RTOL = 1e-6
and id lists equal up to swaps/substitutions among entries whose reference distances differ by
<= TIE = 2e-6 relative (the reference itself is built with -Ofast, so its own float results move in
the last ulp between compilers/ISAs).
"""
import numpy as np

RTOL = 1e-6
TIE = 2e-6

EPS_GRID = (0.0, 0.1, 0.3)
EDGE_GRID = (-1, 0, 5, -2)
INTEGER_EXACT = ("u8l2", "ham", "f32l2", "onng")           # bit-exact cases in synth.npz
FLOAT_CASES = ("glove_l2", "glove_cos", "glove_ang", "glove_ncos", "glove_nang", "glove_nl2", "gist_l2")
NORMALIZED = (5, 6, 9)                                      # ObjectSpace.h:166-180


def grid_key(tag, eps, es):
    return "%s_e%02d_s%d" % (tag, int(round(eps * 100)), es)


def assert_bit_exact(ids, dists, counts, ref_ids, ref_dists, ref_counts=None, what=""):
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    if ref_counts is not None:
        assert (np.asarray(counts) == np.asarray(ref_counts)).all(), what + ": result counts differ"
    for q in range(ids.shape[0]):
        c = int(counts[q]) if counts is not None else ids.shape[1]
        assert (ids[q, :c] == ref_ids[q, :c]).all(), "%s: ids differ for query %d\n%s\n%s" % (
            what, q, ids[q, :c], ref_ids[q, :c])
        a = np.asarray(dists[q, :c], np.float32).view(np.uint32)
        b = np.asarray(ref_dists[q, :c], np.float32).view(np.uint32)
        assert (a == b).all(), "%s: distance bits differ for query %d" % (what, q)


def assert_float_parity(ids, dists, counts, ref_ids, ref_dists, ref_counts=None, what="", rtol=RTOL, tie=TIE):
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    dists, ref_dists = np.asarray(dists, np.float64), np.asarray(ref_dists, np.float64)
    for q in range(ids.shape[0]):
        c = int(counts[q]) if counts is not None else ids.shape[1]
        rc = int(ref_counts[q]) if ref_counts is not None else c
        assert c == rc, "%s: result count differs for query %d (%d vs %d)" % (what, q, c, rc)
        d, rd = dists[q, :c], ref_dists[q, :c]
        scale = np.maximum(np.abs(rd), 1e-6)
        assert (np.abs(d - rd) <= rtol * scale).all(), "%s: distances off for query %d: max rel %g" % (
            what, q, (np.abs(d - rd) / scale).max())
        for i in range(c):
            if ids[q, i] == ref_ids[q, i]:
                continue
            # allowed only as a near-tie: the id sits elsewhere in the reference list at (almost) the
            # same distance, or it replaces the boundary entry at (almost) the boundary distance
            where = np.nonzero(ref_ids[q, :c] == ids[q, i])[0]
            if where.size:
                ok = abs(rd[where[0]] - rd[i]) <= tie * max(abs(rd[i]), 1e-6)
            else:
                ok = abs(d[i] - rd[c - 1]) <= tie * max(abs(rd[c - 1]), 1e-6)
            assert ok, "%s: query %d rank %d: id %d vs reference %d is not a near-tie" % (
                what, q, i, ids[q, i], ref_ids[q, i])

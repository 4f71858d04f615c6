"""NGT on-disk index directory (`prf`, `obj`, `grp`[, `tre`]) <-> arrays.

`prf` is `key<TAB>value` text (lib/NGT/Common.h:649-665; keys lib/NGT/Index.h:105-154 and
lib/NGT/Graph.h:423-454). `obj`/`grp` are streamed by the native readers/writers in
ngt_b200/csrc/index_io.cu. The DVP-tree file `tre` is not produced: indexes written here carry
`IndexType Graph`, which the reference opens without a tree (lib/NGT/Index.cpp:93-111), and seed selection
on the device replaces the tree at search time anyway. Reading ignores `tre`.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

OBJECT_TYPE_NAMES = {"Integer-1": _lib.OBJECT_UINT8, "Float-4": _lib.OBJECT_FLOAT}
DISTANCE_TYPE_NAMES = {
    "L1": _lib.DISTANCE_L1, "L2": _lib.DISTANCE_L2, "Hamming": _lib.DISTANCE_HAMMING, "Angle": _lib.DISTANCE_ANGLE,
    "Cosine": _lib.DISTANCE_COSINE, "NormalizedAngle": _lib.DISTANCE_NORMALIZED_ANGLE,
    "NormalizedCosine": _lib.DISTANCE_NORMALIZED_COSINE, "Jaccard": _lib.DISTANCE_JACCARD,
    "NormalizedL2": _lib.DISTANCE_NORMALIZED_L2,
}

# defaults of NGT::Property (lib/NGT/Index.h:60-103, lib/NGT/Graph.h:385-420) as `ngt create` writes them
DEFAULT_PRF = {
    "AccuracyTable": "", "BatchSizeForCreation": "200", "BuildTimeLimit": "0", "DatabaseType": "Memory",
    "Dimension": "0", "DistanceType": "L2", "DynamicEdgeSizeBase": "30", "DynamicEdgeSizeRate": "20",
    "EdgeSizeForCreation": "10", "EdgeSizeForSearch": "40", "EdgeSizeLimitForCreation": "5",
    "EpsilonForCreation": "0.1", "GraphType": "ANNG", "IncomingEdge": "80",
    "IncrimentalEdgeSizeLimitForTruncation": "0", "IndexType": "Graph", "ObjectAlignment": "False",
    "ObjectType": "Float-4", "OutgoingEdge": "10", "PathAdjustmentInterval": "0", "PrefetchOffset": "0",
    "PrefetchSize": "0", "SeedSize": "10", "SeedType": "None", "ThreadPoolSize": "24",
    "TruncationThreadPoolSize": "8",
}

_io_ready = False


def _io():
    global _io_ready
    lib = _lib.load()
    if not _io_ready:
        P = C.c_void_p
        lib.ngtgpu_io_obj_info.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.ngtgpu_io_read_obj.argtypes = [C.c_char_p, C.c_uint32, P, P]
        lib.ngtgpu_io_write_obj.argtypes = [C.c_char_p, C.c_uint32, P, C.c_uint64, P]
        lib.ngtgpu_io_grp_info.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.ngtgpu_io_read_grp.argtypes = [C.c_char_p, P, P, P, P]
        lib.ngtgpu_io_write_grp.argtypes = [C.c_char_p, C.c_uint64, P, P, P, P]
        _io_ready = True
    return lib


def read_prf(path):
    prop = {}
    with open(os.path.join(path, "prf")) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            key, _, value = line.partition("\t")
            prop[key] = value
    return prop


def write_prf(path, prop):
    with open(os.path.join(path, "prf"), "w") as f:
        for key in sorted(prop):
            f.write("%s\t%s\n" % (key, prop[key]))


def object_dtype(prop):
    return np.uint8 if OBJECT_TYPE_NAMES[prop["ObjectType"]] == _lib.OBJECT_UINT8 else np.float32


def read_objects(path, prop):
    """-> (rows [n, dim] of the object type for ids 1..n, present [n+1] uint8)"""
    dt = object_dtype(prop)
    dim = int(prop["Dimension"])
    rec = dim * np.dtype(dt).itemsize
    lib = _io()
    slots, present = C.c_uint64(0), C.c_uint64(0)
    p = os.path.join(path, "obj").encode()
    _lib.check(lib.ngtgpu_io_obj_info(p, rec, C.byref(slots), C.byref(present)))
    n = max(int(slots.value) - 1, 0)
    rows = np.zeros((n, dim), dt)
    pres = np.zeros(n + 1, np.uint8)
    _lib.check(lib.ngtgpu_io_read_obj(p, rec, rows.ctypes.data, pres.ctypes.data))
    return rows, pres


def write_objects(path, rows, present=None):
    rows = np.ascontiguousarray(rows)
    pp = None
    if present is not None:
        present = np.ascontiguousarray(present, np.uint8)
        pp = present.ctypes.data
    _lib.check(_io().ngtgpu_io_write_obj(os.path.join(path, "obj").encode(), rows.shape[1] * rows.itemsize,
                                         rows.ctypes.data, rows.shape[0], pp))


def read_graph(path):
    """-> (row_ptr [slots+1] over ids 0..slots-1, col, dist, present [slots])"""
    lib = _io()
    slots, nnz = C.c_uint64(0), C.c_uint64(0)
    p = os.path.join(path, "grp").encode()
    _lib.check(lib.ngtgpu_io_grp_info(p, C.byref(slots), C.byref(nnz)))
    row_ptr = np.zeros(int(slots.value) + 1, np.uint64)
    col = np.zeros(max(int(nnz.value), 1), np.uint32)
    dist = np.zeros(max(int(nnz.value), 1), np.float32)
    pres = np.zeros(int(slots.value), np.uint8)
    _lib.check(lib.ngtgpu_io_read_grp(p, row_ptr.ctypes.data, col.ctypes.data, dist.ctypes.data, pres.ctypes.data))
    return row_ptr, col[:int(nnz.value)], dist[:int(nnz.value)], pres


def write_graph(path, row_ptr, col, dist, present=None):
    """row_ptr over ids 0..n (n+2 entries)."""
    row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
    col = np.ascontiguousarray(col, np.uint32)
    dist = np.ascontiguousarray(dist, np.float32)
    pp = None
    if present is not None:
        present = np.ascontiguousarray(present, np.uint8)
        pp = present.ctypes.data
    _lib.check(_io().ngtgpu_io_write_grp(os.path.join(path, "grp").encode(), row_ptr.size - 2, row_ptr.ctypes.data,
                                         col.ctypes.data, dist.ctypes.data, pp))

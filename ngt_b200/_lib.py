"""ctypes binding of libngtgpu.so -- the C ABI declared in include/ngtgpu.h.

The library is built in-tree by `make -C ngt_b200/csrc` (see __graft_entry__.build). There is no
Python/CPU fallback: if the shared object is missing or no sm_100 device is present, every entry point
raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("NGTGPU_SO", os.path.join(HERE, "libngtgpu.so"))   # NGTGPU_SO: development builds of the same library

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE, ERR_ZERO_VECTOR = range(6)

# lib/NGT/ObjectSpace.h:182-186 and :166-180
OBJECT_UINT8, OBJECT_FLOAT = 1, 2
DISTANCE_L1, DISTANCE_L2, DISTANCE_HAMMING, DISTANCE_ANGLE, DISTANCE_COSINE = 0, 1, 2, 3, 4
DISTANCE_NORMALIZED_ANGLE, DISTANCE_NORMALIZED_COSINE, DISTANCE_JACCARD = 5, 6, 7
DISTANCE_NORMALIZED_L2 = 9


class SearchParams(C.Structure):
    _fields_ = [("size", C.c_uint32), ("epsilon", C.c_float), ("radius", C.c_float), ("edge_size", C.c_int64)]


class NgtGpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


# every symbol include/ngtgpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "ngtgpu_last_error": (C.c_char_p, []),
    "ngtgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ngtgpu_index_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_uint32]),
    "ngtgpu_index_destroy": (C.c_int, [_P]),
    "ngtgpu_index_set_objects": (C.c_int, [_P, _P, C.c_uint64, C.c_int, C.c_int]),
    "ngtgpu_index_set_removed": (C.c_int, [_P, _P, C.c_uint64]),
    "ngtgpu_index_set_graph": (C.c_int, [_P, _P, _P, C.c_int]),
    "ngtgpu_index_set_search_property": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64]),
    "ngtgpu_index_set_search_workspace": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "ngtgpu_index_set_onchip_tiers": (C.c_int, [_P, C.c_int]),
    "ngtgpu_index_set_fast_kernel": (C.c_int, [_P, C.c_int]),
    "ngtgpu_index_set_seed_fusion": (C.c_int, [_P, C.c_int]),
    "ngtgpu_index_set_fast_shape": (C.c_int, [_P, C.c_int, C.c_int]),
    "ngtgpu_index_set_stage_bytes": (C.c_int, [_P, C.c_uint32]),
    "ngtgpu_index_set_tensor_core": (C.c_int, [_P, C.c_int]),
    "ngtgpu_index_tensor_core_batches": (C.c_uint64, [_P]),
    "ngtgpu_index_build_seed_table": (C.c_int, [_P, C.c_uint32, C.c_uint64]),
    "ngtgpu_index_build_seed_table_range": (C.c_int, [_P, C.c_uint32, C.c_uint64, C.c_uint64]),
    "ngtgpu_index_set_seed_table_ids": (C.c_int, [_P, _P, C.c_uint32]),
    "ngtgpu_index_size": (C.c_uint64, [_P]),
    "ngtgpu_index_padded_dimension": (C.c_uint32, [_P]),
    "ngtgpu_index_get_object": (C.c_int, [_P, C.c_uint32, _P]),
    "ngtgpu_index_get_objects": (C.c_int, [_P, C.c_uint32, C.c_uint64, _P]),
    "ngtgpu_epsilon_from_accuracy_table": (C.c_int, [C.c_char_p, C.c_double, C.POINTER(C.c_float)]),
    "ngtgpu_index_pairwise_distances": (C.c_int, [_P, _P, C.c_uint32, _P]),
    "ngtgpu_index_launch_count": (C.c_uint64, [_P]),
    "ngtgpu_index_last_overflows": (C.c_uint64, [_P]),
    "ngtgpu_search": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.POINTER(SearchParams), _P, C.c_uint32, _P, _P, _P, _P]),
    "ngtgpu_search_device": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.POINTER(SearchParams), _P, C.c_uint32, _P, _P,
                                       _P, _P, _P]),
    "ngtgpu_linear_search": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.c_uint32, C.c_float, _P, _P, _P]),
    "ngtgpu_graph_reconstruct": (C.c_int, [C.c_uint64, _P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint64, _P, _P, _P,
                                           C.POINTER(C.c_uint64), _P]),
    "ngtgpu_index_refine_anng": (C.c_int, [_P, C.c_float, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64,
                                           _P, _P, _P, C.POINTER(C.c_uint64)]),
    "ngtgpu_graph_from_knn_table": (C.c_int, [C.c_uint64, _P, _P, _P, C.c_uint32, _P, C.c_int, C.c_uint64, _P, _P, _P,
                                              C.POINTER(C.c_uint64), _P]),
    "ngtgpu_graph_select_edges": (C.c_int, [C.c_uint64, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_uint64), _P]),
    "ngtgpu_index_insert_batch": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_int64, C.c_uint32, C.c_uint32,
                                            C.c_uint64, C.c_uint64, _P, _P, _P, C.POINTER(C.c_uint64)]),
    "ngtgpu_graph_adjust_paths": (C.c_int, [C.c_uint64, _P, _P, _P, C.c_uint32, _P, C.POINTER(C.c_uint64), _P]),
    "ngtgpu_linear_search_device": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.c_uint32, C.c_float, _P, _P, _P, _P]),
    "ngtgpu_index_build_onng": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, _P, C.POINTER(C.c_double)]),
    "ngtgpu_device_free": (C.c_int, [_P]),
    "ngtgpu_device_copy": (C.c_int, [_P, _P, C.c_uint64]),
    # multi-GPU (shard.cu)
    "ngtgpu_comm_get_unique_id": (C.c_int, [_P]),
    "ngtgpu_comm_create": (C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int, C.c_int]),
    "ngtgpu_comm_destroy": (C.c_int, [_P]),
    "ngtgpu_comm_set_timing": (C.c_int, [_P, C.c_int]),
    "ngtgpu_comm_pop_timing": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "ngtgpu_shard_search_device": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint32, C.POINTER(SearchParams), C.c_uint32, C.c_uint32,
                                             _P, _P, _P, _P]),
    "ngtgpu_shard_linear_search_device": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32,
                                                    _P, _P, _P, _P]),
    "ngtgpu_sharded_create": (C.c_int, [C.POINTER(_P), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_uint32]),
    "ngtgpu_sharded_destroy": (C.c_int, [_P]),
    "ngtgpu_sharded_set_objects": (C.c_int, [_P, _P, C.c_uint64, C.c_int]),
    "ngtgpu_sharded_build_onng": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int64, C.c_uint32]),
    "ngtgpu_sharded_shard_count": (C.c_int, [_P]),
    "ngtgpu_sharded_shard": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ngtgpu_sharded_search": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.POINTER(SearchParams), C.c_uint32, _P, _P, _P]),
    "ngtgpu_sharded_linear_search": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.c_uint32, C.c_float, _P, _P, _P]),
    "ngtgpu_sharded_last_timing": (C.c_int, [_P, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Loads libngtgpu.so (once). Raises if it has not been built: the product has no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NgtGpuError(ERR_NO_DEVICE, "%s is missing: build it with `make -C ngt_b200/csrc` "
                          "(there is no CPU or PyTorch fallback for the NGT hot path)" % SO_PATH)
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)     # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().ngtgpu_last_error()
        raise NgtGpuError(rc, (msg or b"").decode(errors="replace") or "ngtgpu error %d" % rc)


def device_count():
    n = C.c_int(0)
    check(load().ngtgpu_device_count(C.byref(n)))
    return n.value

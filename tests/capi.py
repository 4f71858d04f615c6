"""ctypes declarations of NGT's C API (lib/NGT/Capi.h:28-212) as python/ngt/base.py binds it, against any library
that exports it -- here ngt_b200/libngtgpu.so."""
import ctypes as C

import numpy as np


class ObjectDistance(C.Structure):
    _fields_ = [("id", C.c_uint32), ("distance", C.c_float)]


class Query(C.Structure):
    _fields_ = [("query", C.POINTER(C.c_float)), ("size", C.c_size_t), ("epsilon", C.c_float), ("accuracy", C.c_float),
                ("radius", C.c_float), ("edge_size", C.c_size_t)]


class AnngEdgeOptimizationParameter(C.Structure):     # Capi.h:49-58
    _fields_ = [("no_of_queries", C.c_size_t), ("no_of_results", C.c_size_t), ("no_of_threads", C.c_size_t),
                ("target_accuracy", C.c_float), ("target_no_of_objects", C.c_size_t), ("no_of_sample_objects", C.c_size_t),
                ("max_of_no_of_edges", C.c_size_t), ("log", C.c_bool)]


# every function lib/NGT/Capi.h:60-212 declares (67 names)
CAPI_H_FUNCTIONS = """ngt_open_index ngt_create_graph_and_tree ngt_create_graph_and_tree_in_memory
ngt_create_property ngt_save_index ngt_get_property ngt_get_property_dimension ngt_set_property_dimension
ngt_set_property_edge_size_for_creation ngt_set_property_edge_size_for_search ngt_get_property_object_type
ngt_is_property_object_type_float ngt_is_property_object_type_integer ngt_set_property_object_type_float
ngt_set_property_object_type_integer ngt_set_property_distance_type_l1 ngt_set_property_distance_type_l2
ngt_set_property_distance_type_angle ngt_set_property_distance_type_hamming ngt_set_property_distance_type_jaccard
ngt_set_property_distance_type_cosine ngt_set_property_distance_type_normalized_angle
ngt_set_property_distance_type_normalized_cosine ngt_create_empty_results ngt_search_index ngt_search_index_as_float
ngt_search_index_with_query ngt_linear_search_index ngt_linear_search_index_as_float ngt_linear_search_index_with_query
ngt_get_size ngt_get_result_size ngt_get_result ngt_insert_index ngt_append_index ngt_insert_index_as_float
ngt_append_index_as_float ngt_batch_append_index ngt_batch_insert_index ngt_create_index ngt_remove_index
ngt_get_object_space ngt_get_object_as_float ngt_get_object_as_integer ngt_destroy_results ngt_destroy_property
ngt_close_index ngt_get_property_edge_size_for_creation ngt_get_property_edge_size_for_search
ngt_get_property_distance_type ngt_create_error_object ngt_get_error_string ngt_clear_error_string
ngt_destroy_error_object ngt_create_optimizer ngt_optimizer_adjust_search_coefficients ngt_optimizer_execute
ngt_optimizer_set ngt_optimizer_set_minimum ngt_optimizer_set_extension ngt_optimizer_set_processing_modes
ngt_destroy_optimizer ngt_refine_anng ngt_get_edges ngt_get_object_repository_size
ngt_get_anng_edge_optimization_parameter ngt_optimize_number_of_edges""".split()


def load_reference_base_py(so_path):
    """The reference's own ctypes binding, python/ngt/base.py, UNMODIFIED (copied by oracle/Makefile into oracle/_ref/,
    or read where it lies under /root/reference), executed with ctypes.util.find_library("ngt") answering `so_path`.
    Returns the module, or None when no copy of the reference is at hand."""
    import ctypes.util
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(root, "oracle", "_ref", "python", "ngt", "base.py"), "/root/reference/python/ngt/base.py"):
        if os.path.exists(cand):
            break
    else:
        return None
    real = ctypes.util.find_library
    ctypes.util.find_library = lambda name: so_path if name == "ngt" else real(name)
    try:
        spec = importlib.util.spec_from_file_location("ngt_reference_base", cand)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        ctypes.util.find_library = real
    return mod


def bind(path):
    lib = C.CDLL(path)
    P, F, D = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_double)
    sig = {
        "ngt_create_error_object": (P, []), "ngt_get_error_string": (C.c_char_p, [P]), "ngt_clear_error_string": (None, [P]),
        "ngt_destroy_error_object": (None, [P]),
        "ngt_create_property": (P, [P]), "ngt_destroy_property": (None, [P]), "ngt_get_property": (C.c_bool, [P, P, P]),
        "ngt_get_property_dimension": (C.c_int32, [P, P]), "ngt_set_property_dimension": (C.c_bool, [P, C.c_int32, P]),
        "ngt_set_property_edge_size_for_creation": (C.c_bool, [P, C.c_int16, P]),
        "ngt_set_property_edge_size_for_search": (C.c_bool, [P, C.c_int16, P]),
        "ngt_get_property_edge_size_for_creation": (C.c_int16, [P, P]), "ngt_get_property_edge_size_for_search": (C.c_int16, [P, P]),
        "ngt_get_property_object_type": (C.c_int32, [P, P]), "ngt_get_property_distance_type": (C.c_int32, [P, P]),
        "ngt_is_property_object_type_float": (C.c_bool, [C.c_int32]), "ngt_is_property_object_type_integer": (C.c_bool, [C.c_int32]),
        "ngt_set_property_object_type_float": (C.c_bool, [P, P]), "ngt_set_property_object_type_integer": (C.c_bool, [P, P]),
        "ngt_set_property_distance_type_l2": (C.c_bool, [P, P]), "ngt_set_property_distance_type_cosine": (C.c_bool, [P, P]),
        "ngt_set_property_distance_type_hamming": (C.c_bool, [P, P]),
        "ngt_set_property_distance_type_normalized_cosine": (C.c_bool, [P, P]),
        "ngt_open_index": (P, [C.c_char_p, P]), "ngt_create_graph_and_tree": (P, [C.c_char_p, P, P]),
        "ngt_create_graph_and_tree_in_memory": (P, [P, P]), "ngt_save_index": (C.c_bool, [P, C.c_char_p, P]),
        "ngt_close_index": (None, [P]),
        "ngt_create_empty_results": (P, [P]), "ngt_destroy_results": (None, [P]), "ngt_get_result_size": (C.c_uint32, [P, P]),
        "ngt_get_size": (C.c_int32, [P, P]), "ngt_get_result": (ObjectDistance, [P, C.c_uint32, P]),
        "ngt_search_index": (C.c_bool, [P, D, C.c_int32, C.c_size_t, C.c_float, C.c_float, P, P]),
        "ngt_search_index_as_float": (C.c_bool, [P, F, C.c_int32, C.c_size_t, C.c_float, C.c_float, P, P]),
        "ngt_search_index_with_query": (C.c_bool, [P, Query, P, P]),
        "ngt_linear_search_index": (C.c_bool, [P, D, C.c_int32, C.c_size_t, P, P]),
        "ngt_linear_search_index_as_float": (C.c_bool, [P, F, C.c_int32, C.c_size_t, P, P]),
        "ngt_linear_search_index_with_query": (C.c_bool, [P, Query, P, P]),
        "ngt_insert_index": (C.c_uint32, [P, D, C.c_uint32, P]), "ngt_append_index": (C.c_uint32, [P, D, C.c_uint32, P]),
        "ngt_insert_index_as_float": (C.c_uint32, [P, F, C.c_uint32, P]), "ngt_append_index_as_float": (C.c_uint32, [P, F, C.c_uint32, P]),
        "ngt_batch_append_index": (C.c_bool, [P, F, C.c_uint32, P]),
        "ngt_batch_insert_index": (C.c_bool, [P, F, C.c_uint32, C.POINTER(C.c_uint32), P]),
        "ngt_create_index": (C.c_bool, [P, C.c_uint32, P]), "ngt_remove_index": (C.c_bool, [P, C.c_uint32, P]),
        "ngt_get_object_space": (P, [P, P]), "ngt_get_object_as_float": (F, [P, C.c_uint32, P]),
        "ngt_get_object_as_integer": (C.POINTER(C.c_uint8), [P, C.c_uint32, P]),
        "ngt_get_edges": (C.c_bool, [P, C.c_uint32, P, P]), "ngt_get_object_repository_size": (C.c_uint32, [P, P]),
        "ngt_refine_anng": (C.c_bool, [P, C.c_float, C.c_float, C.c_int, C.c_int, C.c_size_t, P]),
        "ngt_create_optimizer": (P, [C.c_bool, P]), "ngt_destroy_optimizer": (None, [P]),
        "ngt_optimizer_execute": (C.c_bool, [P, C.c_char_p, C.c_char_p, P]),
        "ngt_optimizer_adjust_search_coefficients": (C.c_bool, [P, C.c_char_p, P]),
        "ngt_optimizer_set": (C.c_bool, [P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double,
                                         C.c_double, P]),
        "ngt_optimizer_set_minimum": (C.c_bool, [P, C.c_int, C.c_int, C.c_int, C.c_int, P]),
        "ngt_optimizer_set_extension": (C.c_bool, [P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, P]),
        "ngt_optimizer_set_processing_modes": (C.c_bool, [P, C.c_bool, C.c_bool, C.c_bool, P]),
        # additive batch entry points
        "ngt_batch_search_index_as_float": (C.c_bool, [P, F, C.c_uint32, C.c_int32, C.c_size_t, C.c_float, C.c_float, C.c_int64,
                                                       C.POINTER(C.c_uint32), F, C.POINTER(C.c_uint32), P]),
        "ngt_batch_linear_search_index_as_float": (C.c_bool, [P, F, C.c_uint32, C.c_int32, C.c_size_t, C.c_float,
                                                              C.POINTER(C.c_uint32), F, C.POINTER(C.c_uint32), P]),
        "ngt_batch_search_index_as_uint8": (C.c_bool, [P, C.POINTER(C.c_uint8), C.c_uint32, C.c_int32, C.c_size_t, C.c_float,
                                                       C.c_float, C.c_int64, C.POINTER(C.c_uint32), F, C.POINTER(C.c_uint32), P]),
        "ngt_batch_linear_search_index_as_uint8": (C.c_bool, [P, C.POINTER(C.c_uint8), C.c_uint32, C.c_int32, C.c_size_t, C.c_float,
                                                              C.POINTER(C.c_uint32), F, C.POINTER(C.c_uint32), P]),
        # Capi.h:208,212
        "ngt_get_anng_edge_optimization_parameter": (AnngEdgeOptimizationParameter, []),
        "ngt_optimize_number_of_edges": (C.c_bool, [C.c_char_p, AnngEdgeOptimizationParameter, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._signatures = sig
    return lib


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def results_of(lib, r, err):
    n = lib.ngt_get_result_size(r, err)
    out = []
    for i in range(n):
        o = lib.ngt_get_result(r, i, err)
        out.append((int(o.id), float(np.float32(o.distance))))
    return out

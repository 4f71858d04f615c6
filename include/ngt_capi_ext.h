/*
 * ngt_capi_ext.h -- the ADDITIVE `ngt_*` entry points libngtgpu.so exports next to the 67 functions of the
 * reference's own C header (lib/NGT/Capi.h:60-212, which a caller keeps including unchanged: the library serves
 * those under their own names and signatures, csrc/capi.cu).
 *
 * Why they exist: the per-query calls of Capi.h (`ngt_search_index*`, Capi.h:108-118) cannot feed a GPU, so the batch
 * forms below take `nq` queries and fill flat result arrays. Same handle types, same error
 * convention as Capi.cpp:25-38: no exception crosses, the message "Capi : <func>() : Error: <what>" goes into *error
 * (or to stderr when error is NULL) and the call returns false / NULL / 0. Plain pointers and sizes only.
 */
#ifndef NGT_CAPI_EXT_H
#define NGT_CAPI_EXT_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the opaque handles of lib/NGT/Capi.h:28-33 */
#ifndef NGT_CAPI_EXT_NO_TYPEDEFS
typedef void *NGTIndex;
typedef void *NGTError;
typedef void *NGTOptimizer;
typedef void *NGTProperty;
#endif

/* Batch form of ngt_search_index_as_float (Capi.h:110 -> Capi.cpp:377-404 -> GraphIndex::search, Index.h:1140-1179).
 * queries: nq x query_dim floats (host). size = k, epsilon / radius / edge_size as the fields of NGTQuery
 * (Capi.h:40-47; radius < 0 = none; edge_size -1 = the index property, 0 = all edges, -2 = dynamic, Graph.h:675-692).
 * Out (host, caller-owned): ids and dists nq x size, ascending by (distance, id), 1-based ids; counts[q] = number of
 * valid entries of row q (fewer than size when the graph yields fewer). */
bool ngt_batch_search_index_as_float(NGTIndex index, const float *queries, uint32_t nq, int32_t query_dim, size_t size,
                                     float epsilon, float radius, int64_t edge_size, uint32_t *ids, float *dists,
                                     uint32_t *counts, NGTError error);
/* The same with byte queries, for Integer-1 indexes (objects are stored as bytes, ObjectSpaceRepository.h:392-420). */
bool ngt_batch_search_index_as_uint8(NGTIndex index, const uint8_t *queries, uint32_t nq, int32_t query_dim, size_t size,
                                     float epsilon, float radius, int64_t edge_size, uint32_t *ids, float *dists,
                                     uint32_t *counts, NGTError error);

/* Batch form of ngt_linear_search_index_as_float (Capi.h:116 -> ObjectSpaceRepository::linearSearch,
 * ObjectSpaceRepository.h:466-502): the exact k nearest of every query; radius < 0 = no radius filter. */
bool ngt_batch_linear_search_index_as_float(NGTIndex index, const float *queries, uint32_t nq, int32_t query_dim,
                                            size_t size, float radius, uint32_t *ids, float *dists, uint32_t *counts,
                                            NGTError error);
bool ngt_batch_linear_search_index_as_uint8(NGTIndex index, const uint8_t *queries, uint32_t nq, int32_t query_dim,
                                            size_t size, float radius, uint32_t *ids, float *dists, uint32_t *counts,
                                            NGTError error);

/* ngt_open_index (Capi.h:60) with the rows of the index sharded over `devices` (one process, one NCCL communicator per
 * handle; per-shard graph and seed table, one all-gather of the per-shard top-k lists and a device merge per call:
 * SURVEY.md section 8e). The handle is read-only. The environment variable NGTGPU_DEVICES=0,1,.. does the same for
 * plain ngt_open_index. */
NGTIndex ngt_open_index_sharded(const char *index_path, const int *devices, int n_devices, NGTError error);

/* DistanceTypeNormalizedL2 (lib/NGT/ObjectSpace.h:166-180) for a property handle: Capi.h:86-104 has setters for the
 * other distance types only, ngtpy.create accepts "Normalized L2" (python/src/ngtpy.cpp:76-77). */
bool ngt_set_property_distance_type_normalized_l2(NGTProperty prop, NGTError error);

/* GraphOptimizer::shortcutReduction (GraphOptimizer.h:73,640-650). Capi.h:188's ngt_optimizer_set_processing_modes
 * carries the three tuning flags only; ngtpy's Optimizer.set_processing_modes (python/src/ngtpy.cpp:593-597) also
 * switches the shortcut reduction of execute(). */
bool ngt_optimizer_set_shortcut_reduction(NGTOptimizer optimizer, bool shortcutReduction, NGTError error);

/* Sum of SearchContainer::distanceComputationCount (Graph.cpp:592; Common.h:2062) over the single-query searches the
 * handle has served: what ngtpy.Index.get_num_of_distance_computations reports (python/src/ngtpy.cpp:181,349). */
uint64_t ngt_get_number_of_distance_computations(NGTIndex index, NGTError error);

#ifdef __cplusplus
}
#endif
#endif /* NGT_CAPI_EXT_H */

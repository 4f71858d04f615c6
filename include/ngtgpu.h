/*
 * ngtgpu.h -- C ABI of the B200 (sm_100a) engine for NGT's data-parallel hot path.
 *
 * This is the whole drop-in boundary: plain pointers and sizes, integer status codes, no
 * exceptions, no torch/C++ types. The reference (NGT v1.13.8, /root/reference) would bind these
 * symbols from the seams listed next to each function; INTEGRATION.md shows the few lines a
 * maintainer adds on the reference side. There is no CPU fallback: without a CUDA device every
 * call fails with NGTGPU_ERR_NO_DEVICE.
 *
 * Conventions shared with the reference
 *   - object ids are 1-based, id 0 is the dummy slot          (lib/NGT/Common.h:1704-1720)
 *   - results are ascending by (distance, id)                  (lib/NGT/Common.h:1946-1959)
 *   - distances are float, the id type is uint32               (lib/NGT/Common.h:46-47)
 *   - object/distance type codes are the reference's enums     (lib/NGT/ObjectSpace.h:166-186)
 *   - rows are zero padded to 16 elements                      (lib/NGT/ObjectSpace.h:249)
 *
 * Every function returns NGTGPU_OK (0) or an NGTGPU_ERR_* code; ngtgpu_last_error() returns the
 * message of the calling thread's last failure (the `Capi : f() : Error: what` string the C API
 * layer of the reference puts into NGTError, lib/NGT/Capi.cpp:25-38, is built from it).
 */
#ifndef NGTGPU_H
#define NGTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGTGPU_OK 0
#define NGTGPU_ERR_INVALID 1    /* bad argument / unsupported type combination            */
#define NGTGPU_ERR_CUDA 2       /* a CUDA runtime call failed                              */
#define NGTGPU_ERR_NO_DEVICE 3  /* no usable sm_100 device: there is no CPU fallback       */
#define NGTGPU_ERR_STATE 4      /* call order (e.g. search before objects/graph are set)   */
#define NGTGPU_ERR_ZERO_VECTOR 5 /* zero vector given to a normalised space (ObjectSpace.h:256-260) */

/* lib/NGT/ObjectSpace.h:182-186 */
#define NGTGPU_OBJECT_UINT8 1
#define NGTGPU_OBJECT_FLOAT 2
/* lib/NGT/ObjectSpace.h:166-180 */
#define NGTGPU_DISTANCE_L2 1
#define NGTGPU_DISTANCE_HAMMING 2
#define NGTGPU_DISTANCE_ANGLE 3
#define NGTGPU_DISTANCE_COSINE 4
#define NGTGPU_DISTANCE_NORMALIZED_ANGLE 5
#define NGTGPU_DISTANCE_NORMALIZED_COSINE 6
#define NGTGPU_DISTANCE_NORMALIZED_L2 9

typedef struct ngtgpu_index ngtgpu_index; /* opaque; one per (index, device) */

/* Per-call search settings == the fields of NGT::SearchContainer a caller can set
 * (lib/NGT/Common.h:2029-2046) == NGTQuery (lib/NGT/Capi.h:40-47). */
typedef struct {
  uint32_t size;      /* k; 0 returns no results (Index.h:1141-1144)                          */
  float epsilon;      /* explorationCoefficient = (float)(epsilon + 1.0) (Common.h:2041)      */
  float radius;       /* < 0 means unbounded (FLT_MAX), as Capi.cpp:384-386                   */
  int64_t edge_size;  /* -1 index property, 0 all edges, >0 cap, -2 dynamic (Graph.h:675-692) */
} ngtgpu_search_params;

const char *ngtgpu_last_error(void);
int ngtgpu_device_count(int *count);

/* ---- index life cycle: replaces GraphIndex construction + loadIndex
 *      (lib/NGT/Index.cpp:587-606, Index.h:665-695) for the device-resident copy ---------------- */
int ngtgpu_index_create(ngtgpu_index **out, int device, int object_type, int distance_type, uint32_t dimension);
int ngtgpu_index_destroy(ngtgpu_index *index);

/* Objects 1..n, row-major, `dimension` elements of the object type each (the payload of `obj`,
 * lib/NGT/Common.h:1776-1793). Rows are padded on the device; for Normalized* distance types they
 * are stored as given unless `normalize` != 0, in which case each row is divided by its L2 norm as
 * ObjectSpace::normalize does (lib/NGT/ObjectSpace.h:251-266). `on_device`: pointer is in HBM. */
int ngtgpu_index_set_objects(ngtgpu_index *index, const void *objects, uint64_t n, int normalize, int on_device);
/* ids whose slots are empty ('-' records): skipped by linearSearch (ObjectSpaceRepository.h:485). */
int ngtgpu_index_set_removed(ngtgpu_index *index, const uint32_t *ids, uint64_t count);
/* Adjacency lists in CSR over ids 0..n: row_ptr has n+2 entries, list of id is
 * col[row_ptr[id] .. row_ptr[id+1]) in `grp` order, i.e. ascending (distance,id)
 * (lib/NGT/Graph.h:62-183, ObjectSpace.h:29). */
int ngtgpu_index_set_graph(ngtgpu_index *index, const uint64_t *row_ptr, const uint32_t *col, int on_device);
/* NeighborhoodGraph::Property fields the search path reads (lib/NGT/Graph.h:383-454). */
int ngtgpu_index_set_search_property(ngtgpu_index *index, int64_t edge_size_for_search,
                                     int64_t dynamic_edge_size_base, int64_t dynamic_edge_size_rate);
/* Batched on-device seed selection that stands in for the DVP-tree leaf lookup
 * (lib/NGT/Index.h:1524-1567, Tree.cpp:400-563): `n_pivots` sampled objects form a table; each query's
 * seeds are its nearest `seed_size` pivots. */
int ngtgpu_index_build_seed_table(ngtgpu_index *index, uint32_t n_pivots, uint64_t rng_seed);
/* The same with pivots drawn from ids 1..limit only (0 = all): used while a graph is grown batch by batch
 * (ngtgpu_index_insert_batch), when later ids are not in the graph yet. */
int ngtgpu_index_build_seed_table_range(ngtgpu_index *index, uint32_t n_pivots, uint64_t rng_seed, uint64_t limit);
/* An explicit pivot list (host ids). ids 1..seedSize with n_seeds == seedSize in the search calls is the reference's
 * SeedTypeFixedNodes (lib/NGT/Index.h:1122-1127). */
int ngtgpu_index_set_seed_table_ids(ngtgpu_index *index, const uint32_t *pivot_ids, uint32_t n_pivots);

uint64_t ngtgpu_index_size(const ngtgpu_index *index);          /* n */
uint32_t ngtgpu_index_padded_dimension(const ngtgpu_index *index);
/* Copies the stored (possibly normalised) object `id` back: ngt_get_object_as_float/_as_integer
 * (lib/NGT/Capi.cpp:750-782). `out` holds `dimension` elements of the object type. */
int ngtgpu_index_get_object(const ngtgpu_index *index, uint32_t id, void *out);
/* Objects first .. first+count-1 into `out` (count x dimension elements, packed) with one strided copy. */
int ngtgpu_index_get_objects(const ngtgpu_index *index, uint32_t first, uint64_t count, void *out);

/* ---- graph beam search: replaces NeighborhoodGraph::search / searchReadOnlyGraph
 *      (lib/NGT/Graph.cpp:398-495, 499-638) behind GraphIndex::search(sc, seeds)
 *      (lib/NGT/Index.h:1140-1179), for a batch of queries -------------------------------------
 * queries: nq rows of `dimension` elements; query_type says whether they are float (as every
 *   ngt_search_index* entry point passes them, Capi.cpp:377-406) or already uint8. They are cast to
 *   the object type / normalised on the device exactly as Index::allocateObject does
 *   (ObjectRepository.h:222-258, ObjectSpaceRepository.h:560-594).
 * seeds: nq x n_seeds explicit seed ids (what GraphIndex::search(sc, seeds) takes), or NULL to use the
 *   device seed table (n_seeds = number of pivots to take per query).
 * ids/dists: nq x size, ascending (distance,id); counts: results per query.
 * stats (nullable): nq x 3 = {distance computations incl. seeds, adjacency entries examined,
 *   nodes expanded} -- the reference's distanceComputationCount / visitCount (Graph.cpp:592,604). */
int ngtgpu_search(ngtgpu_index *index, const void *queries, int query_type, uint32_t nq,
                  const ngtgpu_search_params *params, const uint32_t *seeds, uint32_t n_seeds,
                  uint32_t *ids, float *dists, uint32_t *counts, uint32_t *stats);
/* Same with every buffer already in HBM, enqueued on `stream` (a cudaStream_t); returns without
 * synchronising unless a query overflowed the on-chip working set and had to be re-run. */
int ngtgpu_search_device(ngtgpu_index *index, const void *queries, int query_type, uint32_t nq,
                         const ngtgpu_search_params *params, const uint32_t *seeds, uint32_t n_seeds,
                         uint32_t *ids, float *dists, uint32_t *counts, uint32_t *stats, void *stream);

/* ---- exhaustive scan: replaces ObjectSpaceRepository::linearSearch
 *      (lib/NGT/ObjectSpaceRepository.h:466-502) behind GraphIndex::linearSearch (Index.h:729-749) */
int ngtgpu_linear_search(ngtgpu_index *index, const void *queries, int query_type, uint32_t nq, uint32_t size,
                         float radius, uint32_t *ids, float *dists, uint32_t *counts);
int ngtgpu_linear_search_device(ngtgpu_index *index, const void *queries, int query_type, uint32_t nq,
                                uint32_t size, float radius, uint32_t *ids, float *dists, uint32_t *counts,
                                void *stream);

/* ---- exhaustive kNN of stored objects: the brute-force pass behind graph construction,
 *      GraphIndex::searchForKNNGInsertion (lib/NGT/Index.h:839-856: linearSearch with size k+1 for object
 *      `id`, the object itself dropped, ObjectSpace.h:70-88). Queries are the stored rows
 *      first_id..first_id+count-1; outputs are DEVICE buffers [count x k], enqueued on `stream`. */
int ngtgpu_index_knn_graph(ngtgpu_index *index, uint32_t k, uint32_t first_id, uint32_t count, uint32_t *ids,
                           float *dists, uint32_t *counts, void *stream);

/* The seeds the engine would start these (host) queries from -- the nearest `n_seeds` pivots of the seed
 * table -- so a caller can run the reference's GraphIndex::search(sc, seeds) (Index.h:1140) from the same
 * starting points. seeds_out: nq x n_seeds. */
int ngtgpu_select_seeds(ngtgpu_index *index, const void *queries, int query_type, uint32_t nq, uint32_t n_seeds,
                        uint32_t *seeds_out);

/* Working set of the traversal kernel per query: visited-hash slots = 2^hash_bits (8..17; an exact
 * open-addressing table in a per-CTA slab that stays in L2), unchecked-queue entries in shared memory
 * (64..8192). A query that outgrows it is re-run with a 2^17-slot slab and finally with an exact bitmap +
 * queue in HBM (the reference's own structures, lib/NGT/Graph.h:751-799); ngtgpu_index_last_overflows()
 * says how many queries of the last host-pointer call left the first tier. */
int ngtgpu_index_set_search_workspace(ngtgpu_index *index, uint32_t hash_bits, uint32_t queue_cap);
/* Number of shared-memory tiers tried before the HBM tier: 2 (default: the configured one, then the largest
 * that fits an SM) or 1. */
int ngtgpu_index_set_onchip_tiers(ngtgpu_index *index, int tiers);
/* The first tier of the common case (rows of 80..512 bytes, edge cap <= 128, epsilon >= 0, size <= 32) runs a leaner
 * kernel with the same results; 0 sends it through the general kernel as well (default 1). */
int ngtgpu_index_set_fast_kernel(ngtgpu_index *index, int enabled);
/* Shape of the lean kernel's on-chip tiers: warps per query -- 0 (default): 2 whenever a round fits 64 threads (<= 64
 * edges and seeds; 16 CTAs per SM, twice the queries in flight: rounds are latency-bound), else 4; 4 as asked; 2 as asked
 * also for rounds of up to 128 edges (two per thread; measured slower than 4 on 512-byte rows); 1 as asked for rows of
 * <= 128 bytes and rounds of <= 64 edges (32 CTAs per SM, first tier only; measured slower than 2) -- and a cap on
 * resident CTAs per SM (0: what fits). Same results either way. */
int ngtgpu_index_set_fast_shape(ngtgpu_index *index, int warps_per_query, int ctas_per_sm);
/* Searches without a seed list take the nearest pivots of the seed table. 1: the lean traversal kernel selects them itself
 * (no separate selection launch: +1.4 % queries/s on the 1M x 128 set, but the traversal launch then also carries the
 * table scan); 0 (default): a selection kernel runs first. Same seeds, same results either way. */
int ngtgpu_index_set_seed_fusion(ngtgpu_index *index, int enabled);
/* Shared-memory staging area (bytes per CTA) that neighbour rows are copied into with cp.async. */
int ngtgpu_index_set_stage_bytes(ngtgpu_index *index, uint32_t bytes);
uint64_t ngtgpu_index_last_overflows(const ngtgpu_index *index);

/* Exhaustive float batches (>= 1024 queries against >= 32768 objects, size <= 100, dimension <= 384) are filtered
 * on the tensor cores (tcgen05, bf16-split operands) and re-evaluated exactly; results are identical to the CUDA-core
 * scan. This switch turns the tensor-core filter off (parity tests); the counter says how many batches used it. */
int ngtgpu_index_set_tensor_core(ngtgpu_index *index, int enabled);
uint64_t ngtgpu_index_tensor_core_batches(const ngtgpu_index *index);

/* ---- multi-GPU: per-shard result lists <-> 64-bit keys (ordered distance bits << 32 | global id), and the k-way
 *      merge of all-gathered key lists by (distance, id) (lib/NGT/Common.h:1946-1952). Device buffers. ---------- */
int ngtgpu_pack_keys(const uint32_t *ids, const float *dists, const uint32_t *counts, uint32_t nq, uint32_t k,
                     uint32_t id_offset, uint64_t *keys, void *stream);
/* keys: [n_lists][nq][k] (n_lists <= 32), each list ascending, padded with 0xffffffffffffffff. */
int ngtgpu_merge_keys(const uint64_t *keys, uint32_t n_lists, uint32_t nq, uint32_t k, uint32_t *ids, float *dists,
                      uint32_t *counts, void *stream);

/* ---- multi-GPU: the objects sharded by rows over the GPUs of one box (the reference has no
 *      distribution). Shard g is an ordinary ngtgpu_index over its block of the objects (own graph, own seed table),
 *      global id = local id + id_offset. Every shard answers the whole query batch; the traversal kernel writes its
 *      results as 64-bit keys straight into the shard's slot of the gather buffer, ONE ncclAllGather (in place, NVLink)
 *      exchanges them and a device merge keeps the k smallest (distance, id) per query -- what one priority queue over
 *      the union keeps (lib/NGT/Common.h:1946-1952), so merged == search of the union, bit for bit.
 *      NCCL is bound at run time (libnccl.so.2; NGTGPU_NCCL_SO overrides), only when these entry points are used. ----
 *
 *  (1) one process per GPU (torchrun / MPI): the launcher carries the 128-byte id from rank 0 to the other ranks. */
#define NGTGPU_COMM_ID_BYTES 128
typedef struct ngtgpu_comm ngtgpu_comm;
int ngtgpu_comm_get_unique_id(void *id_out /* NGTGPU_COMM_ID_BYTES */);
int ngtgpu_comm_create(ngtgpu_comm **out, const void *id, int rank, int world, int device);
int ngtgpu_comm_destroy(ngtgpu_comm *comm);
/* This rank's shard `index` searched for the batch (DEVICE queries, identical on every rank), all-gather, merge: every
 * rank receives the merged lists (DEVICE buffers, global ids). Enqueued on `stream`. */
int ngtgpu_shard_search_device(ngtgpu_index *index, ngtgpu_comm *comm, const void *queries, int query_type, uint32_t nq,
                               const ngtgpu_search_params *params, uint32_t n_seeds, uint32_t id_offset, uint32_t *ids,
                               float *dists, uint32_t *counts, void *stream);
int ngtgpu_shard_linear_search_device(ngtgpu_index *index, ngtgpu_comm *comm, const void *queries, int query_type, uint32_t nq,
                                      uint32_t size, float radius, uint32_t id_offset, uint32_t *ids, float *dists,
                                      uint32_t *counts, void *stream);
/* CUDA-event timing of the calls above: ms3 = {this shard's search, all-gather (waits for the slowest shard), merge}. */
int ngtgpu_comm_set_timing(ngtgpu_comm *comm, int enabled);
int ngtgpu_comm_pop_timing(ngtgpu_comm *comm, double *ms3, uint64_t *calls);

/*  (2) one process driving several GPUs -- what a program written against lib/NGT/Capi.h is (ngt_open_index uses this
 *      handle when NGTGPU_DEVICES lists several devices). HOST buffers in and out, like ngtgpu_search. */
typedef struct ngtgpu_sharded ngtgpu_sharded;
int ngtgpu_sharded_create(ngtgpu_sharded **out, const int *devices, int n_devices, int object_type, int distance_type,
                          uint32_t dimension);
int ngtgpu_sharded_destroy(ngtgpu_sharded *sharded);
/* objects 1..n (host rows): rows [g*n/G, (g+1)*n/G) become shard g */
int ngtgpu_sharded_set_objects(ngtgpu_sharded *sharded, const void *objects, uint64_t n, int normalize);
/* every shard builds its own ONNG (ngtgpu_index_build_onng) and seed table, all devices at once */
int ngtgpu_sharded_build_onng(ngtgpu_sharded *sharded, uint32_t knn, uint32_t outgoing, uint32_t incoming,
                              int shortcut_reduction, int64_t edge_size_for_search, uint32_t n_pivots);
int ngtgpu_sharded_shard_count(const ngtgpu_sharded *sharded);
/* shard g's own index handle (to set an explicit graph, tune its workspace, ...), its id offset and size */
int ngtgpu_sharded_shard(ngtgpu_sharded *sharded, int shard, ngtgpu_index **index, uint64_t *id_offset, uint64_t *count);
/* upload to the first device, ncclBroadcast, per-shard search, ncclAllGather, merge, download */
int ngtgpu_sharded_search(ngtgpu_sharded *sharded, const void *queries, int query_type, uint32_t nq,
                          const ngtgpu_search_params *params, uint32_t n_seeds, uint32_t *ids, float *dists, uint32_t *counts);
int ngtgpu_sharded_linear_search(ngtgpu_sharded *sharded, const void *queries, int query_type, uint32_t nq, uint32_t size,
                                 float radius, uint32_t *ids, float *dists, uint32_t *counts);
/* ms5 of the last call, on the first device's stream: upload + broadcast, search, all-gather, merge, download */
int ngtgpu_sharded_last_timing(const ngtgpu_sharded *sharded, double *ms5);

/* ---- graph from the exhaustive kNN pass: a DEVICE [n x k] neighbour table (ngtgpu_index_knn_graph) -> DEVICE CSR over
 *      ids 0..n with distances, lists ascending by (distance, id). symmetric != 0 adds the reverse of every edge, i.e.
 *      the graph insertANNGNode's out-edges + reverse edges converge to (lib/NGT/Graph.h:611-626). valid (nullable):
 *      one byte per id, 0 = removed slot. capacity: entries of out_col / out_dist (n * k * 2 always suffices). */
int ngtgpu_graph_from_knn_table(uint64_t n, const uint32_t *ids, const float *dists, const uint32_t *counts, uint32_t k,
                                const uint8_t *valid, int symmetric, uint64_t capacity, uint64_t *out_row_ptr,
                                uint32_t *out_col, float *out_dist, uint64_t *out_nnz, void *stream);

/* ---- ONNG construction, first step: GraphReconstructor::reconstructGraph (lib/NGT/GraphReconstructor.h:425-561):
 *      every node keeps its first `outgoing` edges (all, when it has fewer) and receives the reverse of the first
 *      `incoming` edges of every node; lists sorted by (distance, id), repeated ids dropped. DEVICE CSR in (ids 0..n),
 *      DEVICE CSR out: out_row_ptr n+2 entries, out_col / out_dist `capacity` entries (2 x the input edges always
 *      suffice); *out_nnz (host) receives the number of edges written. */
int ngtgpu_graph_reconstruct(uint64_t n, const uint64_t *row_ptr, const uint32_t *col, const float *dist,
                             uint32_t outgoing, uint32_t incoming, uint64_t capacity, uint64_t *out_row_ptr,
                             uint32_t *out_col, float *out_dist, uint64_t *out_nnz, void *stream);

/* ---- ONNG construction, second step: GraphReconstructor::adjustPathsEffectively
 *      (lib/NGT/GraphReconstructor.h:197-386; run by GraphOptimizer::execute after reconstructGraph,
 *      lib/NGT/GraphOptimizer.h:279-292): drop edge src->dst when src->path->dst with both hops shorter is already
 *      in the rebuilt graph, unless the node would be left with min_edges edges or fewer. The graph is a DEVICE CSR
 *      over ids 0..n (row_ptr: n+2 entries, id 0 has no edges), lists ascending by (distance, id) as `grp` stores
 *      them; keep[e] (device, one byte per edge) becomes 1 for the edges of the adjusted graph -- the reference's
 *      result exactly. stats (host, nullable): candidates, removed edges, sweep launches, kernels launched. */
int ngtgpu_graph_adjust_paths(uint64_t n, const uint64_t *row_ptr, const uint32_t *col, const float *dist,
                              uint32_t min_edges, uint8_t *keep, uint64_t *stats, void *stream);

/* ---- one batch of the reference's ANNG construction loop / incremental insertion (lib/NGT/Index.cpp:631-719:
 *      searchMultipleQueryForCreation + insertMultipleSearchResults; Index.h:815-837 searchForNNGInsertion;
 *      Graph.h:611-626 insertANNGNode): the stored objects first_id .. first_id+count-1 are searched for in the graph
 *      as it is (size = edge_size_for_creation, epsilon = the creation epsilon, edge_size as in ngtgpu_search_params;
 *      seeds = nearest n_seeds of n_pivots pivots drawn from ids < first_id; n_pivots == 0: the reference's
 *      SeedTypeFixedNodes, ids 1..min(n_seeds, first_id - 1), Index.h:1122-1127), each also gets the distances to the
 *      objects before it in the batch, its list is cut to edge_size_for_creation, becomes its edges, and every listed
 *      node gets the reverse edge. DEVICE CSR with distances, updated in place (capacity entries; grows by at most
 *      2 * count * edge_size_for_creation). An empty graph (first batch) only links the batch internally. The index's
 *      own graph is left equal to the result. */
int ngtgpu_index_insert_batch(ngtgpu_index *index, uint32_t first_id, uint32_t count, uint32_t edge_size_for_creation,
                              float epsilon, int64_t edge_size, uint32_t n_seeds, uint32_t n_pivots, uint64_t pivot_seed,
                              uint64_t capacity, uint64_t *d_row_ptr, uint32_t *d_col, float *d_dist, uint64_t *nnz_out);

/* ---- the reference's ONNG recipe for the objects of one index, all on the device: exact kNN table (the brute-force pass
 *      of lib/NGT/Index.h:839-856) -> kNN graph -> reconstructGraph(outgoing, incoming) (GraphReconstructor.h:425-561) ->
 *      adjustPathsEffectively when shortcut_reduction != 0 (:197-386, min_edges as there) -> the index's graph. graph_out
 *      (nullable) receives device copies of the CSR with distances, to be released with ngtgpu_device_free; seconds
 *      (nullable, 3 entries): kNN pass, reconstructGraph, path adjustment, from CUDA events. */
typedef struct {
  uint64_t n, nnz;
  uint64_t *row_ptr; /* n + 2 */
  uint32_t *col;
  float *dist;
} ngtgpu_graph_buffers;
int ngtgpu_index_build_onng(ngtgpu_index *index, uint32_t knn, uint32_t outgoing, uint32_t incoming, int shortcut_reduction,
                            uint32_t min_edges, ngtgpu_graph_buffers *graph_out, double *seconds);
int ngtgpu_device_free(void *device_pointer);
int ngtgpu_device_copy(void *dst, const void *src, uint64_t bytes); /* device to device, synchronous */
/* Counters of the construction path since the library was loaded (evidence for tools/anng_probe.py):
 * [0] batches merged into the sorted lists, [1] batches that took the full sort, [2] temporary blocks that had to come
 * from cudaMalloc, [3] temporary blocks served by the per-device cache. */
int ngtgpu_construction_counters(uint64_t out[4]);

/* The sub-graph of the edges with keep[e] != 0, order inside the lists preserved (compaction after
 * ngtgpu_graph_adjust_paths). DEVICE buffers; *out_nnz is a host word. */
int ngtgpu_graph_select_edges(uint64_t n, const uint64_t *row_ptr, const uint32_t *col, const float *dist,
                              const uint8_t *keep, uint64_t *out_row_ptr, uint32_t *out_col, float *out_dist,
                              uint64_t *out_nnz, void *stream);

/* ---- GraphReconstructor::refineANNG (lib/NGT/GraphReconstructor.h:814-924; C API ngt_refine_anng): in batches of
 *      batch_size, every stored object is searched for in the current graph (size searched_edges, epsilon, edge_size
 *      as in ngtgpu_search_params, seeds = nearest n_seeds pivots), its results are merged into its edge list and --
 *      when no_of_edges == 0 -- the reverse edges are added; later batches search the refined graph; no_of_edges > 0
 *      finally cuts every list to that length. d_row_ptr / d_col / d_dist: DEVICE CSR with distances, updated in place
 *      (col / dist hold `capacity` entries; nnz + 2 * n * searched_edges always suffices). The index's own graph is left
 *      equal to the result. */
int ngtgpu_index_refine_anng(ngtgpu_index *index, float epsilon, int32_t no_of_edges, int64_t edge_size,
                             uint32_t searched_edges, uint64_t batch_size, uint32_t n_seeds, uint64_t capacity,
                             uint64_t *d_row_ptr, uint32_t *d_col, float *d_dist, uint64_t *nnz_out);

/* Index::AccuracyTable::getEpsilon (lib/NGT/Index.h:293-360): the epsilon GraphIndex::search substitutes when a query
 * carries an expected accuracy (Index.h:1156-1158). table: the `AccuracyTable` line of `prf` ("epsilon:accuracy,...").
 * Host arithmetic; fails with the reference's messages for a table of two or fewer points or a malformed token. */
int ngtgpu_epsilon_from_accuracy_table(const char *table, double accuracy, float *epsilon);

/* The m x m distances among the stored objects `ids` (host ids in, host floats out; m <= 4096), with the engine's exact
 * distance: the comparator calls of NeighborhoodGraph::removeEdgesReliably (lib/NGT/Graph.cpp:641-864), which re-links
 * the neighbours of a node that NGT::Index::remove takes out. */
int ngtgpu_index_pairwise_distances(ngtgpu_index *index, const uint32_t *ids, uint32_t m, float *out);

/* Number of kernels this library launched since the index was created (bench.py's gpu_launches). */
uint64_t ngtgpu_index_launch_count(const ngtgpu_index *index);
/* Device timing of the traversal kernel with CUDA events on the launching stream (bench.py's roofline leg):
 * enable, run searches, then pop the summed milliseconds and the number of traversal launches. */
int ngtgpu_index_set_timing(ngtgpu_index *index, int enabled);
int ngtgpu_index_pop_timing(ngtgpu_index *index, double *total_ms, uint64_t *launches);

/* Development aid: when a device buffer of nq x 8 words is set, the traversal kernel's control warp writes the
 * cycles it spent per phase (merge, pop, adjacency load, visited filter, TMA issue, row wait, distances, other). */
int ngtgpu_index_set_phase_profile(ngtgpu_index *index, uint32_t *device_buffer);

/* ---- NGT index files (host only): `obj` and `grp` as Repository::serialize writes them
 *      (lib/NGT/Common.h:1776-1837, ObjectSpace.h:293-301, Graph.h:151-158). record_bytes =
 *      dimension * sizeof(object element). Slots = repository size = n + 1 (slot 0 is the dummy). */
int ngtgpu_io_obj_info(const char *path, uint32_t record_bytes, uint64_t *slots, uint64_t *present);
int ngtgpu_io_read_obj(const char *path, uint32_t record_bytes, void *rows, uint8_t *present);
int ngtgpu_io_write_obj(const char *path, uint32_t record_bytes, const void *rows, uint64_t n, const uint8_t *present);
int ngtgpu_io_grp_info(const char *path, uint64_t *slots, uint64_t *nnz);
int ngtgpu_io_read_grp(const char *path, uint64_t *row_ptr, uint32_t *col, float *dist, uint8_t *present);
int ngtgpu_io_write_grp(const char *path, uint64_t n, const uint64_t *row_ptr, const uint32_t *col, const float *dist,
                        const uint8_t *present);

#ifdef __cplusplus
}
#endif
#endif /* NGTGPU_H */

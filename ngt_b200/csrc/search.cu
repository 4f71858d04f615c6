// search.cu -- host side of the batched graph search (the C-ABI entry points ngtgpu_search*).
//
// Stands where GraphIndex::search(sc, seeds) (lib/NGT/Index.h:1140-1179) and GraphAndTreeIndex::search
// (Index.h:1570-1577) stand in the reference: resolve the edge-size mode (Graph.h:675-692), get seeds
// (explicit, or nearest pivots of the device seed table instead of the DVP-tree leaf, Index.h:1524-1567),
// run the traversal kernel, hand back ascending (distance,id) lists.
#include <cfloat>
#include <cmath>

#include "search.cuh"

template <int ACC>
cudaError_t search_dispatch(const SearchArgs &a, const SearchLaunch &l, int op, int *blocks);
extern template cudaError_t search_dispatch<ACC_F_L2>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_F_DOT>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_F_COS>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_U8_L2>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_U8_HAM>(const SearchArgs &, const SearchLaunch &, int, int *);

static cudaError_t dispatch(int acc, const SearchArgs &a, const SearchLaunch &l, int op, int *blocks) {
  switch (acc) {
    case ACC_F_L2: return search_dispatch<ACC_F_L2>(a, l, op, blocks);
    case ACC_F_DOT: return search_dispatch<ACC_F_DOT>(a, l, op, blocks);
    case ACC_F_COS: return search_dispatch<ACC_F_COS>(a, l, op, blocks);
    case ACC_U8_L2: return search_dispatch<ACC_U8_L2>(a, l, op, blocks);
    case ACC_U8_HAM: return search_dispatch<ACC_U8_HAM>(a, l, op, blocks);
  }
  return cudaErrorInvalidValue;
}

#define BIG_TIER_CTAS 32u
#define BIG_TIER_QUEUE (1u << 20)

int ngtgpu_traverse(ngtgpu_index *ix, const uint8_t *d_queries, uint32_t nq, const ngtgpu_search_params *params,
                    const uint32_t *d_seeds, uint32_t n_seeds, uint32_t *d_ids, float *d_dists, uint32_t *d_counts,
                    uint32_t *d_stats, cudaStream_t stream) {
  if (!ix->d_row_ptr) NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: the graph is not set");
  int64_t cap = ngtgpu_effective_edge_size(ix, params);
  if (cap < 0) {
    // message of NeighborhoodGraph::getEdgeSize, Graph.h:687-689
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "NGT::getEdgeSize: Invalid edge size parameters " +
                                        std::to_string(params->edge_size) + ":" +
                                        std::to_string(ix->edge_size_for_search));
  }
  const uint32_t k = params->size;
  if (k > 2048) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: size > 2048 is not supported by the on-chip result list");

  SearchArgs a;
  memset(&a, 0, sizeof(a));
  a.objects = ix->d_objects;
  a.row_bytes = ix->row_bytes;
  a.chunks = ix->chunks;
  a.n = ix->n;
  a.row_ptr = ix->d_row_ptr;
  a.col = ix->d_col;
  a.queries = d_queries;
  a.seeds = d_seeds;
  a.n_seeds = n_seeds;
  a.nq = nq;
  a.k = k;
  a.coef = (float)((double)params->epsilon + 1.0);  // Common.h:2041
  a.radius = params->radius < 0.0f ? FLT_MAX : params->radius;  // Capi.cpp:384-386
  a.edge_cap = (uint32_t)cap;
  a.dtype = ix->distance_type;
  a.hash_bits = ix->hash_bits;
  a.hash_limit = (uint32_t)((3ull << ix->hash_bits) / 4);
  a.queue_cap = ix->queue_cap;
  a.ids = d_ids;
  a.dists = d_dists;
  a.counts = d_counts;
  a.stats = d_stats;

  // counters: [0] work counter tier 0, [1] overflow count, [2] work counter tier 1, [3] failed count, [4..] list
  uint32_t *ws = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEARCH_WS, (size_t)(nq + 8) * sizeof(uint32_t), (void **)&ws));
  CUDA_TRY(cudaMemsetAsync(ws, 0, 8 * sizeof(uint32_t), stream));
  a.work_counter = ws + 0;
  a.overflow_count = ws + 1;
  a.overflow_list = ws + 8;
  a.failed_count = ws + 3;

  // ---- tier 0: shared-memory working set
  SearchLaunch l;
  l.group = (int)ix->group;
  if (ix->group < 32) l.cpl = 1;
  else {
    uint32_t per_lane = (ix->chunks + 31) / 32;
    l.cpl = per_lane <= 1 ? 1 : per_lane <= 2 ? 2 : per_lane <= 4 ? 4 : per_lane <= 8 ? 8 : 0;
  }
  l.ws = 0;
  l.stream = stream;
  size_t smem = (k > 32 ? (((size_t)k * 8 + 15) & ~(size_t)15) : 0) + (size_t)a.queue_cap * 8 + ((size_t)4 << a.hash_bits);
  if (l.cpl == 0) smem += ix->row_bytes;
  if (smem > 200 * 1024) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: working set does not fit shared memory; lower hash_bits/queue_cap");
  l.smem = smem;
  int blocks = 0;
  cudaError_t e = dispatch(ix->acc_kind, a, l, 1, &blocks);
  if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("search occupancy query: ") + cudaGetErrorString(e));
  if (blocks < 1) blocks = 1;
  uint64_t grid = (uint64_t)blocks * ix->sm_count;
  if (grid > nq) grid = nq;
  if (grid == 0) return NGTGPU_OK;
  l.grid = (unsigned)grid;
  e = dispatch(ix->acc_kind, a, l, 0, nullptr);
  if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("search kernel launch: ") + cudaGetErrorString(e));
  ix->launches++;

  // ---- tier 1: HBM working set for the queries that overflowed (exits at once when there are none)
  const uint64_t bitmap_words = ((ix->n + 1 + 31) / 32 + 3) & ~(uint64_t)3;
  uint32_t big_queue = BIG_TIER_QUEUE;
  size_t big_bytes = (size_t)BIG_TIER_CTAS * (bitmap_words * 4 + (size_t)big_queue * 8);
  uint8_t *big = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEARCH_BIG, big_bytes, (void **)&big));
  SearchArgs b = a;
  b.work_counter = ws + 2;
  b.query_list = ws + 8;
  b.query_list_count = ws + 1;
  b.queue_cap = big_queue;
  b.big_queues = reinterpret_cast<uint64_t *>(big);
  b.big_bitmaps = reinterpret_cast<uint32_t *>(big + (size_t)BIG_TIER_CTAS * big_queue * 8);
  b.bitmap_words = bitmap_words;
  SearchLaunch lb = l;
  lb.group = 32;
  lb.cpl = 0;
  lb.ws = 1;
  lb.grid = BIG_TIER_CTAS;
  lb.smem = (k > 32 ? (((size_t)k * 8 + 15) & ~(size_t)15) : 0) + ix->row_bytes;
  e = dispatch(ix->acc_kind, b, lb, 0, nullptr);
  if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("search overflow-tier launch: ") + cudaGetErrorString(e));
  ix->launches++;
  return NGTGPU_OK;
}

// seeds for a prepared query batch: nearest `n_seeds` pivots of the seed table
static int select_seeds(ngtgpu_index *ix, const uint8_t *d_queries, uint32_t nq, uint32_t n_seeds, uint32_t **d_seeds,
                        cudaStream_t stream) {
  if (ix->n_pivots == 0)
    NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: no seeds given and no seed table built (ngtgpu_index_build_seed_table)");
  if (n_seeds == 0) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: n_seeds is zero");
  if (n_seeds > ix->n_pivots) n_seeds = ix->n_pivots;
  uint32_t *seeds = nullptr;
  float *sd = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEEDS, (size_t)nq * n_seeds * sizeof(uint32_t), (void **)&seeds));
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEED_DISTS, (size_t)nq * (n_seeds + 1) * sizeof(float) + 256, (void **)&sd));
  ScanParams p;
  p.d_queries = d_queries;
  p.nq = nq;
  p.d_rows = ix->d_pivot_rows;
  p.n_rows = ix->n_pivots;
  p.d_id_map = ix->d_pivot_ids;
  p.k = n_seeds;
  p.radius = -1.0f;
  p.d_ids = seeds;
  p.d_dists = sd + 64;
  p.d_counts = reinterpret_cast<uint32_t *>(sd + 64 + (size_t)nq * n_seeds);
  NGTGPU_TRY(ngtgpu_scan_topk(ix, p, stream));
  *d_seeds = seeds;
  return NGTGPU_OK;
}

static int search_common(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                         const ngtgpu_search_params *params, const uint32_t *seeds, uint32_t n_seeds, uint32_t *ids,
                         float *dists, uint32_t *counts, uint32_t *stats, bool on_device, cudaStream_t stream) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!params) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: null params");
  if (!ix->d_objects || ix->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: the index holds no objects");
  if (nq == 0) return NGTGPU_OK;
  if (!queries || !counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: null buffer");
  const uint32_t k = params->size;
  if (k == 0) {  // Index.h:1141-1144
    if (on_device) CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)nq * 4, stream));
    else memset(counts, 0, (size_t)nq * 4);
    return NGTGPU_OK;
  }
  if (!ids || !dists) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: null result buffer");
  uint8_t *d_q = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_QUERIES, (size_t)nq * ix->row_bytes, (void **)&d_q));
  NGTGPU_TRY(ngtgpu_prepare_queries(ix, queries, query_type, nq, on_device, d_q, stream));
  const uint32_t *d_seeds = seeds;
  uint32_t ns = n_seeds;
  if (seeds) {
    if (!on_device) {
      uint32_t *s = nullptr;
      NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEEDS, (size_t)nq * n_seeds * sizeof(uint32_t), (void **)&s));
      CUDA_TRY(cudaMemcpyAsync(s, seeds, (size_t)nq * n_seeds * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
      d_seeds = s;
    }
  } else {
    uint32_t *s = nullptr;
    if (ns > ix->n_pivots) ns = ix->n_pivots;
    NGTGPU_TRY(select_seeds(ix, d_q, nq, ns, &s, stream));
    d_seeds = s;
  }
  uint32_t *d_ids = ids, *d_counts = counts, *d_stats = stats;
  float *d_dists = dists;
  if (!on_device) {
    size_t words = (size_t)nq * k * 2 + nq + (stats ? (size_t)nq * 3 : 0);
    uint32_t *io = nullptr;
    NGTGPU_TRY(ngtgpu_scratch(ix, SCR_IO, words * 4, (void **)&io));
    d_ids = io;
    d_dists = reinterpret_cast<float *>(io + (size_t)nq * k);
    d_counts = io + (size_t)nq * k * 2;
    d_stats = stats ? d_counts + nq : nullptr;
  }
  NGTGPU_TRY(ngtgpu_traverse(ix, d_q, nq, params, d_seeds, ns, d_ids, d_dists, d_counts, d_stats, stream));
  if (!on_device) {
    uint32_t *ws = (uint32_t *)ix->d_scratch[SCR_SEARCH_WS];
    uint32_t h_ws[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(ids, d_ids, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(dists, d_dists, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(counts, d_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    if (stats) CUDA_TRY(cudaMemcpyAsync(stats, d_stats, (size_t)nq * 12, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_ws, ws, sizeof(h_ws), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    ix->last_overflows = h_ws[1];
    if (h_ws[3])
      NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: " + std::to_string(h_ws[3]) +
                                        " queries outgrew the HBM working set (unchecked queue > 2^20 entries)");
  }
  return NGTGPU_OK;
}

extern "C" int ngtgpu_search(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                             const ngtgpu_search_params *params, const uint32_t *seeds, uint32_t n_seeds, uint32_t *ids,
                             float *dists, uint32_t *counts, uint32_t *stats) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  return search_common(ix, queries, query_type, nq, params, seeds, n_seeds, ids, dists, counts, stats, false, ix->stream);
}

extern "C" int ngtgpu_search_device(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                                    const ngtgpu_search_params *params, const uint32_t *seeds, uint32_t n_seeds,
                                    uint32_t *ids, float *dists, uint32_t *counts, uint32_t *stats, void *stream) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  return search_common(ix, queries, query_type, nq, params, seeds, n_seeds, ids, dists, counts, stats, true,
                       (cudaStream_t)stream);
}

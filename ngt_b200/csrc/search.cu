// search.cu -- host side of the batched graph search (the C-ABI entry points ngtgpu_search*).
//
// Stands where GraphIndex::search(sc, seeds) (lib/NGT/Index.h:1140-1179) and GraphAndTreeIndex::search
// (Index.h:1570-1577) stand in the reference: resolve the edge-size mode (Graph.h:675-692), get seeds
// (explicit, or nearest pivots of the device seed table instead of the DVP-tree leaf, Index.h:1524-1567),
// run the traversal kernel, hand back ascending (distance,id) lists.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "search_fast.cuh"

template <int ACC>
cudaError_t search_dispatch(const SearchArgs &a, const SearchLaunch &l, int op, int *blocks);
extern template cudaError_t search_dispatch<ACC_F_L2>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_F_DOT>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_F_COS>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_U8_L2>(const SearchArgs &, const SearchLaunch &, int, int *);
extern template cudaError_t search_dispatch<ACC_U8_HAM>(const SearchArgs &, const SearchLaunch &, int, int *);

// the lean kernel of the common case (search_fast.cuh)
extern template cudaError_t search_fast_dispatch<ACC_F_L2>(const SearchArgs &, int, int, unsigned, size_t, cudaStream_t, int, int *);
extern template cudaError_t search_fast_dispatch<ACC_F_DOT>(const SearchArgs &, int, int, unsigned, size_t, cudaStream_t, int, int *);
extern template cudaError_t search_fast_dispatch<ACC_F_COS>(const SearchArgs &, int, int, unsigned, size_t, cudaStream_t, int, int *);
extern template cudaError_t search_fast_dispatch<ACC_U8_L2>(const SearchArgs &, int, int, unsigned, size_t, cudaStream_t, int, int *);
extern template cudaError_t search_fast_dispatch<ACC_U8_HAM>(const SearchArgs &, int, int, unsigned, size_t, cudaStream_t, int, int *);

extern template cudaError_t seed_select_dispatch<ACC_F_L2>(const SeedArgs &, cudaStream_t);
extern template cudaError_t seed_select_dispatch<ACC_F_DOT>(const SeedArgs &, cudaStream_t);
extern template cudaError_t seed_select_dispatch<ACC_F_COS>(const SeedArgs &, cudaStream_t);
extern template cudaError_t seed_select_dispatch<ACC_U8_L2>(const SeedArgs &, cudaStream_t);
extern template cudaError_t seed_select_dispatch<ACC_U8_HAM>(const SeedArgs &, cudaStream_t);

static cudaError_t dispatch_seeds(int acc, const SeedArgs &a, cudaStream_t stream) {
  switch (acc) {
    case ACC_F_L2: return seed_select_dispatch<ACC_F_L2>(a, stream);
    case ACC_F_DOT: return seed_select_dispatch<ACC_F_DOT>(a, stream);
    case ACC_F_COS: return seed_select_dispatch<ACC_F_COS>(a, stream);
    case ACC_U8_L2: return seed_select_dispatch<ACC_U8_L2>(a, stream);
    case ACC_U8_HAM: return seed_select_dispatch<ACC_U8_HAM>(a, stream);
  }
  return cudaErrorInvalidValue;
}

static cudaError_t dispatch_fast(int acc, const SearchArgs &a, int ch, int warps, unsigned grid, size_t smem, cudaStream_t stream,
                                 int op, int *blocks) {
  switch (acc) {
    case ACC_F_L2: return search_fast_dispatch<ACC_F_L2>(a, ch, warps, grid, smem, stream, op, blocks);
    case ACC_F_DOT: return search_fast_dispatch<ACC_F_DOT>(a, ch, warps, grid, smem, stream, op, blocks);
    case ACC_F_COS: return search_fast_dispatch<ACC_F_COS>(a, ch, warps, grid, smem, stream, op, blocks);
    case ACC_U8_L2: return search_fast_dispatch<ACC_U8_L2>(a, ch, warps, grid, smem, stream, op, blocks);
    case ACC_U8_HAM: return search_fast_dispatch<ACC_U8_HAM>(a, ch, warps, grid, smem, stream, op, blocks);
  }
  return cudaErrorInvalidValue;
}

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

static int select_seeds(ngtgpu_index *ix, const uint8_t *d_queries, uint32_t nq, uint32_t n_seeds, uint32_t **d_seeds,
                        cudaStream_t stream);

#define BIG_TIER_QUEUE (1u << 18)
#ifndef FAST_MIN_CHUNKS
#define FAST_MIN_CHUNKS 5   // shorter rows stay on the general kernel (many rows per copy instruction)
#endif

int ngtgpu_traverse(ngtgpu_index *ix, const uint8_t *d_queries, uint32_t nq, const ngtgpu_search_params *params,
                    const uint32_t *d_seeds, uint32_t n_seeds, uint32_t *d_ids, float *d_dists, uint32_t *d_counts,
                    uint32_t *d_stats, cudaStream_t stream, uint64_t *d_keys, uint32_t id_offset) {
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
  a.head = ix->d_head;
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
  a.keys_out = d_keys;
  a.id_offset = id_offset;
  a.prof = ix->d_prof;
  // No seed list: the nearest pivots of the seed table, from a selection pass that runs first -- or, when seed fusion is
  // switched on and the table is small, selected by the lean kernel itself in the first tier (and written out for the
  // later tiers).
  uint32_t *d_sel = nullptr;
  bool fused_seeds = false;
  if (!d_seeds) {
    if (ix->n_pivots == 0)
      NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: no seeds given and no seed table built (ngtgpu_index_build_seed_table)");
    if (n_seeds == 0) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: n_seeds is zero");
    if (n_seeds > ix->n_pivots) n_seeds = ix->n_pivots;
    a.n_seeds = n_seeds;
    const bool lean = ix->fast_kernel && ix->chunks >= FAST_MIN_CHUNKS && ix->chunks <= 32 && cap <= 32 * FAST_WARPS &&
                      a.coef >= 1.0f && k <= 32 && (ix->n + 1) * (uint64_t)ix->row_bytes < (1ull << 36);
    if (lean && ix->fuse_seeds && n_seeds <= 32 && ix->n_pivots <= 512) {
      NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEEDS, (size_t)nq * n_seeds * sizeof(uint32_t), (void **)&d_sel));
      fused_seeds = true;
      a.seeds = nullptr;
      a.pivots = ix->d_pivot_rows;
      a.pivot_ids = ix->d_pivot_ids;
      a.n_pivots = ix->n_pivots;
      a.seeds_out = d_sel;
    } else {
      NGTGPU_TRY(select_seeds(ix, d_queries, nq, n_seeds, &d_sel, stream));
      a.seeds = d_sel;
    }
  }

  // counters (16 words): [2t] work counter of tier t, [2t+1] overflow count of tier t (t = 0, 1), [4] work counter
  // of the HBM tier, [5] failed count; then two overflow lists of nq entries each
  uint32_t *ws = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEARCH_WS, ((size_t)2 * nq + 16) * sizeof(uint32_t), (void **)&ws));
  CUDA_TRY(cudaMemsetAsync(ws, 0, 16 * sizeof(uint32_t), stream));
  uint32_t *lists[2] = {ws + 16, ws + 16 + nq};
  a.failed_count = ws + 5;

  // ---- on-chip tiers: the configured working set, then (for the queries that outgrew it) the largest one
  // that still fits one CTA per SM. Later tiers read the previous tier's overflow list and exit at once
  // when it is empty, so no host synchronisation is needed in between.
  SearchLaunch l;
  l.group = (int)ix->group;
  if (ix->group < 32) l.cpl = 1;
  else {
    uint32_t per_lane = (ix->chunks + 31) / 32;
    l.cpl = per_lane <= 1 ? 1 : per_lane <= 2 ? 2 : per_lane <= 4 ? 4 : per_lane <= 8 ? 8 : 0;
  }
  l.ws = 0;
  l.stream = stream;
  const size_t res_bytes = k > 32 ? (((size_t)k * 8 + 15) & ~(size_t)15) : 0;
  // staging area of the TMA row gather: as many rows as fit the budget, 4..64
  // (split evenly between the 4 warps: a multiple of 4, and of 4 x rows-per-instruction for short rows)
  // (rows of the register-query kernels are staged at a stride of 512 * CPL bytes)
  const uint32_t stage_row = (l.group == 32 && l.cpl > 0) ? 512u * (uint32_t)l.cpl : ix->row_bytes;
  uint32_t stage_rows = ix->stage_bytes / stage_row;
  // (the fold8 kernels, CPL 1-2, read whole blocks of 8 slots per warp: a multiple of 32 rows)
  const uint32_t unit = (l.group == 32 && l.cpl > 0 && l.cpl <= 2) ? 8 * SEARCH_WARPS : SEARCH_WARPS * (32 / ix->group);
  if (stage_rows > 128) stage_rows = 128;
  if (stage_rows > 32u * SEARCH_WARPS) stage_rows = 32u * SEARCH_WARPS;   // a warp's slice holds at most 32 rows (one id per lane)
  stage_rows = stage_rows / unit * unit;
  if (stage_rows < unit) stage_rows = unit;
  a.stage_rows = stage_rows;
  const size_t stage_bytes = ((size_t)stage_rows * stage_row + 127) & ~(size_t)127;
  const size_t extra = stage_bytes + res_bytes + (l.cpl == 0 ? ix->row_bytes : 0);
  // first-tier slab: the configured size, or by index size (measured on the 12.5M x 128 uint8 shard at recall 0.96:
  // 10 % of the queries visit more than 12 288 objects; 2^15 slots serve them in the first tier, 1.18 M vs 0.90 M queries/s)
  const uint32_t bits0 = ix->hash_bits_auto ? (ix->n > 4000000ull ? 15u : 14u) : ix->hash_bits;
  uint32_t tier_bits[2] = {bits0, 17};
  uint32_t tier_queue[2] = {ix->queue_cap, 4096};
  int n_tiers = ix->onchip_tiers >= 2 ? 2 : 1;
  if (tier_bits[0] >= 17) n_tiers = 1;
  for (int t = 0; t < n_tiers; t++) {
    a.hash_bits = tier_bits[t];
    a.hash_limit = (uint32_t)((3ull << a.hash_bits) / 4);
    a.queue_cap = tier_queue[t];
    a.work_counter = ws + 2 * t;
    a.overflow_count = ws + 2 * t + 1;
    a.overflow_list = lists[t & 1];
    if (fused_seeds && t > 0) a.seeds = d_sel;   // written by the first tier
    a.query_list = t == 0 ? nullptr : lists[(t - 1) & 1];
    a.query_list_count = t == 0 ? nullptr : ws + 2 * (t - 1) + 1;
    size_t smem = extra + (size_t)a.queue_cap * 8;
    // the on-chip tiers of the common case run the lean kernel: rows of 5..32 chunks, head-table adjacency, set semantics
    // (epsilon >= 0), results in one warp's registers (k <= 128), a seed list that is one round
    const bool fast = ix->fast_kernel && ix->chunks >= FAST_MIN_CHUNKS && ix->chunks <= 32 && cap <= NGTGPU_HEAD_WIDTH &&
                      a.coef >= 1.0f && k <= 32 * FAST_KL_MAX && n_seeds <= 32 * FAST_WARPS && cap <= 32 * FAST_WARPS &&   // one edge per thread
                      (ix->n + 1) * (uint64_t)ix->row_bytes < (1ull << 36);   // 32-bit row offsets in 16-byte units
    const int fast_ch = ix->chunks <= 8 ? 1 : ix->chunks <= 16 ? 2 : 4;
    // two warps per query (16 CTAs per SM) for narrow rows whose rounds fit 64 threads (edge cap and seeds <= 64)
    // two warps per query whenever a round fits 64 threads (measured: uint8 128-byte rows 4.09 -> 3.35 ms per 10k batch,
    // 12.5M shard 7.66 -> 6.73 ms, glove-shape 448-byte rows 4.03 -> 3.74 ms); fast_warps = 4 keeps four
    const bool two_fit = cap <= 64 && n_seeds <= 64;
    // one warp per query (32 CTAs per SM, the back of the unchecked set in a global slab) in the first tier, when asked for
    const bool one_fit = two_fit && fast_ch == 1 && t == 0 && ix->fast_warps == 1;
    // (two warps also take rounds of up to 128 edges, two per thread, when asked for)
    const bool two_asked = ix->fast_warps == 2 && cap <= 128 && n_seeds <= 128;
    // (result lists of more than 32 keys -- construction and refinement searches -- come with four warps only)
    const int fast_w = k > 32 ? FAST_WARPS : one_fit ? 1 : (two_fit && ix->fast_warps != 4) || two_asked ? 2 : FAST_WARPS;
    if (fast) smem = (size_t)fast_w * fast_stage_per_warp(fast_ch, fast_w) + (fast_w == 1 ? 0 : (size_t)a.queue_cap * 8);
    if (smem > 200 * 1024)
      NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: working set does not fit shared memory; lower queue_cap/size");
    l.smem = smem;
    int blocks = 0;
    cudaError_t e = fast ? dispatch_fast(ix->acc_kind, a, fast_ch, fast_w, 0, smem, stream, 1, &blocks)
                         : dispatch(ix->acc_kind, a, l, 1, &blocks);
    if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("search occupancy query: ") + cudaGetErrorString(e));
    if (blocks < 1) blocks = 1;
    if (t > 0 && blocks > (fast ? 4 : 2)) blocks = fast ? 4 : 2;
    if (t == 0 && fast && ix->fast_ctas_per_sm > 0 && blocks > ix->fast_ctas_per_sm) blocks = ix->fast_ctas_per_sm;   // the overflow tier serves few queries: keep its slabs small
    uint64_t grid = (uint64_t)blocks * ix->sm_count;
    if (grid > nq) grid = nq;
    if (grid == 0) return NGTGPU_OK;
    l.grid = (unsigned)grid;
    // visited-hash slabs of this tier: one per CTA, in global memory (they live in L2)
    uint32_t *slabs = nullptr;
    const size_t slab_bytes = (size_t)grid * ((size_t)4 << a.hash_bits);
    const size_t queue_bytes = fast && fast_w == 1 ? (size_t)grid * a.queue_cap * 8 : 0;
    NGTGPU_TRY(ngtgpu_scratch(ix, t == 0 ? SCR_HASH0 : SCR_HASH1, slab_bytes + queue_bytes, (void **)&slabs));
    a.hash_slabs = slabs;
    a.queue_slabs = queue_bytes ? reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(slabs) + slab_bytes) : nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ix->timing && t == 0) {
      CUDA_TRY(cudaEventCreate(&ev0));
      CUDA_TRY(cudaEventCreate(&ev1));
      CUDA_TRY(cudaEventRecord(ev0, stream));
    }
    e = fast ? dispatch_fast(ix->acc_kind, a, fast_ch, fast_w, l.grid, smem, stream, 0, nullptr) : dispatch(ix->acc_kind, a, l, 0, nullptr);
    if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("search kernel launch: ") + cudaGetErrorString(e));
    ix->launches++;
    if (ix->timing && t == 0) {
      CUDA_TRY(cudaEventRecord(ev1, stream));
      std::lock_guard<std::mutex> lock(ix->state_mutex);
      ix->timing_events.push_back(ev0);
      ix->timing_events.push_back(ev1);
    }
  }

  // ---- HBM tier: exact bitmap + large queue in global memory for whatever is still left
  const int last = n_tiers - 1;
  const uint64_t bitmap_words = ((ix->n + 1 + 31) / 32 + 3) & ~(uint64_t)3;
  const uint32_t big_ctas = (uint32_t)ix->sm_count;
  uint32_t big_queue = BIG_TIER_QUEUE;
  size_t big_bytes = (size_t)big_ctas * (bitmap_words * 4 + (size_t)big_queue * 8);
  uint8_t *big = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEARCH_BIG, big_bytes, (void **)&big));
  SearchArgs b = a;
  if (fused_seeds) b.seeds = d_sel;
  b.work_counter = ws + 4;
  b.query_list = lists[last & 1];
  b.query_list_count = ws + 2 * last + 1;
  b.overflow_list = nullptr;
  b.overflow_count = nullptr;
  b.queue_cap = big_queue;
  b.big_queues = reinterpret_cast<uint64_t *>(big);
  b.big_bitmaps = reinterpret_cast<uint32_t *>(big + (size_t)big_ctas * big_queue * 8);
  b.bitmap_words = bitmap_words;
  SearchLaunch lb = l;
  lb.ws = 1;
  lb.grid = big_ctas;
  lb.smem = extra;
  cudaError_t e = dispatch(ix->acc_kind, b, lb, 0, nullptr);
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
  if (ix->chunks <= 32 && n_seeds <= 32 && ix->n_pivots <= 512) {
    // small tables of short rows: one warp per query walks the whole table (exact distances); measured against the tile
    // scan on 10k queries x 128-d: 256 pivots 0.19 vs 0.29 ms, 512 equal, 1024 0.75 vs 0.56 ms
    SeedArgs sa;
    sa.queries = d_queries;
    sa.pivots = ix->d_pivot_rows;
    sa.pivot_ids = ix->d_pivot_ids;
    sa.nq = nq;
    sa.n_pivots = ix->n_pivots;
    sa.row_bytes = ix->row_bytes;
    sa.chunks = ix->chunks;
    sa.k = n_seeds;
    sa.dtype = ix->distance_type;
    sa.seeds = seeds;
    cudaError_t e = dispatch_seeds(ix->acc_kind, sa, stream);
    if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("seed selection launch: ") + cudaGetErrorString(e));
    ix->launches++;
    *d_seeds = seeds;
    return NGTGPU_OK;
  }
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEED_DISTS, (size_t)nq * (n_seeds + 1) * sizeof(float) + 256, (void **)&sd));
  ScanParams p;
  p.d_queries = d_queries;
  p.nq = nq;
  p.d_rows = ix->d_pivot_rows;
  p.n_rows = ix->n_pivots;
  p.d_id_map = ix->d_pivot_ids;
  p.k = n_seeds;
  p.radius = -1.0f;
  p.approx = 1;
  p.d_ids = seeds;
  p.d_dists = sd + 64;
  p.d_counts = reinterpret_cast<uint32_t *>(sd + 64 + (size_t)nq * n_seeds);
  NGTGPU_TRY(ngtgpu_scan_topk(ix, p, stream));
  *d_seeds = seeds;
  return NGTGPU_OK;
}

// prepared (padded, object-type) query rows already on the device, e.g. stored objects: seeds from the pivot table,
// then the traversal. Used by the batched self-search of refineANNG (graph_ops.cu).
int ngtgpu_search_prepared(ngtgpu_index *ix, const uint8_t *d_queries, uint32_t nq, const ngtgpu_search_params *params,
                           uint32_t n_seeds, uint32_t *d_ids, float *d_dists, uint32_t *d_counts, cudaStream_t stream) {
  if (nq == 0) return NGTGPU_OK;
  uint32_t *s = nullptr;
  uint32_t ns = n_seeds;
  if (ns > ix->n_pivots) ns = ix->n_pivots;
  NGTGPU_TRY(select_seeds(ix, d_queries, nq, ns, &s, stream));
  return ngtgpu_traverse(ix, d_queries, nq, params, s, ns, d_ids, d_dists, d_counts, nullptr, stream);
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
    d_seeds = nullptr;   // the traversal takes the nearest pivots of the seed table
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
    uint32_t *ws = nullptr;
    NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEARCH_WS, 16, (void **)&ws));   // this lane's counters (already sized by the traversal)
    uint32_t h_ws[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(ids, d_ids, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(dists, d_dists, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(counts, d_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    if (stats) CUDA_TRY(cudaMemcpyAsync(stats, d_stats, (size_t)nq * 12, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_ws, ws, sizeof(h_ws), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    ix->last_overflows = h_ws[1];
    if (h_ws[5])
      NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: " + std::to_string(h_ws[5]) +
                                        " queries outgrew the HBM working set (unchecked queue > 2^18 entries)");
  }
  return NGTGPU_OK;
}

extern "C" int ngtgpu_search(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                             const ngtgpu_search_params *params, const uint32_t *seeds, uint32_t n_seeds, uint32_t *ids,
                             float *dists, uint32_t *counts, uint32_t *stats) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ngtgpu_lane_guard lane(ix);   // concurrent callers each get a stream and a scratch set of their own
  NGTGPU_TRY(lane.status);
  return search_common(ix, queries, query_type, nq, params, seeds, n_seeds, ids, dists, counts, stats, false, lane.stream());
}

int ngtgpu_search_keys_device(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                              const ngtgpu_search_params *params, uint32_t n_seeds, uint32_t id_offset, uint64_t *d_keys,
                              uint32_t *d_counts, cudaStream_t stream) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!params || !queries || !d_keys || !d_counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: null argument");
  if (!ix->d_objects || ix->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: the index holds no objects");
  if (nq == 0 || params->size == 0) return NGTGPU_OK;
  if ((uint64_t)id_offset + ix->n > 0xffffffffull) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: global ids do not fit 32 bits");
  uint8_t *d_q = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_QUERIES, (size_t)nq * ix->row_bytes, (void **)&d_q));
  NGTGPU_TRY(ngtgpu_prepare_queries(ix, queries, query_type, nq, true, d_q, stream));
  // a query that outgrows every tier leaves no keys behind: start from all-KEY_NONE
  CUDA_TRY(cudaMemsetAsync(d_keys, 0xff, (size_t)nq * params->size * 8, stream));
  return ngtgpu_traverse(ix, d_q, nq, params, nullptr, n_seeds, nullptr, nullptr, d_counts, nullptr, stream, d_keys, id_offset);
}

extern "C" int ngtgpu_search_device(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                                    const ngtgpu_search_params *params, const uint32_t *seeds, uint32_t n_seeds,
                                    uint32_t *ids, float *dists, uint32_t *counts, uint32_t *stats, void *stream) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  return search_common(ix, queries, query_type, nq, params, seeds, n_seeds, ids, dists, counts, stats, true,
                       (cudaStream_t)stream);
}

// ---- device timing of the traversal kernel (CUDA events on the launching stream) -----------------------
extern "C" int ngtgpu_index_set_timing(ngtgpu_index *ix, int enabled) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ix->timing = enabled != 0;
  return NGTGPU_OK;
}

// Sums and clears the recorded (start, stop) pairs: total milliseconds and number of traversal launches.
extern "C" int ngtgpu_index_pop_timing(ngtgpu_index *ix, double *total_ms, uint64_t *launches) {
  if (!ix || !total_ms || !launches) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_pop_timing: null argument");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  double sum = 0.0;
  uint64_t cnt = 0;
  std::lock_guard<std::mutex> lock(ix->state_mutex);
  for (size_t i = 0; i + 1 < ix->timing_events.size(); i += 2) {
    CUDA_TRY(cudaEventSynchronize(ix->timing_events[i + 1]));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ix->timing_events[i], ix->timing_events[i + 1]));
    sum += ms;
    cnt++;
    cudaEventDestroy(ix->timing_events[i]);
    cudaEventDestroy(ix->timing_events[i + 1]);
  }
  ix->timing_events.clear();
  *total_ms = sum;
  *launches = cnt;
  return NGTGPU_OK;
}

// The seeds the engine would use for these queries (nearest pivots of the seed table): lets a caller run
// the reference's GraphIndex::search(sc, seeds) from identical starting points.
extern "C" int ngtgpu_select_seeds(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq, uint32_t n_seeds,
                                   uint32_t *seeds_out) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!queries || !seeds_out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_select_seeds: null buffer");
  if (nq == 0) return NGTGPU_OK;
  uint8_t *d_q = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_QUERIES, (size_t)nq * ix->row_bytes, (void **)&d_q));
  NGTGPU_TRY(ngtgpu_prepare_queries(ix, queries, query_type, nq, false, d_q, ix->stream));
  uint32_t *s = nullptr;
  if (n_seeds > ix->n_pivots) n_seeds = ix->n_pivots;
  NGTGPU_TRY(select_seeds(ix, d_q, nq, n_seeds, &s, ix->stream));
  CUDA_TRY(cudaMemcpyAsync(seeds_out, s, (size_t)nq * n_seeds * 4, cudaMemcpyDeviceToHost, ix->stream));
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  return NGTGPU_OK;
}

// development aid: per-phase cycle counters of warp 0 (nq x 8 words, device buffer owned by the caller)
extern "C" int ngtgpu_index_set_phase_profile(ngtgpu_index *ix, uint32_t *d_buffer) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ix->d_prof = d_buffer;
  return NGTGPU_OK;
}

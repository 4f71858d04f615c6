// ngtgpu_internal.cuh -- shared declarations of the sm_100a engine (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ngtgpu.h"

// ---- error plumbing ------------------------------------------------------------------------------
void ngtgpu_set_error(const std::string &msg);
#define NGTGPU_FAIL(code, msg)   \
  do {                           \
    ngtgpu_set_error(msg);       \
    return (code);               \
  } while (0)
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ngtgpu_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
      return NGTGPU_ERR_CUDA;                                                                   \
    }                                                                                           \
  } while (0)
#define NGTGPU_TRY(expr)         \
  do {                           \
    int _rc = (expr);            \
    if (_rc != NGTGPU_OK) return _rc; \
  } while (0)

// ---- accumulate kinds: which sums a (object type, distance type) pair needs ---------------------
// F_L2  : sum (a-b)^2                     PrimitiveComparator.h:143-198
// F_DOT : sum a*b                         :446-477   (Normalized Cosine / Angle / L2)
// F_COS : sum a*a, sum b*b and sum a*b    :487-553   (Cosine, Angle)
// U8_L2 : exact integer sum (a-b)^2       :200-223
// U8_HAM: popcount(a^b)                   :340-353
#define NGTGPU_HEAD_WIDTH 128   // edges per node in the fixed-stride adjacency table read by the traversal kernel

enum AccKind { ACC_F_L2 = 0, ACC_F_DOT = 1, ACC_F_COS = 2, ACC_U8_L2 = 3, ACC_U8_HAM = 4 };

// scratch slots (grown on demand, reused between calls)
enum ScratchSlot {
  SCR_QUERIES = 0,     // prepared (padded, cast, normalised) query rows
  SCR_SEEDS = 1,       // nq x n_seeds seed ids from the pivot table
  SCR_SEED_DISTS = 2,  // nq x n_seeds (discarded) + counts
  SCR_PARTIAL = 3,     // partial top-k keys of the exhaustive scan
  SCR_IO = 4,          // device copies of host-side results (ids | dists | counts | stats)
  SCR_RAW_QUERIES = 5, // host queries as given, before preparation
  SCR_SEARCH_WS = 6,   // counters + overflow list of the traversal kernel
  SCR_SEARCH_BIG = 7,  // global-memory working sets of the overflow tier
  SCR_HASH0 = 8,       // visited-hash slabs of the first traversal tier (one per resident CTA)
  SCR_HASH1 = 9,       // ... of the second tier
  SCR_TC_MISC = 10,    // tensor-core kNN: flags
  SCR_TC_QUERY = 11,   // ... packed query operand + norms
  SCR_TC_CAND = 12,    // ... candidate lists
  SCR_COUNT = 13
};

// Concurrent host-pointer searches on one handle (the reference's contract: many threads search a read-only index,
// SURVEY.md 8b "Threading") each take a LANE: a stream and a scratch set of their own. Lane 0 is the index's own
// stream / scratch (what the device-pointer entry points and the construction functions use); the others are created
// the first time two calls overlap. A call that finds every lane busy waits for lane 0.
#define NGTGPU_LANES 3
struct ngtgpu_lane {
  void *d_scratch[32] = {nullptr};
  size_t scratch_bytes[32] = {0};
  cudaStream_t stream = nullptr;
};

struct ngtgpu_index {
  int device = 0;
  int object_type = 0;
  int distance_type = 0;
  int acc_kind = 0;
  uint32_t dim = 0;          // elements per object as the caller sees them
  uint32_t padded_dim = 0;   // ObjectSpace.h:249
  uint32_t elem_size = 0;    // 1 or 4
  uint32_t row_bytes = 0;    // padded_dim * elem_size, multiple of 16
  uint32_t chunks = 0;       // row_bytes / 16
  uint32_t group = 0;        // lanes that share one row: min(32, pow2ceil(chunks))
  bool normalizes = false;   // ObjectSpaceRepository.h:356-441 sets normalization for Normalized* types
  uint64_t n = 0;            // objects 1..n
  uint8_t *d_objects = nullptr;   // (n+1) x row_bytes, row 0 zero
  uint8_t *d_valid = nullptr;     // (n+1) bytes, 0 = empty slot; nullptr when nothing was removed
  uint64_t *d_row_ptr = nullptr;  // n+2
  uint32_t *d_col = nullptr;
  uint32_t *d_head = nullptr;     // (n+1) x 128: the first 128 edges of every node, zero padded (one coalesced read)
  uint64_t nnz = 0;
  int64_t edge_size_for_search = 40;   // Graph.h:401 defaults
  int64_t dyn_base = 30;
  int64_t dyn_rate = 20;
  // seed table
  uint32_t n_pivots = 0;
  uint8_t *d_pivot_rows = nullptr;     // n_pivots x row_bytes
  uint32_t *d_pivot_ids = nullptr;
  // traversal working-set sizing (on-chip tier)
  uint32_t hash_bits = 14;             // visited hash slots = 1 << hash_bits (4 B each, per-CTA slab in L2)
  bool hash_bits_auto = true;          // until ngtgpu_index_set_search_workspace: 15 for indexes over 4M objects (deeper searches)
  uint32_t queue_cap = 512;            // unchecked queue entries (8 B each, shared memory)
  uint32_t stage_bytes = 16384;        // shared-memory staging area the TMA engine fills with neighbour rows
  int onchip_tiers = 2;                // 1: overflow goes straight to the HBM tier (tests)
  bool fuse_seeds = false;             // seed selection inside the lean traversal kernel instead of its own launch
  int fast_warps = 0;                  // warps per query of the lean kernel: 0 = two when a round fits 64 threads, else four; 2 or 4 as asked (ngtgpu_index_set_fast_shape)
  int fast_ctas_per_sm = 0;            // cap on resident CTAs per SM of its first tier (0: what fits)
  bool fast_kernel = true;             // first tier of the common case on search_fast_kernel (off: tests of the general kernel)
  std::atomic<uint64_t> last_overflows{0};   // queries of the last call that fell to the global-memory tier
  // capacities of the graph / pivot buffers (elements): ngtgpu_index_set_graph and the seed table builders reuse them --
  // the construction loop sets a graph and a seed table once per batch of 200 objects
  uint64_t row_ptr_cap = 0, col_cap = 0, head_cap = 0, pivot_cap = 0;
  const void *graph_source = nullptr;   // caller's device row_ptr the graph was last set from (ngtgpu_index_insert_batch)
  // scratch
  void *d_scratch[SCR_COUNT] = {nullptr};
  size_t scratch_bytes[SCR_COUNT] = {0};
  cudaStream_t stream = nullptr;       // owned; host-pointer entry points run here
  int sm_count = 0;
  std::atomic<uint64_t> launches{0};
  ngtgpu_lane lanes[NGTGPU_LANES - 1];   // lanes 1.. (lane 0 = d_scratch / stream above)
  std::mutex lane_mutex[NGTGPU_LANES];
  std::mutex state_mutex;                // timing_events
  // optional device timing of the traversal kernel (bench.py's roofline leg)
  // tensor-core kNN (knn_tc.cu): the row operand packed once per set_objects
  bool tc_enabled = true;
  bool tc_rows_valid = false;
  uint8_t *d_tc_tiles = nullptr;
  float *d_tc_norms = nullptr;
  uint32_t tc_nseg = 0, tc_kchunks = 0;
  int tc_fold = 0;                     // L2: squared norms folded into the GEMM as an extra k-chunk
  int tc_i8 = 0;                       // integer kinds: u8 x u8 -> s32 operands (tcgen05 kind::i8)
  int tc_i8_m = 0, tc_i8_w = 0;        // ... bound of the integer row norms, and the K slots that carry (bound - norm) / 2
  float tc_max_norm = 0.f;
  uint64_t tc_batches = 0;             // batches answered by the tensor-core path
  uint32_t *d_prof = nullptr;          // development aid: per-phase cycle counters of the traversal kernel
  bool timing = false;
  std::vector<cudaEvent_t> timing_events;   // start, stop, start, stop ...
};

int ngtgpu_scratch(ngtgpu_index *ix, int slot, size_t bytes, void **out);   // of the calling thread's lane

// Takes a free lane of `ix` for the calling thread until it goes out of scope; ngtgpu_scratch then serves that lane.
struct ngtgpu_lane_guard {
  ngtgpu_index *ix;
  int lane;
  ngtgpu_lane *prev;
  explicit ngtgpu_lane_guard(ngtgpu_index *index);
  ~ngtgpu_lane_guard();
  cudaStream_t stream() const;
  int status = NGTGPU_OK;   // creating the lane's stream can fail
};
int ngtgpu_check_device(ngtgpu_index *ix);   // cudaSetDevice + sanity

// queries (host or device, float or uint8, `dim` wide) -> padded object-type rows in HBM.
int ngtgpu_prepare_queries(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq, bool on_device,
                           uint8_t *d_out, cudaStream_t stream);

// exhaustive top-k of prepared queries over `rows` (n_rows x row_bytes; the id of row r is
// id_map[r] when given, else first_row_id + r). exclude_self: skip the row whose id equals
// self_base + query index (kNN-graph construction, Index.h:839-856 drops the object itself).
struct ScanParams {
  const uint8_t *d_queries = nullptr;
  uint32_t nq = 0;
  const uint8_t *d_rows = nullptr;
  uint64_t n_rows = 0;
  uint32_t first_row_id = 0;
  const uint8_t *d_valid = nullptr;     // indexed by row id
  const uint32_t *d_id_map = nullptr;
  uint32_t k = 0;
  float radius = -1.0f;
  int exclude_self = 0;
  uint32_t self_base = 0;
  int approx = 0;                       // float kinds: skip the exact re-evaluation (ranking pivots for seeds)
  uint32_t *d_ids = nullptr;
  float *d_dists = nullptr;
  uint32_t *d_counts = nullptr;
};
int ngtgpu_scan_topk(ngtgpu_index *ix, const ScanParams &p, cudaStream_t stream);
int ngtgpu_scan_topk_tc(ngtgpu_index *ix, const ScanParams &p, cudaStream_t stream, int *used);

// graph traversal over prepared queries; everything in HBM.
int ngtgpu_traverse(ngtgpu_index *ix, const uint8_t *d_queries, uint32_t nq, const ngtgpu_search_params *params,
                    const uint32_t *d_seeds, uint32_t n_seeds, uint32_t *d_ids, float *d_dists, uint32_t *d_counts,
                    uint32_t *d_stats, cudaStream_t stream, uint64_t *d_keys = nullptr, uint32_t id_offset = 0);
// device-pointer search whose results leave as 64-bit keys (ordered distance bits << 32 | id + id_offset, KEY_NONE padded)
// written by the traversal kernel itself: the send buffer of the row-sharded search's all-gather (shard.cu)
int ngtgpu_search_keys_device(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                              const ngtgpu_search_params *params, uint32_t n_seeds, uint32_t id_offset, uint64_t *d_keys,
                              uint32_t *d_counts, cudaStream_t stream);

int64_t ngtgpu_effective_edge_size(const ngtgpu_index *ix, const ngtgpu_search_params *p);

#ifdef __CUDACC__
// ---- (distance,id) keys: unsigned order == ObjectDistance::operator< (Common.h:1946-1952) --------
__host__ __device__ __forceinline__ uint32_t ord_of_bits(uint32_t b) {
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint32_t ord_of_float(float d) {
  d += 0.0f;  // -0 -> +0 so equal distances tie-break on id alone
  return ord_of_bits(__float_as_uint(d));
}
__device__ __forceinline__ float float_of_ord(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t make_key(float d, uint32_t id) { return ((uint64_t)ord_of_float(d) << 32) | id; }
__device__ __forceinline__ float key_dist(uint64_t k) { return float_of_ord((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_id(uint64_t k) { return (uint32_t)k; }
#define KEY_NONE 0xffffffffffffffffull

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src, unsigned mask = 0xffffffffu) {
  uint32_t lo = __shfl_sync(mask, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(mask, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int o) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, o);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), o);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
  uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
  uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
  return ((uint64_t)hi << 32) | lo;
}

// ---- per-chunk (16 B) accumulation -----------------------------------------------------------------
// Float kinds keep TWO accumulators per sum, one for the even and one for the odd elements of the chunks a lane
// owns (the packed f32x2 pipes of sm_100 compute both halves in one FADD2 / FFMA2); a lane's total is their sum.
struct Sums {
  float f0, f0h;   // F_L2: sum sq ; F_DOT/F_COS: dot          (even elements, odd elements)
  float f1, f1h;   // F_COS: row norm^2
  uint32_t u;      // integer kinds
};
__device__ __forceinline__ Sums zero_sums() {
  Sums s;
  s.f0 = s.f0h = 0.f;
  s.f1 = s.f1h = 0.f;
  s.u = 0u;
  return s;
}

template <int ACC>
__device__ __forceinline__ void acc_chunk(Sums &s, const uint4 &q, const uint4 &r) {
  if (ACC == ACC_F_L2) {
    float d0 = __uint_as_float(q.x) - __uint_as_float(r.x);
    float d1 = __uint_as_float(q.y) - __uint_as_float(r.y);
    float d2 = __uint_as_float(q.z) - __uint_as_float(r.z);
    float d3 = __uint_as_float(q.w) - __uint_as_float(r.w);
    s.f0 = fmaf(d0, d0, s.f0);
    s.f0h = fmaf(d1, d1, s.f0h);
    s.f0 = fmaf(d2, d2, s.f0);
    s.f0h = fmaf(d3, d3, s.f0h);
  } else if (ACC == ACC_F_DOT) {
    s.f0 = fmaf(__uint_as_float(q.x), __uint_as_float(r.x), s.f0);
    s.f0h = fmaf(__uint_as_float(q.y), __uint_as_float(r.y), s.f0h);
    s.f0 = fmaf(__uint_as_float(q.z), __uint_as_float(r.z), s.f0);
    s.f0h = fmaf(__uint_as_float(q.w), __uint_as_float(r.w), s.f0h);
  } else if (ACC == ACC_F_COS) {
    float r0 = __uint_as_float(r.x), r1 = __uint_as_float(r.y), r2 = __uint_as_float(r.z), r3 = __uint_as_float(r.w);
    s.f0 = fmaf(__uint_as_float(q.x), r0, s.f0);
    s.f0h = fmaf(__uint_as_float(q.y), r1, s.f0h);
    s.f0 = fmaf(__uint_as_float(q.z), r2, s.f0);
    s.f0h = fmaf(__uint_as_float(q.w), r3, s.f0h);
    s.f1 = fmaf(r0, r0, s.f1);
    s.f1h = fmaf(r1, r1, s.f1h);
    s.f1 = fmaf(r2, r2, s.f1);
    s.f1h = fmaf(r3, r3, s.f1h);
  } else if (ACC == ACC_U8_L2) {
    uint32_t d;
    d = __vabsdiffu4(q.x, r.x); s.u = __dp4a(d, d, s.u);
    d = __vabsdiffu4(q.y, r.y); s.u = __dp4a(d, d, s.u);
    d = __vabsdiffu4(q.z, r.z); s.u = __dp4a(d, d, s.u);
    d = __vabsdiffu4(q.w, r.w); s.u = __dp4a(d, d, s.u);
  } else {
    s.u += __popc(q.x ^ r.x) + __popc(q.y ^ r.y) + __popc(q.z ^ r.z) + __popc(q.w ^ r.w);
  }
}

// One accumulator per sum, element after element: for values that only FILTER (the tile pass of the exhaustive scan,
// which re-evaluates survivors in the engine's order) -- half the registers of the two-accumulator form.
template <int ACC>
__device__ __forceinline__ void acc_chunk_seq(Sums &s, const uint4 &q, const uint4 &r) {
  if (ACC == ACC_F_L2) {
    float d0 = __uint_as_float(q.x) - __uint_as_float(r.x);
    float d1 = __uint_as_float(q.y) - __uint_as_float(r.y);
    float d2 = __uint_as_float(q.z) - __uint_as_float(r.z);
    float d3 = __uint_as_float(q.w) - __uint_as_float(r.w);
    s.f0 = fmaf(d0, d0, s.f0);
    s.f0 = fmaf(d1, d1, s.f0);
    s.f0 = fmaf(d2, d2, s.f0);
    s.f0 = fmaf(d3, d3, s.f0);
  } else if (ACC == ACC_F_DOT) {
    s.f0 = fmaf(__uint_as_float(q.x), __uint_as_float(r.x), s.f0);
    s.f0 = fmaf(__uint_as_float(q.y), __uint_as_float(r.y), s.f0);
    s.f0 = fmaf(__uint_as_float(q.z), __uint_as_float(r.z), s.f0);
    s.f0 = fmaf(__uint_as_float(q.w), __uint_as_float(r.w), s.f0);
  } else if (ACC == ACC_F_COS) {
    float r0 = __uint_as_float(r.x), r1 = __uint_as_float(r.y), r2 = __uint_as_float(r.z), r3 = __uint_as_float(r.w);
    s.f0 = fmaf(__uint_as_float(q.x), r0, s.f0);
    s.f0 = fmaf(__uint_as_float(q.y), r1, s.f0);
    s.f0 = fmaf(__uint_as_float(q.z), r2, s.f0);
    s.f0 = fmaf(__uint_as_float(q.w), r3, s.f0);
    s.f1 = fmaf(r0, r0, s.f1);
    s.f1 = fmaf(r1, r1, s.f1);
    s.f1 = fmaf(r2, r2, s.f1);
    s.f1 = fmaf(r3, r3, s.f1);
  } else {
    acc_chunk<ACC>(s, q, r);
  }
}

// The same sums with the packed pipes: one FADD2 + one FFMA2 per pair of elements (identical bits: every half is
// the IEEE operation of the scalar version).
__device__ __forceinline__ uint64_t pack_f2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub_f2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma_f2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <int ACC>
__device__ __forceinline__ void acc_chunk_packed(Sums &s, const uint4 &q, const uint4 &r) {
  if (ACC == ACC_F_L2 || ACC == ACC_F_DOT || ACC == ACC_F_COS) {
    const uint64_t q01 = pack_f2(q.x, q.y), q23 = pack_f2(q.z, q.w);
    const uint64_t r01 = pack_f2(r.x, r.y), r23 = pack_f2(r.z, r.w);
    uint64_t a0 = pack_f2(__float_as_uint(s.f0), __float_as_uint(s.f0h));
    if (ACC == ACC_F_L2) {
      const uint64_t d01 = sub_f2(q01, r01), d23 = sub_f2(q23, r23);
      a0 = fma_f2(d01, d01, a0);
      a0 = fma_f2(d23, d23, a0);
    } else {
      a0 = fma_f2(q01, r01, a0);
      a0 = fma_f2(q23, r23, a0);
    }
    unpack_f2(a0, s.f0, s.f0h);
    if (ACC == ACC_F_COS) {
      uint64_t a1 = pack_f2(__float_as_uint(s.f1), __float_as_uint(s.f1h));
      a1 = fma_f2(r01, r01, a1);
      a1 = fma_f2(r23, r23, a1);
      unpack_f2(a1, s.f1, s.f1h);
    }
  } else {
    acc_chunk<ACC>(s, q, r);
  }
}

// a lane's total of the chunks it owns
template <int ACC>
__device__ __forceinline__ void lane_total(Sums &s) {
  if (ACC == ACC_F_L2 || ACC == ACC_F_DOT || ACC == ACC_F_COS) s.f0 += s.f0h;
  if (ACC == ACC_F_COS) s.f1 += s.f1h;
}

// The engine's ONE summation order, used by every kernel that reports a distance, so the same
// (query, object) pair gives the same float bits whichever path computed it:
//   chunk c (16 B) is accumulated with fma by lane (c mod G) of a group of G = min(32, pow2ceil(chunks)) lanes,
//   chunks in increasing order, even and odd elements in separate accumulators that are added at the end
//   (lane_total); the G lane totals are then folded by an xor butterfly in float (as the reference folds its
//   SIMD lanes in float, PrimitiveComparator.h:153-193). Integer kinds are exact whatever the order.
template <int ACC, int G>
__device__ __forceinline__ void group_fold(Sums &s) {
  lane_total<ACC>(s);
  if (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s.u += __shfl_xor_sync(0xffffffffu, s.u, o);
  } else {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s.f0 += __shfl_xor_sync(0xffffffffu, s.f0, o);
    if (ACC == ACC_F_COS) {
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) s.f1 += __shfl_xor_sync(0xffffffffu, s.f1, o);
    }
  }
}

__device__ __forceinline__ double clamp_acos(double c) {  // PrimitiveComparator.h:571-593
  if (c >= 1.0) return 0.0;
  if (c <= -1.0) return acos(-1.0);
  return acos(c);
}

// The scalar tail of each comparator. The reference finishes in double and narrows to float
// (Common.h:47); for the square roots of a float sum, (float)sqrt((double)x) == sqrtf_rn(x) because
// double carries more than 2*24+2 bits, so __fsqrt_rn is used there.
// s = folded sums, qn = query norm^2 (F_COS only).
template <int ACC>
__device__ __forceinline__ float finish_distance(int dtype, const Sums &s, float qn) {
  if (ACC == ACC_U8_HAM) return (float)s.u;
  if (ACC == ACC_U8_L2) return __fsqrt_rn((float)s.u);
  if (ACC == ACC_F_L2) return __fsqrt_rn(s.f0);
  if (ACC == ACC_F_DOT) {
    double a = (double)s.f0;
    if (dtype == NGTGPU_DISTANCE_NORMALIZED_L2) {
      double v = 2.0 - 2.0 * a;
      return v < 0.0 ? 0.0f : (float)sqrt(v);
    }
    if (dtype == NGTGPU_DISTANCE_NORMALIZED_COSINE) {
      double v = 1.0 - a;
      return v < 0.0 ? 0.0f : (float)v;
    }
    return (float)clamp_acos(a);  // NGTGPU_DISTANCE_NORMALIZED_ANGLE
  }
  // ACC_F_COS
  double c = (double)s.f0 / sqrt((double)qn * (double)s.f1);
  if (dtype == NGTGPU_DISTANCE_COSINE) return (float)(1.0 - c);
  return (float)clamp_acos(c);  // NGTGPU_DISTANCE_ANGLE
}

__device__ __forceinline__ uint4 ldg16(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
// 16-byte load that does not allocate in L1 (object rows are touched once per query)
__device__ __forceinline__ uint4 ldg16_stream(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 zero16() { return make_uint4(0u, 0u, 0u, 0u); }

// Distance of one query row to one object row by a group of G lanes, both read from global memory
// (used where the query is not register resident: exact re-evaluation in the scan epilogue).
template <int ACC, int G>
__device__ __forceinline__ float group_distance_gmem(const uint8_t *q, const uint8_t *r, uint32_t chunks, int gl,
                                                     int dtype) {
  Sums s = zero_sums();
  float qn = 0.f;
  for (uint32_t c = gl; c < chunks; c += G) {
    uint4 a = ldg16(q + (size_t)c * 16);
    uint4 b = ldg16(r + (size_t)c * 16);
    acc_chunk<ACC>(s, a, b);
    if (ACC == ACC_F_COS) {
      float a0 = __uint_as_float(a.x), a1 = __uint_as_float(a.y), a2 = __uint_as_float(a.z), a3 = __uint_as_float(a.w);
      qn = fmaf(a0, a0, qn);
      qn = fmaf(a1, a1, qn);
      qn = fmaf(a2, a2, qn);
      qn = fmaf(a3, a3, qn);
    }
  }
  group_fold<ACC, G>(s);
  if (ACC == ACC_F_COS) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
  }
  return finish_distance<ACC>(dtype, s, qn);
}

// Insert `key` into the ascending array arr[0..n) of capacity k (the largest entry falls off when full); one warp.
__device__ __forceinline__ void sorted_insert_u64(uint64_t *arr, uint32_t &n_io, uint32_t k, uint64_t key, int lane) {
  uint32_t n = n_io;
  if (n == k) {
    if (key >= arr[k - 1]) return;
    n = k - 1;
  }
  uint32_t pos = 0;
  for (uint32_t i0 = 0; i0 < n; i0 += 32) {
    bool less = i0 + lane < n && arr[i0 + lane] < key;
    pos += __popc(__ballot_sync(0xffffffffu, less));
  }
  for (uint32_t hi = n; hi > pos;) {
    uint32_t lo = hi - pos > 32 ? hi - 32 : pos;
    uint32_t idx = lo + lane;
    uint64_t v = idx < hi ? arr[idx] : 0;
    __syncwarp();
    if (idx < hi) arr[idx + 1] = v;
    __syncwarp();
    hi = lo;
  }
  if (lane == 0) arr[pos] = key;
  __syncwarp();
  n_io = n + 1;
}

#endif  // __CUDACC__

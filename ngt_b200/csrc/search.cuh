// search.cuh -- the graph beam search kernel (one CTA per query), templated on the accumulate kind.
//
// Restates NeighborhoodGraph::search / searchReadOnlyGraph (lib/NGT/Graph.cpp:398-495, 499-638) for a
// batch of queries:
//
//   state per query   unchecked  min-queue of evaluated, not yet expanded nodes   (UncheckedSet)
//                     results    the k smallest (distance,id) seen so far          (ResultSet)
//                     visited    ids whose distance was evaluated                  (DistanceCheckedSet)
//   loop              pop the smallest unchecked node t; stop when t.d > explorationRadius;
//                     for the first min(deg, edgeSize) edges of t that are unvisited: mark, evaluate d;
//                     d <= explorationRadius -> unchecked; d <= radius -> results (keep k, then
//                     radius = k-th distance, explorationRadius = coef * radius).
//
// Because ids are unique, (distance,id) keys are totally ordered and the outcome of the loop depends
// only on SETS: `results` is the k smallest keys evaluated so far, and an unchecked entry whose distance
// exceeds the current explorationRadius can never be expanded (the radius only shrinks). That lets the
// neighbours of one node be evaluated in parallel and merged afterwards with results identical to the
// sequential loop whenever coef >= 1. For coef < 1 (negative epsilon) the order inside one adjacency
// list matters (results are gated by the shrinking explorationRadius), so filter and merge fall back to
// the reference's element-by-element order.
//
// One round of the loop inside the CTA (4 warps):
//   control   warp 0 merges the previous round's keys into results/unchecked, pops the next node.
//   filter    all threads: thread e reads edge e of the node -- one coalesced access to a fixed-stride
//             table of the first 64 edges of every node (the CSR only when edgeSize > 64) -- and
//             test-and-sets it in the visited set (shared-memory hash, atomicCAS).
//   gather    every warp copies the rows of its share of the surviving neighbours from HBM into a
//             shared-memory staging area with cp.async (16 B per lane: one warp instruction moves a whole
//             512-byte row, no registers held while the copies are in flight -- the GPU counterpart of the
//             reference's _mm_prefetch pipeline, Graph.cpp:446-456), waits for its own copies only, and
//             evaluates the distances from shared memory, G lanes per row.
//
// Where the state lives. Results and the unchecked queue are in shared memory (registers for k <= 32). The
// visited set of WS == 0 is an exact open-addressing hash in a per-CTA slab of GLOBAL memory: the slabs of
// all resident CTAs (8 per SM x 32 KB) stay in the 126 MB L2, whose capacity is of little use to the
// 512 MB of randomly gathered rows anyway, and taking the hash out of shared memory is what lets 8 CTAs
// share an SM instead of 3. A query that outgrows its slab or queue is appended to an overflow list and
// re-run by a later launch with a larger slab or, finally (WS == 1), with an exact bitmap and a large
// queue in HBM (the reference's own structures, Graph.h:751-799).
//
// The first tier of the common case (rows of 80..512 bytes, edge cap <= 128, epsilon >= 0, size <= 32) runs the leaner
// search_fast_kernel (search_fast.cuh) instead; the unchecked-set helpers below (sorted register front + unsorted
// back) are its queue.
#pragma once
#include <cfloat>

#include "ngtgpu_internal.cuh"

#ifndef SEARCH_WARPS
#define SEARCH_WARPS 4        // warps per query (1, 2 or 4)
#endif
#define SEARCH_THREADS (SEARCH_WARPS * 32)
#ifndef SEARCH_MIN_CTAS
#define SEARCH_MIN_CTAS (32 / SEARCH_WARPS)   // 64 registers per thread at 4 warps; nine CTAs (56 registers) spill and measured slower
#endif
#define SEARCH_CMAX 128       // upper bound of edges filtered / rows staged per round
#define SEARCH_HEAD 128       // edges per node in the fixed-stride adjacency table (== SEARCH_CMAX)
#define SEARCH_EPT (SEARCH_CMAX / SEARCH_THREADS)   // edges filtered per thread and round

struct SearchArgs {
  const uint8_t *objects;
  uint32_t row_bytes;
  uint32_t chunks;
  uint64_t n;
  const uint64_t *row_ptr;
  const uint32_t *col;
  const uint32_t *head;    // (n+1) x SEARCH_HEAD, zero padded
  const uint8_t *queries;  // prepared rows, nq x row_bytes
  const uint32_t *seeds;   // nq x n_seeds; null (lean kernel, first tier only): the nearest n_seeds pivots, selected by the kernel
  uint32_t n_seeds;
  uint32_t nq;
  uint32_t k;
  float coef;      // explorationCoefficient
  float radius;    // initial radius (FLT_MAX when unbounded)
  uint32_t edge_cap;
  int dtype;
  uint32_t hash_bits;   // WS == 0
  uint32_t hash_limit;  // max visited entries before overflow (WS == 0)
  uint32_t queue_cap;
  uint32_t stage_rows;  // rows the staging area holds (<= SEARCH_CMAX)
  uint32_t *ids;
  float *dists;
  uint32_t *counts;
  uint32_t *stats;  // nullable, nq x 3
  uint32_t *work_counter;
  const uint32_t *query_list;        // nullable: run only these queries (the overflow list of the previous tier)
  const uint32_t *query_list_count;
  uint32_t *overflow_list;           // WS == 0: where overflowing queries go
  uint32_t *overflow_count;
  uint32_t *failed_count;            // WS == 1: queries that outgrew even the HBM tier
  uint32_t *hash_slabs;              // WS == 0: gridDim.x x 2^hash_bits words, L2-resident
  uint64_t *queue_slabs;             // lean kernel, one warp per query: gridDim.x x queue_cap keys (the back of the unchecked set)
  uint32_t *big_bitmaps;             // WS == 1: gridDim.x x bitmap_words
  uint64_t *big_queues;              // WS == 1: gridDim.x x queue_cap
  uint64_t bitmap_words;
  uint32_t *prof;                    // nullable, nq x 8: cycles of warp 0 per phase (development aid)
  const uint8_t *pivots;             // seed table (seeds == null): n_pivots x row_bytes
  const uint32_t *pivot_ids;
  uint32_t n_pivots;
  uint32_t *seeds_out;               // nq x n_seeds: the seeds the kernel selected, for the later tiers
  uint64_t *keys_out;                // nullable, nq x k: results as (ordered distance bits << 32 | id + id_offset) keys, KEY_NONE
  uint32_t id_offset;                //   padded -- the all-gather send buffer of a row-sharded search (shard.cu); ids/dists are then
};                                   //   not written

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ---- cp.async: 16 bytes global -> shared without a register round trip (LDGSTS) -----------------------
__device__ __forceinline__ void cp_async_row16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
// the same with the destination already a 32-bit shared-window address (no generic->shared conversion per call)
__device__ __forceinline__ void cp_async_s16(uint32_t smem_addr, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
// base + a * b in one instruction (IMAD.WIDE.U32 with a 64-bit addend)
__device__ __forceinline__ const uint8_t *mad_wide_ptr(uint32_t a, uint32_t b, const uint8_t *base) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((uint64_t)(uintptr_t)base));
  return reinterpret_cast<const uint8_t *>((uintptr_t)r);
}
// 16-byte copy of which only the first src_bytes (0 or 16) are read from global memory; the rest is zero-filled
__device__ __forceinline__ void cp_async_s16z(uint32_t smem_addr, const void *gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t smem_addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr));
  return r;
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---- result list: k smallest keys ---------------------------------------------------------------------
// k <= 32: one key per lane of warp 0 (registers). k > 32: sorted array in shared memory.
struct ResultList {
  uint64_t reg;      // lane's key when k <= 32
  uint64_t *smem;    // k > 32
  uint32_t k;
  uint32_t n;
};

__device__ __forceinline__ void result_insert(ResultList &R, uint64_t key, int lane) {
  if (R.k <= 32) {
    uint32_t pos = __popc(__ballot_sync(0xffffffffu, R.reg < key));
    uint64_t up = shfl_up_u64(R.reg, 1);
    if ((uint32_t)lane == pos) R.reg = key;
    else if ((uint32_t)lane > pos) R.reg = up;
    if ((uint32_t)lane >= R.k) R.reg = KEY_NONE;
    if (R.n < R.k) R.n++;
  } else {
    uint32_t n = R.n;
    if (n == R.k) {
      if (key >= R.smem[R.k - 1]) return;
      n = R.k - 1;
    }
    uint32_t pos = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
      bool less = i0 + lane < n && R.smem[i0 + lane] < key;
      pos += __popc(__ballot_sync(0xffffffffu, less));
    }
    for (uint32_t hi = n; hi > pos;) {
      uint32_t lo = hi - pos > 32 ? hi - 32 : pos;
      uint32_t idx = lo + lane;
      uint64_t v = idx < hi ? R.smem[idx] : 0;
      __syncwarp();
      if (idx < hi) R.smem[idx + 1] = v;
      __syncwarp();
      hi = lo;
    }
    if (lane == 0) R.smem[pos] = key;
    __syncwarp();
    R.n = n + 1;
  }
}

__device__ __forceinline__ uint64_t result_kth(const ResultList &R) {  // valid when R.n == R.k
  if (R.k <= 32) return shfl_u64(R.reg, (int)R.k - 1);
  return R.smem[R.k - 1];
}

// Fold eight per-lane partial sums (eight rows) over the 32 lanes with 9 shuffles instead of 40: at each
// of the first three butterfly levels a lane keeps half of its rows and hands the other half to its
// partner. Every row is still summed over the pairs (lane, lane ^ 16), (.., ^ 8), (.., ^ 4), (.., ^ 2),
// (.., ^ 1) in that order, i.e. exactly group_fold<ACC, 32>, so the float bits are the same. On return
// the total of row ((lane >> 2) & 7) is in v[0] of every lane.
template <typename T>
__device__ __forceinline__ void fold8(T (&v)[8], int lane) {
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      T send = up ? v[i] : v[i + 4];
      T keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      T send = up ? v[i] : v[i + 2];
      T keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
    T send = up ? v[0] : v[1];
    T keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- unchecked set (UncheckedSet, Graph.h:757-799): a priority queue split in two ------------------------
// "front": the smallest unchecked keys, sorted, one per lane of warp 0 (registers; empty lanes hold KEY_NONE);
// "back":  everything else, unsorted, in `queue` (shared memory; global memory in the HBM tier).
// Invariant: every front key < T <= every back key, where T is the smallest key of the back (KEY_NONE when
// the back is empty). Popping is a lane shift; a new key that is not below T is appended to the back without
// looking at anything else; when the front runs dry it is refilled from the back in two passes.
struct Unchecked {
  uint64_t front;   // this lane's key
  uint32_t fn;      // keys in the front
  uint32_t qsize;   // keys in the back
  uint64_t T;       // smallest key of the back
  uint64_t *queue;
  uint32_t cap;
};

__device__ __forceinline__ uint64_t shfl_down_u64(uint64_t v, int d) {
  uint32_t lo = __shfl_down_sync(0xffffffffu, (uint32_t)v, d);
  uint32_t hi = __shfl_down_sync(0xffffffffu, (uint32_t)(v >> 32), d);
  return ((uint64_t)hi << 32) | lo;
}

// drop the back entries beyond the exploration radius: they can never be expanded (the radius only shrinks)
__device__ __forceinline__ void back_compact(Unchecked &U, float er, int lane) {
  uint32_t w = 0;
  for (uint32_t i0 = 0; i0 < U.qsize; i0 += 32) {
    uint64_t v = i0 + lane < U.qsize ? U.queue[i0 + lane] : KEY_NONE;
    bool keep = v != KEY_NONE && key_dist(v) <= er;
    uint32_t km = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) U.queue[w + __popc(km & lanemask_lt())] = v;
    __syncwarp();
    w += __popc(km);
  }
  U.qsize = w;
  if (w == 0) U.T = KEY_NONE;   // T is the back's minimum: it goes only when everything goes
}

// one key appended to the back by lane 0; false when the back is full even after compaction
__device__ __forceinline__ bool back_push(Unchecked &U, uint64_t key, float er, int lane) {
  if (U.qsize >= U.cap) {
    back_compact(U, er, lane);
    if (U.qsize >= U.cap) return false;
  }
  if (lane == 0) U.queue[U.qsize] = key;
  U.qsize++;
  __syncwarp();
  return true;
}

__device__ __forceinline__ void prefetch_head_row(const uint32_t *head, uint32_t id, uint32_t edge_cap) {
  const uint32_t *hp = head + (size_t)id * SEARCH_HEAD;
  const uint32_t lines = ((edge_cap < SEARCH_HEAD ? edge_cap : SEARCH_HEAD) + 31) / 32;
  for (uint32_t l = 0; l < lines; l++) asm volatile("prefetch.global.L2 [%0];" ::"l"(hp + l * 32));
}

// a key into the sorted front; the key that falls off the end (KEY_NONE unless the front was full) is returned
__device__ __forceinline__ uint64_t front_insert(Unchecked &U, uint64_t key, int lane) {
  const uint64_t evicted = shfl_u64(U.front, 31);
  const uint32_t pos = __popc(__ballot_sync(0xffffffffu, U.front < key));
  const uint64_t up = shfl_up_u64(U.front, 1);
  if ((uint32_t)lane == pos) U.front = key;
  else if ((uint32_t)lane > pos) U.front = up;
  if (U.fn < 32) U.fn++;
  return evicted;
}

// general insertion of one key (warp 0, every lane holds the same `key`); false on overflow of the back.
// head != nullptr: pull the head-table row of a key that enters the front towards L2 (it is likely to be popped)
__device__ __forceinline__ bool unchecked_insert(Unchecked &U, uint64_t key, float er, int lane, const uint32_t *head,
                                                 uint32_t edge_cap) {
  if (key >= U.T) return back_push(U, key, er, lane);
  if (U.fn == 32 && key > shfl_u64(U.front, 31)) {
    // between the front's largest key and T: it becomes the back's new minimum
    if (!back_push(U, key, er, lane)) return false;
    U.T = key;
    return true;
  }
  const uint64_t ev = front_insert(U, key, lane);
  if (head && lane == 0) prefetch_head_row(head, key_id(key), edge_cap);
  if (ev != KEY_NONE) {
    if (!back_push(U, ev, er, lane)) return false;
    U.T = ev;
  }
  return true;
}

// ascending bitonic sort of one key per lane
__device__ __forceinline__ uint64_t warp_sort_u64(uint64_t v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = shfl_xor_u64(v, j);
      const bool take_min = (((lane & k) == 0) == ((lane & j) == 0));
      const uint64_t lo = v < o ? v : o, hi = v < o ? o : v;
      v = take_min ? lo : hi;
    }
  }
  return v;
}

// The front is empty: move the smallest keys of the back into it. Pass 1: every lane finds the two smallest keys of
// its stride; with T' = the smallest of the lanes' SECOND keys, at most one key per lane (its first) is below T' and
// the overall minimum always is. Pass 2: the back is rewritten without them (and without keys beyond the exploration
// radius). T' stays in the back and is its new minimum.
__device__ __forceinline__ void front_refill(Unchecked &U, float er, int lane) {
  uint64_t m1 = KEY_NONE, m2 = KEY_NONE;
  for (uint32_t i = lane; i < U.qsize; i += 32) {
    const uint64_t v = U.queue[i];
    if (v < m1) {
      m2 = m1;
      m1 = v;
    } else if (v < m2) m2 = v;
  }
  uint64_t t = m2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t ot = shfl_xor_u64(t, o);
    if (ot < t) t = ot;
  }
  const bool mv = m1 < t;
  uint32_t w = 0;
  for (uint32_t i0 = 0; i0 < U.qsize; i0 += 32) {
    uint64_t v = i0 + lane < U.qsize ? U.queue[i0 + lane] : KEY_NONE;
    bool keep = v != KEY_NONE && v >= t && key_dist(v) <= er;
    uint32_t km = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) U.queue[w + __popc(km & lanemask_lt())] = v;
    __syncwarp();
    w += __popc(km);
  }
  U.qsize = w;
  U.T = w ? t : KEY_NONE;
  U.fn = __popc(__ballot_sync(0xffffffffu, mv));
  U.front = warp_sort_u64(mv ? m1 : KEY_NONE, lane);
}

// ---- visited set of WS == 0: exact hash in a per-CTA slab of global memory (L2 resident) ----------------
// The slab is an array of 32-byte buckets of eight ids. A lookup is ONE 32-byte read of the home bucket
// (ld.global.cg: served by L2, never a stale L1 line): the id is there, or the bucket has a free slot and
// the id is new (moving on to the next bucket only when all eight slots are taken by other ids). The
// insertion -- atomicCAS on the free slot -- does not gate anything and is issued later, while the row
// copies of the round are in flight.
struct BucketProbe {
  uint32_t bucket;   // where the id belongs (first bucket with a free slot)
  uint32_t slot;     // first free slot seen there
};

__device__ __forceinline__ bool hash_lookup(const uint32_t *hash, uint32_t bucket_bits, uint32_t nid, BucketProbe &bp) {
  const uint32_t bmask = (1u << bucket_bits) - 1u;
  uint32_t b = (nid * 2654435761u) >> (32 - bucket_bits);
  for (;;) {
    const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(hash + (size_t)b * 8));
    const uint4 hi = __ldcg(reinterpret_cast<const uint4 *>(hash + (size_t)b * 8) + 1);
    const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t free_slot = 8;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
      if (v[i] == nid) return true;   // visited
      if (v[i] == 0u) free_slot = i;
    }
    if (free_slot < 8) {
      bp.bucket = b;
      bp.slot = free_slot;
      return false;
    }
    b = (b + 1) & bmask;
  }
}

// returns false when the id turned out to be in the table already (a duplicate inside one round)
__device__ __forceinline__ bool hash_insert(uint32_t *hash, uint32_t bucket_bits, uint32_t nid, BucketProbe bp) {
  const uint32_t bmask = (1u << bucket_bits) - 1u;
  uint32_t b = bp.bucket, s0 = bp.slot;
  for (;;) {
    for (uint32_t i = s0; i < 8; i++) {
      uint32_t old = atomicCAS(&hash[(size_t)b * 8 + i], 0u, nid);
      if (old == 0u) return true;
      if (old == nid) return false;
    }
    b = (b + 1) & bmask;
    s0 = 0;
  }
}

// WS == 1: exact bitmap in HBM, test-and-set in one atomic
__device__ __forceinline__ bool bitmap_visit(uint32_t *bitmap, uint32_t nid) {
  uint32_t bit = 1u << (nid & 31);
  uint32_t old = atomicOr(&bitmap[nid >> 5], bit);
  return (old & bit) == 0;
}

template <int ACC, int G, int CPL, int WS>
__global__ void __launch_bounds__(SEARCH_THREADS, SEARCH_MIN_CTAS) search_kernel(const SearchArgs a) {
  constexpr int R = 32 / G;                 // rows per warp instruction (G < 32)
  constexpr int NCH = CPL > 0 ? CPL : 1;    // register-resident query chunks per lane
  // rows of <= 32 chunks: distances are evaluated by 8 lanes per row (four rows per warp instruction), lane j of a
  // row taking chunks j, j + 8, j + 16, j + 24 -- the adds of the first two butterfly levels of group_fold<ACC, 32>
  // become local (same operands, same order: same float bits), three shuffles remain
  constexpr bool ROW8 = (G == 32 && CPL == 1);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint32_t s_cand_ids[SEARCH_CMAX];
  __shared__ uint64_t s_cand_keys[SEARCH_CMAX];
  __shared__ uint32_t s_cand_n;
  __shared__ uint32_t s_edge_n;   // non-empty edges seen by the filter (head-table mode)
  __shared__ int s_state;         // 0 run, 1 finished, 2 overflow
  __shared__ uint32_t s_query;
  __shared__ int s_seeding;       // the round reads a seed list
  __shared__ uint32_t s_take;     // edges to filter this round
  __shared__ const uint32_t *s_src;  // where they are (head row, CSR slice or seed list)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int gl = lane % G;   // lane inside its row group
  const int grp = lane / G;  // row group inside the warp

  // ---- carve dynamic shared memory: [staging][results k>32][queue][query copy (CPL==0)]
  uint8_t *sp = smem_raw;
  uint8_t *stage = sp;
  sp += ((size_t)a.stage_rows * ((G == 32 && CPL > 0) ? 512u * (CPL > 0 ? CPL : 1) : a.row_bytes) + 127) & ~(size_t)127;
  uint64_t *s_results = reinterpret_cast<uint64_t *>(sp);
  sp += a.k > 32 ? (((size_t)a.k * 8 + 15) & ~(size_t)15) : 0;
  uint64_t *queue;
  uint32_t *hash = nullptr;
  uint32_t *bitmap = nullptr;
  if (WS == 0) {
    queue = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)a.queue_cap * 8;
    hash = a.hash_slabs + ((size_t)blockIdx.x << a.hash_bits);
  } else {
    queue = a.big_queues + (size_t)blockIdx.x * a.queue_cap;
    bitmap = a.big_bitmaps + (size_t)blockIdx.x * a.bitmap_words;
  }
  const uint4 *s_query_row = reinterpret_cast<const uint4 *>(sp);  // CPL == 0 only
  const uint32_t round_cap = SEARCH_CMAX;
  const bool ordered = a.coef < 1.0f;           // negative epsilon: keep the reference's element order
  const bool use_head = a.edge_cap <= SEARCH_HEAD;

  if (G == 32 && CPL > 0) {
    // padded tails of staged rows (chunks >= row chunks) are never written by the copies: make them zero once
    uint4 *z = reinterpret_cast<uint4 *>(stage);
    for (uint32_t i = tid; i < a.stage_rows * (512u * (CPL > 0 ? CPL : 1)) / 16; i += SEARCH_THREADS) z[i] = zero16();
    __syncthreads();
  }
  for (;;) {
    // ---- next query (dynamic scheduling over a persistent grid)
    if (tid == 0) {
      uint32_t w = atomicAdd(a.work_counter, 1u);
      uint32_t total = a.query_list ? *a.query_list_count : a.nq;
      s_query = w < total ? (a.query_list ? a.query_list[w] : w) : 0xffffffffu;
      s_state = 0;
      s_cand_n = 0;
      s_edge_n = 0;
    }
    __syncthreads();
    const uint32_t q = s_query;
    if (q == 0xffffffffu) break;

    // ---- per-query initialisation
    if (WS == 0) {
      uint4 *h4 = reinterpret_cast<uint4 *>(hash);
      for (uint32_t i = tid; i < (1u << a.hash_bits) / 4; i += SEARCH_THREADS) h4[i] = zero16();
    } else {
      uint4 *b4 = reinterpret_cast<uint4 *>(bitmap);
      for (uint64_t i = tid; i < a.bitmap_words / 4; i += SEARCH_THREADS) b4[i] = zero16();
    }
    const uint8_t *qrow = a.queries + (size_t)q * a.row_bytes;
    uint4 qreg[NCH];
    uint4 q8[4];   // ROW8 layout: lane (rr, j) holds query chunks j, j + 8, j + 16, j + 24
    float qn = 0.f;
    if (ROW8) {
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const uint32_t c = (lane & 7) + m * 8;
        q8[m] = c < a.chunks ? ldg16(qrow + (size_t)c * 16) : zero16();
      }
      qreg[0] = zero16();
    } else if (CPL > 0) {
#pragma unroll
      for (int i = 0; i < NCH; i++) {
        uint32_t c = gl + i * G;
        qreg[i] = c < a.chunks ? ldg16(qrow + (size_t)c * 16) : zero16();
      }
    } else {
      uint4 *dst = const_cast<uint4 *>(s_query_row);
      for (uint32_t c = tid; c < a.chunks; c += SEARCH_THREADS) dst[c] = ldg16(qrow + (size_t)c * 16);
      qreg[0] = zero16();
    }
    __syncthreads();
    if (ACC == ACC_F_COS) {
      // query norm^2 in the engine's summation order (PrimitiveComparator.h:487-553 recomputes it per call)
      if (ROW8) {
        // lane c owns chunk c here (one load, once per query) so the fold below is group_fold's
        const uint4 v = (uint32_t)lane < a.chunks ? ldg16(qrow + (size_t)lane * 16) : zero16();
        float a0 = __uint_as_float(v.x), a1 = __uint_as_float(v.y), a2 = __uint_as_float(v.z), a3 = __uint_as_float(v.w);
        qn = fmaf(a0, a0, qn);
        qn = fmaf(a1, a1, qn);
        qn = fmaf(a2, a2, qn);
        qn = fmaf(a3, a3, qn);
      } else if (CPL > 0) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
          float a0 = __uint_as_float(qreg[i].x), a1 = __uint_as_float(qreg[i].y), a2 = __uint_as_float(qreg[i].z),
                a3 = __uint_as_float(qreg[i].w);
          qn = fmaf(a0, a0, qn);
          qn = fmaf(a1, a1, qn);
          qn = fmaf(a2, a2, qn);
          qn = fmaf(a3, a3, qn);
        }
      } else {
        for (uint32_t c = gl; c < a.chunks; c += G) {
          uint4 v = s_query_row[c];
          float a0 = __uint_as_float(v.x), a1 = __uint_as_float(v.y), a2 = __uint_as_float(v.z), a3 = __uint_as_float(v.w);
          qn = fmaf(a0, a0, qn);
          qn = fmaf(a1, a1, qn);
          qn = fmaf(a2, a2, qn);
          qn = fmaf(a3, a3, qn);
        }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
    }

    // ---- control-warp state (registers of warp 0; other warps carry dead copies)
    ResultList res;
    res.reg = KEY_NONE;
    res.smem = s_results;
    res.k = a.k;
    res.n = 0;
    uint32_t qsize = 0;           // unchecked entries
    float radius = a.radius;      // sc.radius
    float er = a.coef * radius;   // explorationRadius (Graph.cpp:420)
    uint32_t visited_n = 0;
    uint32_t st_dist = 0, st_edge = 0, st_exp = 0;
    // adjacency cursor: the seeds first (setupDistances/setupSeeds, Graph.cpp:243-394), then popped nodes
    const uint32_t *cur = a.seeds + (size_t)q * a.n_seeds;
    uint32_t cur_deg = a.n_seeds, cur_pos = 0;
    bool seeding = true;
    bool head_round = false;      // the round in flight read a head-table row (its edge count comes from s_edge_n)
    uint32_t cand_n = 0;
    // per-phase cycle counters of warp 0: compiled in only with -DSEARCH_PHASE_PROFILE (they cost ten registers)
#ifdef SEARCH_PHASE_PROFILE
    uint32_t pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t spec_second = KEY_NONE, spec_third = KEY_NONE;
    long long tp = a.prof ? clock64() : 0;
#define PROF_MARK(i)                     \
  if (a.prof) {                          \
    long long _t = clock64();            \
    pf[i] += (uint32_t)(_t - tp);        \
    tp = _t;                             \
  }
#else
#define PROF_MARK(i)
#endif

    for (;;) {
      // ================= control (warp 0): merge the previous round, choose the next edges =================
      if (warp == 0) {
        bool overflow = false;
        PROF_MARK(7)
        if (head_round) st_edge += s_edge_n;
        visited_n += cand_n;
        st_dist += cand_n;
        if (cand_n) {
          if (!ordered && !seeding) {
            // set semantics: results first, then everything within the final explorationRadius
            for (uint32_t j0 = 0; j0 < cand_n; j0 += 32) {
              uint64_t key = j0 + lane < cand_n ? s_cand_keys[j0 + lane] : KEY_NONE;
              uint32_t m = __ballot_sync(0xffffffffu, key != KEY_NONE && key_dist(key) <= radius);
              while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                uint64_t kk = shfl_u64(key, src);
                if (key_dist(kk) > radius) continue;  // radius shrank meanwhile
                result_insert(res, kk, lane);
                if (res.n >= res.k) radius = key_dist(result_kth(res));
              }
            }
            er = a.coef * radius;
            for (uint32_t j0 = 0; j0 < cand_n && !overflow; j0 += 32) {
              uint64_t key = j0 + lane < cand_n ? s_cand_keys[j0 + lane] : KEY_NONE;
              bool acc = key != KEY_NONE && key_dist(key) <= er;
              uint32_t m = __ballot_sync(0xffffffffu, acc);
              uint32_t cnt = __popc(m);
              if (qsize + cnt > a.queue_cap) {
                // compact: entries beyond explorationRadius can never be expanded
                uint32_t w = 0;
                for (uint32_t i0 = 0; i0 < qsize; i0 += 32) {
                  uint64_t v = i0 + lane < qsize ? queue[i0 + lane] : KEY_NONE;
                  bool keep = v != KEY_NONE && key_dist(v) <= er;
                  uint32_t km = __ballot_sync(0xffffffffu, keep);
                  __syncwarp();
                  if (keep) queue[w + __popc(km & lanemask_lt())] = v;
                  __syncwarp();
                  w += __popc(km);
                }
                qsize = w;
              }
              if (qsize + cnt > a.queue_cap) {
                overflow = true;
                break;
              }
              if (acc) {
                queue[qsize + __popc(m & lanemask_lt())] = key;
                if (use_head) {
                  // its edges will be wanted when it is popped: pull that row of the head table towards L2 now
                  const uint32_t *hp = a.head + (size_t)key_id(key) * SEARCH_HEAD;
                  const uint32_t lines = ((a.edge_cap < SEARCH_HEAD ? a.edge_cap : SEARCH_HEAD) + 31) / 32;
                  for (uint32_t l = 0; l < lines; l++) asm volatile("prefetch.global.L2 [%0];" ::"l"(hp + l * 32));
                }
              }
              qsize += cnt;
            }
            __syncwarp();
          } else {
            // the reference's order: seeds (all go to unchecked, Graph.cpp:352-366) and coef < 1
            for (uint32_t j0 = 0; j0 < cand_n && !overflow; j0 += 32) {
              uint64_t key = j0 + lane < cand_n ? s_cand_keys[j0 + lane] : KEY_NONE;
              uint32_t m = __ballot_sync(0xffffffffu, key != KEY_NONE);
              while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                uint64_t kk = shfl_u64(key, src);
                float d = key_dist(kk);
                if (!seeding && d > er) continue;
                if (qsize >= a.queue_cap) {
                  overflow = true;
                  break;
                }
                if (lane == 0) queue[qsize] = kk;
                qsize++;
                if (d <= radius) {
                  result_insert(res, kk, lane);
                  if (!seeding && res.n >= res.k) {
                    radius = key_dist(result_kth(res));
                    er = a.coef * radius;
                  }
                }
              }
            }
            __syncwarp();
          }
        }
        cand_n = 0;
        head_round = false;
        PROF_MARK(0)
        // ---- next edges: the rest of the current list, or pop the smallest unchecked node
        bool finished = false;
        uint32_t take = 0;
        const uint32_t *src_ptr = nullptr;
        while (!overflow && !finished && take == 0) {
          if (cur_pos < cur_deg) {
            take = cur_deg - cur_pos;
            if (take > round_cap) take = round_cap;
            src_ptr = cur + cur_pos;
            cur_pos += take;
            break;
          }
          if (seeding) {
            // setupSeeds: radius from the seeds once k of them are within it (Graph.cpp:349-351)
            seeding = false;
            if (res.n >= res.k) radius = key_dist(result_kth(res));
            er = a.coef * radius;
          }
          uint64_t best = KEY_NONE;
          uint32_t bi = 0;
          for (uint32_t i = lane; i < qsize; i += 32) {
            uint64_t v = queue[i];
            if (v < best) {
              best = v;
              bi = i;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            uint64_t ob = shfl_xor_u64(best, o);
            uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob < best) {
              best = ob;
              bi = oi;
            }
          }
          if (best == KEY_NONE || key_dist(best) > er) {  // Graph.cpp:430-435
            finished = true;
            break;
          }
          if (lane == 0) queue[bi] = queue[qsize - 1];
          qsize--;
          __syncwarp();
          const uint32_t t = key_id(best);
          st_exp++;
#ifdef SEARCH_PHASE_PROFILE
          // development probe: how often is the popped node the one that was second best at the previous pop?
          // (slot 4 of the phase record: it was the second best; slot 2: it was the second or third best)
          if (a.prof) {
            if (best == spec_second) pf[4]++;
            if (best == spec_second || best == spec_third) pf[2]++;
            uint64_t s2 = KEY_NONE, s3 = KEY_NONE;   // the two smallest remaining keys
            for (uint32_t i = lane; i < qsize; i += 32) {
              uint64_t v = queue[i];
              if (v < s2) { s3 = s2; s2 = v; } else if (v < s3) s3 = v;
            }
            for (int o = 16; o > 0; o >>= 1) {
              uint64_t o2 = shfl_xor_u64(s2, o), o3 = shfl_xor_u64(s3, o);
              uint64_t lo = s2 < o2 ? s2 : o2, hi = s2 < o2 ? o2 : s2;
              uint64_t m3 = s3 < o3 ? s3 : o3;
              s2 = lo;
              s3 = hi < m3 ? hi : m3;
            }
            spec_second = s2;
            spec_third = s3;
          }
#endif
          PROF_MARK(1)
          if (use_head) {
            // the whole (capped) list is one row of the head table; empty slots are zero
            take = a.edge_cap < SEARCH_HEAD ? a.edge_cap : SEARCH_HEAD;
            src_ptr = a.head + (size_t)t * SEARCH_HEAD;
            head_round = true;
            cur_deg = 0;
            cur_pos = 0;
          } else {
            // longer lists (edgeSize > 64): walk the CSR slice (Graph.cpp:438 caps it)
            uint64_t b = 0, e = 0;
            if (lane < 2) b = a.row_ptr[(size_t)t + lane];
            e = shfl_u64(b, 1);
            b = shfl_u64(b, 0);
            uint64_t deg = e - b;
            if (deg > a.edge_cap) deg = a.edge_cap;
            cur = a.col + b;
            cur_deg = (uint32_t)deg;
            cur_pos = 0;
            st_edge += cur_deg;
          }
          PROF_MARK(1)
        }
        if (WS == 0 && !overflow && !finished && visited_n + take > a.hash_limit) overflow = true;
        if (lane == 0) {
          s_cand_n = 0;
          s_edge_n = 0;
          s_take = take;
          s_src = src_ptr;
          s_seeding = seeding ? 1 : 0;
          if (overflow) s_state = 2;
          else if (finished) s_state = 1;
        }
      }
      __syncthreads();  // (A) the round is published
      if (s_state != 0) break;
      const bool seeding_round = s_seeding != 0;

      // ================= filter: one edge per thread, looked up in the visited set =================
      // bit i: this thread found edge i new and its insertion is still to be issued
      uint32_t pending_mask = 0;
      uint32_t pending_id[SEARCH_EPT];
      BucketProbe bp[SEARCH_EPT];
#pragma unroll
      for (int i = 0; i < SEARCH_EPT; i++) {
        pending_id[i] = 0;
        bp[i].bucket = 0;
        bp[i].slot = 0;
      }
      {
        const uint32_t take = s_take;
        const uint32_t *src = s_src;
        const bool immediate = seeding_round;   // seed lists may repeat an id: insert at once so the second copy is seen
        if (!ordered) {
          // thread t looks at edges t, t + THREADS, ...: all edge loads first, then all bucket lookups
          uint32_t nid[SEARCH_EPT];
#pragma unroll
          for (int i = 0; i < SEARCH_EPT; i++) {
            const uint32_t e = (uint32_t)tid + i * SEARCH_THREADS;
            nid[i] = e < take ? __ldg(src + e) : 0u;
          }
          uint32_t new_mask = 0, n_valid = 0;
#pragma unroll
          for (int i = 0; i < SEARCH_EPT; i++) {
            const bool valid = nid[i] != 0u && nid[i] <= a.n;
            bool isnew = false;
            if (valid) {
              n_valid++;
              if (WS == 0) {
                isnew = !hash_lookup(hash, a.hash_bits - 3, nid[i], bp[i]);
                if (isnew && immediate) isnew = hash_insert(hash, a.hash_bits - 3, nid[i], bp[i]);
                else if (isnew) {
                  pending_mask |= 1u << i;
                  pending_id[i] = nid[i];
                }
              } else {
                isnew = bitmap_visit(bitmap, nid[i]);
              }
            }
            if (isnew) new_mask |= 1u << i;
          }
          // position of this thread's new ids in the candidate list: exclusive count inside the warp + the warp's base
          uint32_t before, warp_new, warp_valid;
          if (SEARCH_EPT == 1) {
            const uint32_t m = __ballot_sync(0xffffffffu, new_mask != 0);
            before = __popc(m & lanemask_lt());
            warp_new = __popc(m);
            warp_valid = __popc(__ballot_sync(0xffffffffu, n_valid != 0));
          } else {
            const uint32_t n_new = __popc(new_mask);
            uint32_t incl = n_new;
            warp_valid = n_valid;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += v;
              warp_valid += __shfl_xor_sync(0xffffffffu, warp_valid, o);
            }
            before = incl - n_new;
            warp_new = __shfl_sync(0xffffffffu, incl, 31);
          }
          uint32_t base = 0;
          if (lane == 0) {
            if (SEARCH_WARPS == 1) {
              s_cand_n = warp_new;
              s_edge_n = warp_valid;
            } else {
              if (warp_new) base = atomicAdd(&s_cand_n, warp_new);
              if (warp_valid) atomicAdd(&s_edge_n, warp_valid);
            }
          }
          base = __shfl_sync(0xffffffffu, base, 0) + before;
#pragma unroll
          for (int i = 0; i < SEARCH_EPT; i++)
            if (new_mask & (1u << i)) s_cand_ids[base++] = nid[i];
        } else if (warp == 0) {
          // element order kept: warp 0 walks the edges 32 at a time
          uint32_t cn = 0, en = 0;
          for (uint32_t e0 = 0; e0 < take; e0 += 32) {
            uint32_t nid = e0 + lane < take ? __ldg(src + e0 + lane) : 0u;
            const bool valid = nid != 0u && nid <= a.n;
            bool isnew = false;
            if (valid) {
              if (WS == 0) {
                isnew = !hash_lookup(hash, a.hash_bits - 3, nid, bp[0]);
                if (isnew) isnew = hash_insert(hash, a.hash_bits - 3, nid, bp[0]);
              } else {
                isnew = bitmap_visit(bitmap, nid);
              }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, isnew);
            en += __popc(__ballot_sync(0xffffffffu, valid));
            if (isnew) s_cand_ids[cn + __popc(m & lanemask_lt())] = nid;
            cn += __popc(m);
            __syncwarp();
          }
          if (lane == 0) {
            s_cand_n = cn;
            s_edge_n = en;
          }
        }
      }
      __syncthreads();  // (B) the candidate list is complete
      const uint32_t cn = s_cand_n;
      if (warp == 0) {
        cand_n = cn;
        PROF_MARK(3)
      }

      // ================= gather: each warp stages and evaluates its own slice of the rows =================
      // The staging area is split evenly between the warps (wrows rows each); candidates are taken in passes
      // of 4 * wrows, warp w owning the w-th slice of every pass, so no CTA-wide synchronisation is needed
      // between copying a row and reading it.
      {
        const uint32_t wrows = a.stage_rows / SEARCH_WARPS;
        // rows of the register-query path (G == 32, CPL > 0) are staged at a compile-time stride of 512 * CPL bytes
        // (>= row_bytes; the tail was zeroed once at kernel start), so shared addresses are 32-bit constants + shifts
        constexpr uint32_t SROW = 512u * (CPL > 0 ? CPL : 1);
        const uint32_t srow_bytes = (G == 32 && CPL > 0) ? SROW : a.row_bytes;
        uint8_t *wstage = stage + (size_t)warp * wrows * srow_bytes;
        const uint32_t wstage_s = (uint32_t)__cvta_generic_to_shared(wstage) + (uint32_t)lane * 16u;
        for (uint32_t j0 = warp * wrows; j0 < cn; j0 += SEARCH_WARPS * wrows) {
          const uint32_t nr = cn - j0 < wrows ? cn - j0 : wrows;
          if (G == 32 && CPL > 0) {
            // lane r fetches the id of the slice's r-th row once; rows are issued eight at a time, fully unrolled and
            // branch-free: a copy past the slice's last row, or of a chunk past the row's end, has source size 0
            // (nothing is read, the slot is zero-filled)
            const uint32_t my_id = (uint32_t)lane < nr ? s_cand_ids[j0 + lane] : 0u;
            const uint8_t *lane_src = a.objects + (size_t)lane * 16;
            const bool last_ok = (uint32_t)lane + (NCH - 1) * 32 < a.chunks;
            for (uint32_t r0 = 0; r0 < nr; r0 += 8) {
              const uint32_t left = nr - r0;
#pragma unroll
              for (int i = 0; i < 8; i++) {
                // (slices of the fold kernels, CPL <= 2, are whole blocks of eight slots; longer rows may have fewer)
                if (CPL > 2 && r0 + i >= wrows) break;
                const uint32_t id = __shfl_sync(0xffffffffu, my_id, (int)(r0 + i));   // r0 + i < 32: a slice has <= 32 rows
                const uint8_t *srow = lane_src + (size_t)id * a.row_bytes;
                const uint32_t drow = wstage_s + (r0 + i) * SROW;
                const bool in = (uint32_t)i < left;
#pragma unroll
                for (int c = 0; c < NCH; c++)
                  cp_async_s16z(drow + c * 512, srow + c * 512, (in && (c < NCH - 1 || last_ok)) ? 16u : 0u);
#ifdef SEARCH_EXPERIMENT_EXTRA_TRAFFIC
                // development probe: one more random row pulled from HBM into L2 per row copied (nobody waits for it)
                if (in) {
                  const uint32_t other = (id * 2654435761u + 12345u) % (uint32_t)a.n;
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(lane_src + (size_t)other * a.row_bytes));
                }
#endif
              }
            }
          } else if (G == 32) {
            for (uint32_t r = 0; r < nr; r++) {
              const uint8_t *srow = a.objects + (size_t)s_cand_ids[j0 + r] * a.row_bytes;
              uint8_t *drow = wstage + (size_t)r * a.row_bytes;
              for (uint32_t c = lane; c < a.chunks; c += 32) cp_async_row16(drow + (size_t)c * 16, srow + (size_t)c * 16);
            }
          } else {
            // short rows: the slice is one flat run of nr * chunks 16-byte pieces
            const uint32_t total = nr * a.chunks;
            for (uint32_t x = lane; x < total; x += 32) {
              const uint32_t r = x / a.chunks, c = x - r * a.chunks;
              cp_async_row16(wstage + (size_t)r * a.row_bytes + (size_t)c * 16,
                             a.objects + (size_t)s_cand_ids[j0 + r] * a.row_bytes + (size_t)c * 16);
            }
          }
          if (WS == 0 && pending_mask) {
            // the insertions of this thread's new ids, overlapped with the row copies in flight
#pragma unroll
            for (int i = 0; i < SEARCH_EPT; i++)
              if (pending_mask & (1u << i)) hash_insert(hash, a.hash_bits - 3, pending_id[i], bp[i]);
            pending_mask = 0;
          }
          cp_async_commit_wait_all();
          __syncwarp();
          if (ROW8) {
            // four rows per step, eight lanes each. Slots past the slice's last row hold stale rows of earlier
            // rounds (or the initial zeros): they are read like the others and their result is dropped; chunks
            // past the row's end read the zeroed tail against zero query chunks.
            const uint32_t rr = (uint32_t)lane >> 3;
            const uint32_t ra0 = wstage_s - (uint32_t)lane * 16u + rr * SROW + ((uint32_t)lane & 7u) * 16u;
            for (uint32_t r0 = 0; r0 < nr; r0 += 4) {
              Sums p[4];
#pragma unroll
              for (int m = 0; m < 4; m++) {
                p[m] = zero_sums();
                acc_chunk<ACC>(p[m], q8[m], lds16(ra0 + r0 * SROW + m * 128));
                lane_total<ACC>(p[m]);
              }
              Sums tot = zero_sums();
              if (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM) {
                tot.u = (p[0].u + p[2].u) + (p[1].u + p[3].u);
                tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 4);
                tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 2);
                tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 1);
              } else {
                tot.f0 = (p[0].f0 + p[2].f0) + (p[1].f0 + p[3].f0);
                tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 4);
                tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 2);
                tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 1);
                if (ACC == ACC_F_COS) {
                  tot.f1 = (p[0].f1 + p[2].f1) + (p[1].f1 + p[3].f1);
                  tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 4);
                  tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 2);
                  tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 1);
                }
              }
              const uint32_t r = r0 + rr;
              if ((lane & 7) == 0 && r < nr)
                s_cand_keys[j0 + r] = make_key(finish_distance<ACC>(a.dtype, tot, qn), s_cand_ids[j0 + r]);
            }
          } else if (CPL > 0 && CPL <= 2 && G == 32) {
            // eight rows at a time, folded together (fold8). Slots past the slice's last row hold stale rows of
            // earlier rounds (or the initial zeros): they are read like the others and their result is dropped;
            // chunks past the row's end read the zeroed tail. No predicates and no address selects in the loop.
            for (uint32_t r0 = 0; r0 < nr; r0 += 8) {
              Sums s[8];
              const uint32_t ra = wstage_s + r0 * SROW;
#pragma unroll
              for (int i = 0; i < 8; i++) {
                s[i] = zero_sums();
#pragma unroll
                for (int c = 0; c < NCH; c++) acc_chunk<ACC>(s[i], qreg[c], lds16(ra + i * SROW + c * 512));
                lane_total<ACC>(s[i]);
              }
              Sums tot = zero_sums();
              if (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM) {
                uint32_t v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = s[i].u;
                fold8<uint32_t>(v, lane);
                tot.u = v[0];
              } else {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = s[i].f0;
                fold8<float>(v, lane);
                tot.f0 = v[0];
                if (ACC == ACC_F_COS) {
#pragma unroll
                  for (int i = 0; i < 8; i++) v[i] = s[i].f1;
                  fold8<float>(v, lane);
                  tot.f1 = v[0];
                }
              }
              const uint32_t r = r0 + ((lane >> 2) & 7);
              if ((lane & 3) == 0 && r < nr)
                s_cand_keys[j0 + r] = make_key(finish_distance<ACC>(a.dtype, tot, qn), s_cand_ids[j0 + r]);
            }
          } else if (G == 32) {
            // long rows: one row at a time (query chunks in registers, or in shared memory when CPL == 0)
            for (uint32_t r = 0; r < nr; r++) {
              const uint4 *rp = reinterpret_cast<const uint4 *>(wstage + (size_t)r * srow_bytes);
              Sums s = zero_sums();
              if (CPL > 0) {
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                  const uint32_t ch = lane + c * 32;
                  if (ch < a.chunks) acc_chunk<ACC>(s, qreg[c], rp[ch]);
                }
              } else {
                for (uint32_t c = lane; c < a.chunks; c += 32) acc_chunk<ACC>(s, s_query_row[c], rp[c]);
              }
              group_fold<ACC, 32>(s);
              if (lane == 0) s_cand_keys[j0 + r] = make_key(finish_distance<ACC>(a.dtype, s, qn), s_cand_ids[j0 + r]);
            }
          } else {
            // short rows: R rows per warp instruction, G lanes each
            for (uint32_t r0 = 0; r0 < nr; r0 += R) {
              const uint32_t r = r0 + grp;
              Sums s = zero_sums();
              if (r < nr && (uint32_t)gl < a.chunks)
                acc_chunk<ACC>(s, qreg[0], reinterpret_cast<const uint4 *>(wstage + (size_t)r * a.row_bytes)[gl]);
              group_fold<ACC, G>(s);
              if (gl == 0 && r < nr) s_cand_keys[j0 + r] = make_key(finish_distance<ACC>(a.dtype, s, qn), s_cand_ids[j0 + r]);
            }
          }
          __syncwarp();
        }
        if (WS == 0 && pending_mask) {   // warps without rows this round
#pragma unroll
          for (int i = 0; i < SEARCH_EPT; i++)
            if (pending_mask & (1u << i)) hash_insert(hash, a.hash_bits - 3, pending_id[i], bp[i]);
        }
        if (warp == 0) { PROF_MARK(5) }
      }
      __syncthreads();  // (C) keys are published; the staging area may be overwritten
      if (warp == 0) { PROF_MARK(6) }
    }

    // ---- write the outcome
    const int state = s_state;
#ifdef SEARCH_PHASE_PROFILE
    if (a.prof && tid == 0 && state == 1) {
      for (int i = 0; i < 8; i++) a.prof[(size_t)q * 8 + i] = pf[i];
    }
#endif
    if (warp == 0) {
      if (state == 1) {
        for (uint32_t i = lane; i < a.k; i += 32) {
          uint64_t key = KEY_NONE;
          if (a.k <= 32) key = res.reg;
          else if (i < res.n) key = res.smem[i];
          bool ok = i < res.n;
          if (a.keys_out) {
            a.keys_out[(size_t)q * a.k + i] = ok ? key + a.id_offset : KEY_NONE;
          } else {
            a.ids[(size_t)q * a.k + i] = ok ? key_id(key) : 0u;
            a.dists[(size_t)q * a.k + i] = ok ? key_dist(key) : 0.f;
          }
        }
        if (lane == 0) {
          a.counts[q] = res.n;
          if (a.stats) {
            a.stats[(size_t)q * 3 + 0] = st_dist;
            a.stats[(size_t)q * 3 + 1] = st_edge;
            a.stats[(size_t)q * 3 + 2] = st_exp;
          }
        }
      } else if (lane == 0) {
        a.counts[q] = 0xffffffffu;
        if (WS == 0) {
          uint32_t slot = atomicAdd(a.overflow_count, 1u);
          a.overflow_list[slot] = q;
        } else {
          atomicAdd(a.failed_count, 1u);
        }
      }
    }
    __syncthreads();  // s_query / s_state are rewritten by thread 0 next
  }
}

// ---- launch plumbing: one translation unit per accumulate kind instantiates its kernels -------------
struct SearchLaunch {
  int group;       // G
  int cpl;         // chunks per lane kept in registers (0: query in shared memory)
  int ws;          // 0 shared-memory tier, 1 HBM tier
  unsigned grid;
  size_t smem;
  cudaStream_t stream;
};

template <int ACC, int G, int CPL, int WS>
static cudaError_t launch_one(const SearchArgs &a, const SearchLaunch &l) {
  cudaError_t e = cudaFuncSetAttribute(search_kernel<ACC, G, CPL, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)l.smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(search_kernel<ACC, G, CPL, WS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return e;
  search_kernel<ACC, G, CPL, WS><<<l.grid, SEARCH_THREADS, l.smem, l.stream>>>(a);
  return cudaGetLastError();
}

template <int ACC, int G, int CPL, int WS>
static cudaError_t occupancy_one(size_t smem, int *blocks) {
  cudaError_t e = cudaFuncSetAttribute(search_kernel<ACC, G, CPL, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(search_kernel<ACC, G, CPL, WS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, search_kernel<ACC, G, CPL, WS>, SEARCH_THREADS, smem);
}

// op == 0: launch, op == 1: occupancy query
template <int ACC>
cudaError_t search_dispatch(const SearchArgs &a, const SearchLaunch &l, int op, int *blocks) {
#define SEARCH_CASE(GG, CC, WW)                                   \
  if (l.group == GG && l.cpl == CC && l.ws == WW)                 \
    return op == 0 ? launch_one<ACC, GG, CC, WW>(a, l) : occupancy_one<ACC, GG, CC, WW>(l.smem, blocks);
#define SEARCH_CASES(WW) \
  SEARCH_CASE(1, 1, WW)  \
  SEARCH_CASE(2, 1, WW)  \
  SEARCH_CASE(4, 1, WW)  \
  SEARCH_CASE(8, 1, WW)  \
  SEARCH_CASE(16, 1, WW) \
  SEARCH_CASE(32, 1, WW) \
  SEARCH_CASE(32, 2, WW) \
  SEARCH_CASE(32, 4, WW) \
  SEARCH_CASE(32, 8, WW) \
  SEARCH_CASE(32, 0, WW)
  SEARCH_CASES(0)
  SEARCH_CASES(1)
#undef SEARCH_CASES
#undef SEARCH_CASE
  return cudaErrorInvalidValue;
}

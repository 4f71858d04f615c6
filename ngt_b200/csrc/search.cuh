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
// list matters (results are gated by the shrinking explorationRadius), so the merge falls back to the
// reference's element-by-element order.
//
// Work split inside the CTA (4 warps): warp 0 is the control warp (queue pop, adjacency read, visited
// filter, merge); all warps gather the surviving neighbours' rows from HBM with 128-bit loads, G lanes
// per row (G = min(32, pow2ceil(row_bytes/16))), several rows in flight per lane.
//
// Tiers: WS == 0 keeps the visited set (open-addressing hash) and the queue in shared memory; a query
// that outgrows them is appended to an overflow list and re-run by the WS == 1 instantiation, which
// keeps an exact bitmap and a large queue in HBM (the reference's own choice, Graph.h:751-799).
#pragma once
#include "ngtgpu_internal.cuh"

#define SEARCH_WARPS 4
#define SEARCH_THREADS (SEARCH_WARPS * 32)
#define SEARCH_CMAX 128  // adjacency entries filtered per round

struct SearchArgs {
  const uint8_t *objects;
  uint32_t row_bytes;
  uint32_t chunks;
  uint64_t n;
  const uint64_t *row_ptr;
  const uint32_t *col;
  const uint8_t *queries;  // prepared rows, nq x row_bytes
  const uint32_t *seeds;   // nq x n_seeds
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
  uint32_t *ids;
  float *dists;
  uint32_t *counts;
  uint32_t *stats;  // nullable, nq x 3
  uint32_t *work_counter;
  const uint32_t *query_list;        // WS == 1: the overflow list of the first tier
  const uint32_t *query_list_count;  // WS == 1
  uint32_t *overflow_list;           // WS == 0: where overflowing queries go
  uint32_t *overflow_count;
  uint32_t *failed_count;            // WS == 1: queries that outgrew even the HBM tier
  uint32_t *big_bitmaps;             // WS == 1: gridDim.x x bitmap_words
  uint64_t *big_queues;              // WS == 1: gridDim.x x queue_cap
  uint64_t bitmap_words;
};

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
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

template <int ACC, int G, int CPL, int WS>
__global__ void __launch_bounds__(SEARCH_THREADS) search_kernel(const SearchArgs a) {
  constexpr int R = 32 / G;                                  // rows per warp instruction
  constexpr int NCH = CPL > 0 ? CPL : 1;                      // register-resident query chunks per lane
  constexpr int U = CPL == 0 ? 2 : (CPL >= 8 ? 2 : (CPL >= 4 ? 2 : (CPL == 2 ? 4 : 8)));  // row groups in flight per warp
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_cand_ids[SEARCH_CMAX];
  __shared__ uint64_t s_cand_keys[SEARCH_CMAX];
  __shared__ uint32_t s_cand_n;
  __shared__ int s_state;  // 0 run, 1 finished, 2 overflow
  __shared__ uint32_t s_query;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int gl = lane % G;   // lane inside its row group
  const int grp = lane / G;  // row group inside the warp

  // ---- carve dynamic shared memory: [results k>32][queue][hash][query copy (CPL==0)]
  uint8_t *sp = smem_raw;
  uint64_t *s_results = reinterpret_cast<uint64_t *>(sp);
  sp += a.k > 32 ? (((size_t)a.k * 8 + 15) & ~(size_t)15) : 0;
  uint64_t *queue;
  uint32_t *hash = nullptr;
  uint32_t *bitmap = nullptr;
  if (WS == 0) {
    queue = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)a.queue_cap * 8;
    hash = reinterpret_cast<uint32_t *>(sp);
    sp += (size_t)4 << a.hash_bits;
  } else {
    queue = a.big_queues + (size_t)blockIdx.x * a.queue_cap;
    bitmap = a.big_bitmaps + (size_t)blockIdx.x * a.bitmap_words;
  }
  const uint4 *s_query_row = reinterpret_cast<const uint4 *>(sp);  // CPL == 0 only
  const uint32_t hash_mask = (1u << a.hash_bits) - 1u;

  for (;;) {
    // ---- next query (dynamic scheduling over a persistent grid)
    if (tid == 0) {
      uint32_t w = atomicAdd(a.work_counter, 1u);
      uint32_t total = WS == 0 ? a.nq : *a.query_list_count;
      s_query = w < total ? (WS == 0 ? w : a.query_list[w]) : 0xffffffffu;
      s_state = 0;
      s_cand_n = 0;
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
    float qn = 0.f;
    if (CPL > 0) {
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
      if (CPL > 0) {
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
    uint32_t cand_n = 0;

    for (;;) {
      if (warp == 0) {
        // ======== merge the keys evaluated in the previous round ========
        if (cand_n) {
          if (a.coef >= 1.0f && !seeding) {
            // set semantics: results first, then everything within the final explorationRadius
            for (uint32_t j0 = 0; j0 < cand_n; j0 += 32) {
              uint64_t key = j0 + lane < cand_n ? s_cand_keys[j0 + lane] : KEY_NONE;
              float d = key_dist(key);
              uint32_t m = __ballot_sync(0xffffffffu, key != KEY_NONE && d <= radius);
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
            for (uint32_t j0 = 0; j0 < cand_n; j0 += 32) {
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
                if (lane == 0) s_state = 2;
                break;
              }
              if (acc) queue[qsize + __popc(m & lanemask_lt())] = key;
              qsize += cnt;
            }
            __syncwarp();
          } else {
            // the reference's order: seeds (all go to unchecked, Graph.cpp:352-366) and coef < 1
            bool full = false;
            for (uint32_t j0 = 0; j0 < cand_n && !full; j0 += 32) {
              uint64_t key = j0 + lane < cand_n ? s_cand_keys[j0 + lane] : KEY_NONE;
              uint32_t m = __ballot_sync(0xffffffffu, key != KEY_NONE);
              while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                uint64_t kk = shfl_u64(key, src);
                float d = key_dist(kk);
                if (!seeding && d > er) continue;
                if (qsize >= a.queue_cap) {
                  if (lane == 0) s_state = 2;
                  full = true;
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
        // ======== pop / filter until some unvisited neighbours are found ========
        bool overflow = false;
        __syncwarp();
        if (*(volatile int *)&s_state == 2) overflow = true;
        while (!overflow && cand_n == 0) {
          if (cur_pos >= cur_deg) {
            if (seeding) {
              // setupSeeds: radius from the seeds once k of them are within it (Graph.cpp:349-351)
              seeding = false;
              if (res.n >= res.k) radius = key_dist(result_kth(res));
              er = a.coef * radius;
            }
            // pop the smallest unchecked key
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
              if (lane == 0) s_state = 1;
              break;
            }
            if (lane == 0) queue[bi] = queue[qsize - 1];
            qsize--;
            __syncwarp();
            uint32_t t = key_id(best);
            uint64_t b = 0, e = 0;
            if (lane < 2) b = a.row_ptr[(size_t)t + lane];
            e = shfl_u64(b, 1);
            b = shfl_u64(b, 0);
            uint64_t deg = e - b;
            if (deg > a.edge_cap) deg = a.edge_cap;  // Graph.cpp:438
            cur = a.col + b;
            cur_deg = (uint32_t)deg;
            cur_pos = 0;
            st_edge += cur_deg;
            st_exp++;
            if (cur_deg == 0) continue;
          }
          uint32_t take = cur_deg - cur_pos;
          if (take > SEARCH_CMAX) take = SEARCH_CMAX;
          if (WS == 0 && visited_n + take > a.hash_limit) {
            if (lane == 0) s_state = 2;
            overflow = true;
            break;
          }
          for (uint32_t e0 = 0; e0 < take; e0 += 32) {
            uint32_t ei = e0 + lane;
            uint32_t nid = ei < take ? __ldg(cur + cur_pos + ei) : 0u;
            bool isnew = false;
            if (nid != 0u && nid <= a.n) {
              if (WS == 0) {
                uint32_t slot = (nid * 2654435761u) >> (32 - a.hash_bits);
                for (;;) {
                  uint32_t old = atomicCAS(&hash[slot], 0u, nid);
                  if (old == 0u) {
                    isnew = true;
                    break;
                  }
                  if (old == nid) break;
                  slot = (slot + 1) & hash_mask;
                }
              } else {
                uint32_t bit = 1u << (nid & 31);
                uint32_t old = atomicOr(&bitmap[nid >> 5], bit);
                isnew = (old & bit) == 0;
              }
            }
            uint32_t m = __ballot_sync(0xffffffffu, isnew);
            if (isnew) s_cand_ids[cand_n + __popc(m & lanemask_lt())] = nid;
            cand_n += __popc(m);
          }
          cur_pos += take;
          visited_n += cand_n;
        }
        st_dist += cand_n;
        if (lane == 0) s_cand_n = cand_n;
      }
      __syncthreads();  // candidate list (or the final state) is published
      if (s_state != 0) break;
      const uint32_t cn = s_cand_n;

      // ======== gather: all warps evaluate the candidates' distances ========
      {
        const uint32_t rounds = (cn + SEARCH_WARPS * R * U - 1) / (SEARCH_WARPS * R * U);
        for (uint32_t it = 0; it < rounds; it++) {
          if (CPL > 0) {
            uint4 rows[U][NCH];
            uint32_t cid[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
              uint32_t j = (it * U + u) * (SEARCH_WARPS * R) + warp * R + grp;
              cid[u] = j < cn ? s_cand_ids[j] : 0u;
              const uint8_t *rp = a.objects + (size_t)cid[u] * a.row_bytes;
#pragma unroll
              for (int i = 0; i < NCH; i++) {
                uint32_t c = gl + i * G;
                rows[u][i] = (j < cn && c < a.chunks) ? ldg16_stream(rp + (size_t)c * 16) : zero16();
              }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
              uint32_t j = (it * U + u) * (SEARCH_WARPS * R) + warp * R + grp;
              Sums s = zero_sums();
#pragma unroll
              for (int i = 0; i < NCH; i++) acc_chunk<ACC>(s, qreg[i], rows[u][i]);
              group_fold<ACC, G>(s);
              if (gl == 0 && j < cn) s_cand_keys[j] = make_key(finish_distance<ACC>(a.dtype, s, qn), cid[u]);
            }
          } else {
            // long rows: the query sits in shared memory, G == 32, one row per warp at a time
#pragma unroll
            for (int u = 0; u < U; u++) {
              uint32_t j = (it * U + u) * (SEARCH_WARPS * R) + warp * R + grp;
              uint32_t id = j < cn ? s_cand_ids[j] : 0u;
              const uint8_t *rp = a.objects + (size_t)id * a.row_bytes;
              Sums s = zero_sums();
              if (j < cn) {
                uint32_t c = gl;
                for (; c + 3 * G < a.chunks; c += 4 * G) {
                  uint4 r0 = ldg16_stream(rp + (size_t)c * 16);
                  uint4 r1 = ldg16_stream(rp + (size_t)(c + G) * 16);
                  uint4 r2 = ldg16_stream(rp + (size_t)(c + 2 * G) * 16);
                  uint4 r3 = ldg16_stream(rp + (size_t)(c + 3 * G) * 16);
                  acc_chunk<ACC>(s, s_query_row[c], r0);
                  acc_chunk<ACC>(s, s_query_row[c + G], r1);
                  acc_chunk<ACC>(s, s_query_row[c + 2 * G], r2);
                  acc_chunk<ACC>(s, s_query_row[c + 3 * G], r3);
                }
                for (; c < a.chunks; c += G) acc_chunk<ACC>(s, s_query_row[c], ldg16_stream(rp + (size_t)c * 16));
              }
              group_fold<ACC, G>(s);
              if (gl == 0 && j < cn) s_cand_keys[j] = make_key(finish_distance<ACC>(a.dtype, s, qn), id);
            }
          }
        }
      }
      __syncthreads();  // keys are published
    }

    // ---- write the outcome
    const int state = s_state;
    if (warp == 0) {
      if (state == 1) {
        for (uint32_t i = lane; i < a.k; i += 32) {
          uint64_t key = KEY_NONE;
          if (a.k <= 32) key = res.reg;
          else if (i < res.n) key = res.smem[i];
          bool ok = i < res.n;
          a.ids[(size_t)q * a.k + i] = ok ? key_id(key) : 0u;
          a.dists[(size_t)q * a.k + i] = ok ? key_dist(key) : 0.f;
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
  search_kernel<ACC, G, CPL, WS><<<l.grid, SEARCH_THREADS, l.smem, l.stream>>>(a);
  return cudaGetLastError();
}

template <int ACC, int G, int CPL, int WS>
static cudaError_t occupancy_one(size_t smem, int *blocks) {
  cudaError_t e = cudaFuncSetAttribute(search_kernel<ACC, G, CPL, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, search_kernel<ACC, G, CPL, WS>, SEARCH_THREADS, smem);
}

// op == 0: launch, op == 1: occupancy query
template <int ACC>
cudaError_t search_dispatch(const SearchArgs &a, const SearchLaunch &l, int op, int *blocks) {
#define SEARCH_CASE(GG, CC, WW)                                   \
  if (l.group == GG && l.cpl == CC && l.ws == WW)                 \
    return op == 0 ? launch_one<ACC, GG, CC, WW>(a, l) : occupancy_one<ACC, GG, CC, WW>(l.smem, blocks);
  SEARCH_CASE(1, 1, 0)
  SEARCH_CASE(2, 1, 0)
  SEARCH_CASE(4, 1, 0)
  SEARCH_CASE(8, 1, 0)
  SEARCH_CASE(16, 1, 0)
  SEARCH_CASE(32, 1, 0)
  SEARCH_CASE(32, 2, 0)
  SEARCH_CASE(32, 4, 0)
  SEARCH_CASE(32, 8, 0)
  SEARCH_CASE(32, 0, 0)
  SEARCH_CASE(32, 0, 1)
#undef SEARCH_CASE
  return cudaErrorInvalidValue;
}

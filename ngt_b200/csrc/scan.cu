// scan.cu -- exhaustive k-nearest scan for a batch of queries (CUDA-core tile kernel + merge).
//
// Replaces ObjectSpaceRepository::linearSearch (lib/NGT/ObjectSpaceRepository.h:466-502): every
// non-empty object is compared with the query, the k smallest (distance,id) within `radius` are kept,
// output ascending. The reference does this one query at a time, streaming the whole repository per
// query; here a CTA owns a tile of 64 queries and walks the object rows in tiles of 64, so one pass
// over HBM serves 64 queries (and all CTAs walk in the same order, so the pass is shared through L2).
//
// Exactness. Integer kinds (uint8 L2, Hamming) are computed exactly in the tile (dp4a / popc) and
// finished as the reference does. Float kinds are accumulated in the tile in plain fp32, which is only a
// FILTER: a pair whose approximate value is within a safety margin of the query's current k-th distance
// is re-evaluated with the engine's one summation order (ngtgpu_internal.cuh: group_fold) before it may
// enter the result list, so the distances reported here are bit-identical to the ones the graph search
// reports for the same pair. The margin (1e-3 relative for sums of squares, 1e-3..2e-3 absolute in cosine
// space) is two orders above the fp32 accumulation error bound dim * 2^-24 for dim <= 16384.
#include <cfloat>
#include <cstring>

#include <cstdlib>

#include "ngtgpu_internal.cuh"

#define SCAN_TQ 64
#define SCAN_TR 64
#define SCAN_KC 8          // 16-byte chunks per pipeline stage
#define SCAN_LD (SCAN_KC + 1)
#define SCAN_THREADS 256
#define SCAN_MAX_SPLIT 32

struct ScanArgs {
  const uint8_t *queries;
  uint32_t nq;
  const uint8_t *rows;
  uint64_t n_rows;
  uint32_t first_row_id;
  const uint8_t *valid;
  const uint32_t *id_map;
  uint32_t row_bytes;
  uint32_t chunks;
  uint32_t k;
  float radius;
  int exclude_self;
  uint32_t self_base;
  int dtype;
  uint32_t qtiles;
  uint32_t nsplit;
  uint64_t tiles_per_split;
  uint64_t *partial;        // nq x nsplit x k keys
  int approx;               // float kinds: rank by the tile's own fp32 sums (seed selection), no exact re-evaluation
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// threshold of the tile filter in the accumulate domain, from the current k-th distance (or the radius)
template <int ACC>
__device__ __forceinline__ uint32_t raw_threshold(int dtype, float thr_d) {
  if (ACC == ACC_U8_HAM) {
    if (!(thr_d < 4.0e9f)) return 0xffffffffu;
    return thr_d < 0.f ? 0u : (uint32_t)thr_d;
  }
  if (ACC == ACC_U8_L2) {
    if (!(thr_d < 60000.0f)) return 0xffffffffu;
    if (thr_d < 0.f) return 0u;
    float t = thr_d * thr_d * 1.000001f + 1.0f;
    return (uint32_t)t;
  }
  if (ACC == ACC_F_L2) {
    if (!(thr_d < 1.0e18f)) return __float_as_uint(__int_as_float(0x7f800000));
    float t = thr_d * 1.001f + 1e-30f;
    return __float_as_uint(t * t);
  }
  // similarity domains: pass when the similarity is >= threshold
  const float ninf = __int_as_float(0xff800000);
  if (!(thr_d < 1.0e18f)) return __float_as_uint(ninf);
  float t;
  if (ACC == ACC_F_DOT) {
    if (dtype == NGTGPU_DISTANCE_NORMALIZED_L2) t = 1.0f - 0.5f * thr_d * thr_d - 1e-3f;
    else if (dtype == NGTGPU_DISTANCE_NORMALIZED_COSINE) t = 1.0f - thr_d - 1e-3f;
    else t = (thr_d >= 3.1415927f ? -1.0f : cosf(thr_d)) - 1e-3f;
  } else {
    if (dtype == NGTGPU_DISTANCE_COSINE) t = 1.0f - thr_d - 2e-3f;
    else t = (thr_d >= 3.1415927f ? -1.0f : cosf(thr_d)) - 2e-3f;
  }
  return __float_as_uint(t);
}

template <int ACC>
__device__ __forceinline__ bool tile_pass(const Sums &s, uint32_t thr_raw, float qn) {
  if (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM) return s.u <= thr_raw;
  float t = __uint_as_float(thr_raw);
  if (ACC == ACC_F_L2) return s.f0 <= t;
  if (ACC == ACC_F_DOT) return s.f0 >= t;
  float c = s.f0 * rsqrtf(qn * s.f1);
  return !(c < t);  // NaN (zero vectors) goes to the exact path
}

// warp-parallel insertion into an ascending array of at most k keys
__device__ __forceinline__ void sorted_insert(uint64_t *arr, uint32_t &n_io, uint32_t k, uint64_t key, int lane) {
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

template <int ACC, int G>
__global__ void __launch_bounds__(SCAN_THREADS, 2) scan_tile_kernel(const ScanArgs a) {
  constexpr bool IS_INT = ACC == ACC_U8_L2 || ACC == ACC_U8_HAM;
  constexpr int RG = IS_INT ? 1 : G;  // lanes per re-evaluated candidate
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4 *stages = reinterpret_cast<uint4 *>(smem_raw);                                  // 2 x (Q | R) x 64 x LD
  uint64_t *buf = reinterpret_cast<uint64_t *>(smem_raw + 2 * 2 * SCAN_TQ * SCAN_LD * 16);  // TQ x TR
  uint64_t *topk = buf + SCAN_TQ * SCAN_TR;                                             // TQ x k
  uint32_t *thr_raw = reinterpret_cast<uint32_t *>(topk + (size_t)SCAN_TQ * a.k);      // TQ
  uint32_t *cnt = thr_raw + SCAN_TQ;                                                    // TQ
  uint32_t *tn = cnt + SCAN_TQ;                                                         // TQ
  float *qnorm = reinterpret_cast<float *>(tn + SCAN_TQ);                               // TQ

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int tx = tid & 15;
  const int ty = tid >> 4;
  const uint32_t qtile = blockIdx.x % a.qtiles;
  const uint32_t split = blockIdx.x / a.qtiles;
  const uint32_t q0 = qtile * SCAN_TQ;
  const uint64_t total_tiles = (a.n_rows + SCAN_TR - 1) / SCAN_TR;
  const uint64_t tile_begin = (uint64_t)split * a.tiles_per_split;
  uint64_t tile_end = tile_begin + a.tiles_per_split;
  if (tile_end > total_tiles) tile_end = total_tiles;
  const uint32_t nkc = (a.chunks + SCAN_KC - 1) / SCAN_KC;
  const float thr0 = a.radius < 0.f ? FLT_MAX : a.radius;

  // ---- init per-query state
  if (tid < SCAN_TQ) {
    thr_raw[tid] = raw_threshold<ACC>(a.dtype, thr0);
    cnt[tid] = 0;
    tn[tid] = 0;
    qnorm[tid] = 1.f;
  }
  if (ACC == ACC_F_COS) {
    for (int m = warp; m < SCAN_TQ; m += SCAN_THREADS / 32) {
      float s = 0.f;
      if (q0 + m < a.nq) {
        const float *qr = reinterpret_cast<const float *>(a.queries + (size_t)(q0 + m) * a.row_bytes);
        for (uint32_t i = lane; i < a.row_bytes / 4; i += 32) s = fmaf(qr[i], qr[i], s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      __syncwarp();
      if (lane == 0) qnorm[m] = s;
    }
  }
  __syncthreads();

  auto load_stage = [&](int sbuf, uint64_t tile, uint32_t kc) {
    uint4 *st = stages + (size_t)sbuf * (2 * SCAN_TQ * SCAN_LD);
    const uint64_t row0 = tile * SCAN_TR;
#pragma unroll
    for (int x = tid; x < 2 * SCAN_TQ * SCAN_KC; x += SCAN_THREADS) {
      const bool is_r = x >= SCAN_TQ * SCAN_KC;
      const int idx = x & (SCAN_TQ * SCAN_KC - 1);
      const int r = idx / SCAN_KC, c = idx % SCAN_KC;
      const uint32_t chunk = kc * SCAN_KC + c;
      const uint8_t *src;
      bool ok = chunk < a.chunks;
      if (is_r) {
        uint64_t row = row0 + r;
        ok = ok && row < a.n_rows;
        src = a.rows + (ok ? row : 0) * a.row_bytes + (size_t)(ok ? chunk : 0) * 16;
      } else {
        uint32_t qq = q0 + r;
        ok = ok && qq < a.nq;
        src = a.queries + (size_t)(ok ? qq : 0) * a.row_bytes + (size_t)(ok ? chunk : 0) * 16;
      }
      cp_async16(st + (is_r ? SCAN_TQ * SCAN_LD : 0) + r * SCAN_LD + c, src, ok ? 16u : 0u);
    }
    cp_async_commit();
  };

  const uint64_t ntiles = tile_end > tile_begin ? tile_end - tile_begin : 0;
  const uint64_t iters = ntiles * nkc;
  Sums acc[4][4];
  if (iters) load_stage(0, tile_begin, 0);
  for (uint64_t it = 0; it < iters; it++) {
    const uint64_t tile = tile_begin + it / nkc;
    const uint32_t kc = (uint32_t)(it % nkc);
    if (it + 1 < iters) {
      load_stage((int)((it + 1) & 1), tile_begin + (it + 1) / nkc, (uint32_t)((it + 1) % nkc));
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kc == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = zero_sums();
    }
    {
      const uint4 *qs = stages + (size_t)(it & 1) * (2 * SCAN_TQ * SCAN_LD);
      const uint4 *rs = qs + SCAN_TQ * SCAN_LD;
      uint32_t kcn = a.chunks - kc * SCAN_KC;
      if (kcn > SCAN_KC) kcn = SCAN_KC;
      for (uint32_t c = 0; c < kcn; c++) {
        uint4 qv[4], rv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) qv[i] = qs[(ty + 16 * i) * SCAN_LD + c];
#pragma unroll
        for (int j = 0; j < 4; j++) rv[j] = rs[(tx + 16 * j) * SCAN_LD + c];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc_chunk_seq<ACC>(acc[i][j], qv[i], rv[j]);
      }
    }
    if (kc == nkc - 1) {
      // ---- tile epilogue: filter against each query's threshold
      const uint64_t row0 = tile * SCAN_TR;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int m = ty + 16 * i;
        const uint32_t thr = thr_raw[m];
        const float qn = qnorm[m];
        const uint32_t qq = q0 + m;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (tile_pass<ACC>(acc[i][j], thr, qn)) {
            const int nn = tx + 16 * j;
            const uint64_t row = row0 + nn;
            if (qq < a.nq && row < a.n_rows) {
              const uint32_t id = a.id_map ? a.id_map[row] : a.first_row_id + (uint32_t)row;
              bool ok = !(a.valid && a.valid[id] == 0);
              if (a.exclude_self && id == a.self_base + qq) ok = false;
              if (ok) {
                uint32_t slot = atomicAdd(&cnt[m], 1u);
                uint32_t raw = IS_INT ? acc[i][j].u : __float_as_uint(acc[i][j].f0);
                if (ACC == ACC_F_COS && a.approx) raw = __float_as_uint(acc[i][j].f0 * rsqrtf(qn * acc[i][j].f1));
                buf[m * SCAN_TR + slot] = ((uint64_t)raw << 32) | (uint32_t)nn;
              }
            }
          }
        }
      }
      __syncthreads();
      // ---- exact evaluation + insertion, one warp per query
      for (int m = warp; m < SCAN_TQ; m += SCAN_THREADS / 32) {
        const uint32_t c = cnt[m];
        if (c == 0) continue;
        const int rg = (IS_INT || a.approx) ? 1 : RG;   // lanes per candidate
        const int rr = 32 / rg;
        const int gl = lane % rg, grp = lane / rg;
        uint64_t *mytop = topk + (size_t)m * a.k;
        uint32_t n = tn[m];
        const uint8_t *qptr = a.queries + (size_t)(q0 + m) * a.row_bytes;
        for (uint32_t i0 = 0; i0 < c; i0 += rr) {
          const uint32_t i = i0 + grp;
          const bool act = i < c;
          const uint64_t entry = act ? buf[m * SCAN_TR + i] : 0ull;
          const uint64_t row = row0 + (uint32_t)(entry & 63u);
          uint32_t id = 0;
          if (act) id = a.id_map ? a.id_map[row] : a.first_row_id + (uint32_t)row;
          float d;
          if (IS_INT) {
            Sums s = zero_sums();
            s.u = (uint32_t)(entry >> 32);
            d = finish_distance<ACC>(a.dtype, s, 0.f);
          } else if (a.approx) {
            // seed selection: the tile's fp32 value is good enough to rank pivots
            Sums s = zero_sums();
            s.f0 = __uint_as_float((uint32_t)(entry >> 32));
            s.f1 = 1.0f;   // F_COS: f0 already holds the cosine
            d = finish_distance<ACC>(a.dtype, s, 1.0f);
          } else {
            d = group_distance_gmem<ACC, RG>(qptr, act ? a.rows + row * a.row_bytes : qptr, a.chunks, gl, a.dtype);
          }
          uint64_t key = KEY_NONE;
          if (act && gl == 0 && (a.radius < 0.f || d <= a.radius)) key = make_key(d, id);
          uint32_t mm = __ballot_sync(0xffffffffu, key != KEY_NONE);
          while (mm) {
            int src = __ffs(mm) - 1;
            mm &= mm - 1;
            uint64_t kk = shfl_u64(key, src);
            sorted_insert(mytop, n, a.k, kk, lane);
          }
        }
        if (lane == 0) {
          tn[m] = n;
          cnt[m] = 0;
          if (n == a.k) thr_raw[m] = raw_threshold<ACC>(a.dtype, key_dist(mytop[a.k - 1]));
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }

  // ---- partial result of this (query tile, split)
  for (int m = warp; m < SCAN_TQ; m += SCAN_THREADS / 32) {
    const uint32_t qq = q0 + m;
    if (qq >= a.nq) continue;
    const uint32_t n = tn[m];
    uint64_t *out = a.partial + ((size_t)qq * a.nsplit + split) * a.k;
    for (uint32_t i = lane; i < a.k; i += 32) out[i] = i < n ? topk[(size_t)m * a.k + i] : KEY_NONE;
  }
}

// one warp per query: merge nsplit ascending lists (<= 32 of them) into the final top-k
__global__ void scan_merge_kernel(const uint64_t *__restrict__ partial, uint32_t nq, uint32_t nsplit, uint32_t k,
                                  uint64_t query_stride, uint64_t list_stride, uint32_t *__restrict__ ids,
                                  float *__restrict__ dists, uint32_t *__restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  const uint64_t *mine = partial + (size_t)q * query_stride + (size_t)(lane < (int)nsplit ? lane : 0) * list_stride;
  uint32_t pos = 0;
  uint64_t head = lane < (int)nsplit ? mine[0] : KEY_NONE;
  uint32_t n = 0;
  for (uint32_t i = 0; i < k; i++) {
    uint64_t best = head;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      uint64_t ob = shfl_xor_u64(best, o);
      if (ob < best) best = ob;
    }
    if (best == KEY_NONE) break;
    if (head == best) {  // keys are unique (ids are), so exactly one lane advances
      pos++;
      head = pos < k ? mine[pos] : KEY_NONE;
    }
    if (lane == 0) {
      ids[(size_t)q * k + i] = key_id(best);
      dists[(size_t)q * k + i] = key_dist(best);
    }
    n++;
  }
  for (uint32_t i = n + lane; i < k; i += 32) {
    ids[(size_t)q * k + i] = 0;
    dists[(size_t)q * k + i] = 0.f;
  }
  if (lane == 0) counts[q] = n;
}

template <int ACC, int G>
static cudaError_t launch_scan(const ScanArgs &a, unsigned grid, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(scan_tile_kernel<ACC, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan_tile_kernel<ACC, G><<<grid, SCAN_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int ACC>
static cudaError_t launch_scan_g(int group, const ScanArgs &a, unsigned grid, size_t smem, cudaStream_t stream) {
  switch (group) {
    case 4: return launch_scan<ACC, 4>(a, grid, smem, stream);
    case 8: return launch_scan<ACC, 8>(a, grid, smem, stream);
    case 16: return launch_scan<ACC, 16>(a, grid, smem, stream);
    case 32: return launch_scan<ACC, 32>(a, grid, smem, stream);
  }
  return cudaErrorInvalidValue;
}

int ngtgpu_scan_topk(ngtgpu_index *ix, const ScanParams &p, cudaStream_t stream) {
  if (p.nq == 0) return NGTGPU_OK;
  if (p.k == 0) {
    CUDA_TRY(cudaMemsetAsync(p.d_counts, 0, (size_t)p.nq * 4, stream));
    return NGTGPU_OK;
  }
  {
    // float batches against the whole repository go to the tensor cores when the shape allows (knn_tc.cu)
    int used = 0;
    NGTGPU_TRY(ngtgpu_scan_topk_tc(ix, p, stream, &used));
    if (used) {
      ix->tc_batches++;
      return NGTGPU_OK;
    }
  }
  const size_t fixed = (size_t)2 * 2 * SCAN_TQ * SCAN_LD * 16 + (size_t)SCAN_TQ * SCAN_TR * 8 + 4 * SCAN_TQ * 4;
  const size_t smem = fixed + (size_t)SCAN_TQ * p.k * 8;
  if (smem > 220 * 1024)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "linear search: size " + std::to_string(p.k) + " exceeds the on-chip result lists (max " +
                                        std::to_string((220 * 1024 - fixed) / (SCAN_TQ * 8)) + ")");
  ScanArgs a;
  memset(&a, 0, sizeof(a));
  a.queries = p.d_queries;
  a.nq = p.nq;
  a.rows = p.d_rows;
  a.n_rows = p.n_rows;
  a.first_row_id = p.first_row_id;
  a.valid = p.d_valid;
  a.id_map = p.d_id_map;
  a.row_bytes = ix->row_bytes;
  a.chunks = ix->chunks;
  a.k = p.k;
  a.radius = p.radius;
  a.exclude_self = p.exclude_self;
  a.self_base = p.self_base;
  a.dtype = ix->distance_type;
  a.approx = p.approx;
  a.qtiles = (p.nq + SCAN_TQ - 1) / SCAN_TQ;
  const uint64_t total_tiles = (p.n_rows + SCAN_TR - 1) / SCAN_TR;
  // two CTAs per SM in flight; more row splits measured slower on the seed scan (1024 pivots: 2 -> 16 splits, 0.56 -> 1.0 ms)
  uint64_t want = ((uint64_t)2 * ix->sm_count + a.qtiles - 1) / a.qtiles;
  if (want > SCAN_MAX_SPLIT) want = SCAN_MAX_SPLIT;
  if (want > total_tiles) want = total_tiles;
  if (want < 1) want = 1;
  a.tiles_per_split = (total_tiles + want - 1) / want;
  if (a.tiles_per_split < 1) a.tiles_per_split = 1;
  a.nsplit = (uint32_t)((total_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
  if (a.nsplit < 1) a.nsplit = 1;
  uint64_t *partial = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_PARTIAL, (size_t)p.nq * a.nsplit * p.k * 8, (void **)&partial));
  a.partial = partial;
  const unsigned grid = a.qtiles * a.nsplit;
  cudaError_t e;
  switch (ix->acc_kind) {
    case ACC_F_L2: e = launch_scan_g<ACC_F_L2>((int)ix->group, a, grid, smem, stream); break;
    case ACC_F_DOT: e = launch_scan_g<ACC_F_DOT>((int)ix->group, a, grid, smem, stream); break;
    case ACC_F_COS: e = launch_scan_g<ACC_F_COS>((int)ix->group, a, grid, smem, stream); break;
    case ACC_U8_L2: e = launch_scan<ACC_U8_L2, 32>(a, grid, smem, stream); break;
    case ACC_U8_HAM: e = launch_scan<ACC_U8_HAM, 32>(a, grid, smem, stream); break;
    default: e = cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("scan kernel launch: ") + cudaGetErrorString(e));
  ix->launches++;
  const unsigned mblocks = (p.nq + 7) / 8;
  scan_merge_kernel<<<mblocks, 256, 0, stream>>>(partial, p.nq, a.nsplit, p.k, (uint64_t)a.nsplit * p.k, (uint64_t)p.k, p.d_ids,
                                                  p.d_dists, p.d_counts);
  ix->launches++;
  CUDA_TRY(cudaGetLastError());
  return NGTGPU_OK;
}

static int linear_common(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq, uint32_t size, float radius,
                         uint32_t *ids, float *dists, uint32_t *counts, bool on_device, cudaStream_t stream) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!ix->d_objects) NGTGPU_FAIL(NGTGPU_ERR_STATE, "linear search: the index holds no objects");
  if (nq == 0) return NGTGPU_OK;
  if (!queries || !counts || (size && (!ids || !dists))) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "linear search: null buffer");
  if (size == 0) {
    if (on_device) CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)nq * 4, stream));
    else memset(counts, 0, (size_t)nq * 4);
    return NGTGPU_OK;
  }
  uint8_t *d_q = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_QUERIES, (size_t)nq * ix->row_bytes, (void **)&d_q));
  NGTGPU_TRY(ngtgpu_prepare_queries(ix, queries, query_type, nq, on_device, d_q, stream));
  ScanParams p;
  p.d_queries = d_q;
  p.nq = nq;
  p.d_rows = ix->d_objects + ix->row_bytes;  // row of id 1
  p.n_rows = ix->n;
  p.first_row_id = 1;
  p.d_valid = ix->d_valid;
  p.k = size;
  p.radius = radius;
  if (on_device) {
    p.d_ids = ids;
    p.d_dists = dists;
    p.d_counts = counts;
    return ngtgpu_scan_topk(ix, p, stream);
  }
  uint32_t *io = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_IO, ((size_t)nq * size * 2 + nq) * 4, (void **)&io));
  p.d_ids = io;
  p.d_dists = reinterpret_cast<float *>(io + (size_t)nq * size);
  p.d_counts = io + (size_t)nq * size * 2;
  NGTGPU_TRY(ngtgpu_scan_topk(ix, p, stream));
  CUDA_TRY(cudaMemcpyAsync(ids, p.d_ids, (size_t)nq * size * 4, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaMemcpyAsync(dists, p.d_dists, (size_t)nq * size * 4, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaMemcpyAsync(counts, p.d_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return NGTGPU_OK;
}

extern "C" int ngtgpu_linear_search(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq, uint32_t size,
                                    float radius, uint32_t *ids, float *dists, uint32_t *counts) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ngtgpu_lane_guard lane(ix);
  NGTGPU_TRY(lane.status);
  return linear_common(ix, queries, query_type, nq, size, radius, ids, dists, counts, false, lane.stream());
}

extern "C" int ngtgpu_linear_search_device(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq,
                                           uint32_t size, float radius, uint32_t *ids, float *dists, uint32_t *counts,
                                           void *stream) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  return linear_common(ix, queries, query_type, nq, size, radius, ids, dists, counts, true, (cudaStream_t)stream);
}

// ---- exhaustive kNN of stored objects against the whole repository -------------------------------------
// The brute-force pass behind graph construction: GraphIndex::searchForKNNGInsertion (lib/NGT/Index.h:839-856)
// runs linearSearch with size k+1 for object `id` and drops the object itself (ObjectSpace.h:70-88); here the
// queries are the stored rows first_id .. first_id+count-1 and the object itself is skipped inside the scan.
// Outputs are device buffers [count x k]; enqueued on `stream`.
extern "C" int ngtgpu_index_knn_graph(ngtgpu_index *ix, uint32_t k, uint32_t first_id, uint32_t count, uint32_t *d_ids,
                                      float *d_dists, uint32_t *d_counts, void *stream) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!ix->d_objects) NGTGPU_FAIL(NGTGPU_ERR_STATE, "knn graph: the index holds no objects");
  if (count == 0) return NGTGPU_OK;
  if (first_id == 0 || (uint64_t)first_id + count - 1 > ix->n) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "knn graph: id range out of bounds");
  if (!d_ids || !d_dists || !d_counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "knn graph: null buffer");
  ScanParams p;
  p.d_queries = ix->d_objects + (size_t)first_id * ix->row_bytes;
  p.nq = count;
  p.d_rows = ix->d_objects + ix->row_bytes;
  p.n_rows = ix->n;
  p.first_row_id = 1;
  p.d_valid = ix->d_valid;
  p.k = k;
  p.radius = -1.0f;
  p.exclude_self = 1;
  p.self_base = first_id;
  p.d_ids = d_ids;
  p.d_dists = d_dists;
  p.d_counts = d_counts;
  return ngtgpu_scan_topk(ix, p, (cudaStream_t)stream);
}

// ---- merging per-shard result lists (multi-GPU) ----------------------------------
// Each GPU searches its own shard of the rows; the per-shard top-k lists are exchanged with one all-gather
// and merged per query by (distance, id) -- ObjectDistance's order (lib/NGT/Common.h:1946-1952) -- exactly as
// one priority queue over the union would keep them. Lists travel as 64-bit keys (ordered distance bits in
// the high word, GLOBAL id in the low word) so the exchange is one flat buffer.
__global__ void pack_keys_kernel(const uint32_t *__restrict__ ids, const float *__restrict__ dists,
                                 const uint32_t *__restrict__ counts, uint32_t nq, uint32_t k, uint32_t id_offset,
                                 uint64_t *__restrict__ keys) {
  const uint64_t total = (uint64_t)nq * k;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t q = (uint32_t)(i / k), r = (uint32_t)(i % k);
    uint32_t c = counts[q];
    if (c == 0xffffffffu) c = 0;   // a query that outgrew every tier of the device-pointer search: no results, not k zeros
    keys[i] = r < c ? make_key(dists[i], ids[i] + id_offset) : KEY_NONE;
  }
}

extern "C" int ngtgpu_pack_keys(const uint32_t *d_ids, const float *d_dists, const uint32_t *d_counts, uint32_t nq,
                                uint32_t k, uint32_t id_offset, uint64_t *d_keys, void *stream) {
  if (nq == 0 || k == 0) return NGTGPU_OK;
  if (!d_ids || !d_dists || !d_counts || !d_keys) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_pack_keys: null buffer");
  uint64_t total = (uint64_t)nq * k;
  unsigned blocks = (unsigned)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  pack_keys_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_ids, d_dists, d_counts, nq, k, id_offset, d_keys);
  CUDA_TRY(cudaGetLastError());
  return NGTGPU_OK;
}

// keys: [n_lists][nq][k] ascending lists (KEY_NONE padded), as an all-gather of pack_keys outputs lays them out.
extern "C" int ngtgpu_merge_keys(const uint64_t *d_keys, uint32_t n_lists, uint32_t nq, uint32_t k, uint32_t *d_ids,
                                 float *d_dists, uint32_t *d_counts, void *stream) {
  if (nq == 0 || k == 0) return NGTGPU_OK;
  if (!d_keys || !d_ids || !d_dists || !d_counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_merge_keys: null buffer");
  if (n_lists == 0 || n_lists > 32) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_merge_keys: 1..32 lists");
  const unsigned mblocks = (nq + 7) / 8;
  scan_merge_kernel<<<mblocks, 256, 0, (cudaStream_t)stream>>>(d_keys, nq, n_lists, k, (uint64_t)k, (uint64_t)nq * k, d_ids,
                                                                d_dists, d_counts);
  CUDA_TRY(cudaGetLastError());
  return NGTGPU_OK;
}

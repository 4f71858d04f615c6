// knn_tc.cu -- exhaustive kNN for float objects on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// The brute-force pass behind linearSearch batches, kNN-graph construction and ground truth
// (lib/NGT/ObjectSpaceRepository.h:466-502, Index.h:839-856, Index.cpp:670-719) is a dense contraction:
//     L2      ||q - x||^2 = ||q||^2 + ||x||^2 - 2 q.x
//     cosine  q.x / (||q|| ||x||),  normalised kinds  q.x
// so q.x for a 128-query x 128-row tile is one tcgen05.mma chain with the accumulator in tensor memory.
// fp32 operands are split into bf16 parts, x = hi + lo, and the three products that matter are obtained from
// ONE bf16 GEMM over a concatenated K axis:  A' = [q_hi | q_hi | q_lo],  B' = [x_hi | x_lo | x_hi]
// (q_lo.x_lo, relative 2^-16, is dropped). When every value is exactly a bf16 number (SIFT-like integer
// data) a single segment is used and the products are exact. uint8 and Hamming objects go through kind::i8 as bytes
// (u8 x u8 -> s32, exact; tc_pack_i8_kernel for how the norms ride along).
//
// The tensor cores only FILTER. Each epilogue thread owns one query (one TMEM lane) and a buffer of (score, row) pairs
// in global memory (L2-resident): a row whose approximate score is within an error margin of the query's running
// threshold is appended with two predicated instructions; when a buffer is nearly full the WARP compacts it together
// (its 32 lanes hold the entries in registers, find the k-th smallest score by bisection over the ordered bit
// patterns with one warp-wide add per step, and write back the entries within the margin of it), which also
// refreshes the threshold. Between compactions the threshold is stale, i.e. larger: the tests admit a superset. There
// is no per-thread heap and no shared-memory state; what is left in the buffers at the end IS the candidate list
// (k plus the few rows inside the margin).
// A second kernel re-evaluates the candidates with the engine's one exact summation order
// (ngtgpu_internal.cuh) and selects the top k by (distance, id), so results are bit-identical to the
// CUDA-core scan (scan.cu) and to the reference where that is exact. A true k-nearest row cannot be dropped:
// its approximate score is <= d_k(1+e), the running threshold is >= d_k(1-e), and the margin is > 2e.
//
// Kernel anatomy (320 threads, 1 CTA/SM): warps 0-7 epilogue in two groups of four (tcgen05.ld, 32 TMEM lanes per
// warp; group g owns accumulator g, i.e. the even / the odd row tiles, with buffers and thresholds of its own, so the
// two groups never share state), warp 8 producer (cp.async.bulk of pre-swizzled 16 KB operand tiles, mbarrier
// complete_tx), warp 9 MMA issuer (one elected thread; M=128, N=128, K=16 per instruction, SWIZZLE_128B K-major
// descriptors). A CTA owns 256 queries (two query tiles; every row tile feeds both). Short K axes (<= 6 chunks of 64
// elements): the query operand stays resident in shared memory and row tiles stream through a ring of up to ten 16 KB
// stages. Long K axes (960-d data: 46 chunks as split floats): one k-chunk of both query tiles travels with every row
// chunk through a ring of four 48 KB stages, and the grid is ordered so that a wave is a few query-tile pairs times all
// row splits (the re-read query operand of a wave stays in L2). The two sets of 128-column accumulators in TMEM
// alternate, so each epilogue group has two tile times per tile.
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>

#include "ngtgpu_internal.cuh"

#define TC_TILE 128            // queries per CTA tile == rows per streamed tile
#define TC_KCHUNK 64           // bf16 elements per swizzle row (128 bytes)
#define TC_TILE_BYTES (TC_TILE * TC_KCHUNK * 2)   // 16 KB: one operand tile of one k-chunk
#define TC_STAGES 10
#define TC_MAX_K 128
// Epilogue warps: four (one per 32 TMEM lanes) per query tile and per column half. TC_HALVES = 1: a thread owns one query
// and all 128 columns of a row tile. TC_HALVES = 2 (a thread owns 64 columns, with a candidate buffer and a threshold of
// its own; sixteen epilogue warps) was measured SLOWER on every shape -- uint8 1M 0.434 -> 0.487 s, float 128-d 0.544 ->
// 0.649 s, 960-d 0.54 -> 0.67 s: two thresholds per query admit more candidates, and the accumulator read-back was not
// waiting for warps.
#ifndef TC_HALVES
#define TC_HALVES 1
#endif
#define TC_EPI_WARPS (8 * TC_HALVES)
#define TC_COLS (128 / TC_HALVES)              // accumulator columns per epilogue thread and row tile
#define TC_BLK (TC_HALVES == 1 ? 32 : 16)      // columns per tcgen05.ld
#define TC_THREADS ((TC_EPI_WARPS + 2) * 32)
#define TC_CAND_MAX 512        // largest (score, row) buffer per (query, split): 16 entries per lane during a compaction

struct TcPacked {
  uint8_t *tiles = nullptr;    // [n_tiles][kchunks][16 KB], SWIZZLE_128B K-major images
  float *norms = nullptr;      // [n_tiles * 128], +inf for padding / empty rows
  uint64_t n_rows = 0;
  uint64_t n_tiles = 0;
  uint32_t kchunks = 0;
  uint32_t nseg = 0;
};

// ---- operand preparation --------------------------------------------------------------------------------
// Element d of a stored row as a float, for the three object layouts the engine holds (kind 0: float32, 1: uint8,
// 2: bit vectors for Hamming -- one K element per bit). uint8 values and bits are bf16 numbers exactly, their products
// and sums (< 2^24) are exact in the fp32 accumulator, so for integer kinds the tensor-core score IS the squared L2
// distance / the Hamming distance (||a||^2 + ||b||^2 - 2ab with 0/1 entries).
__device__ __forceinline__ float tc_elem(const uint8_t *rows, int kind, uint64_t row, uint32_t row_bytes, uint32_t d) {
  const uint8_t *r = rows + row * (uint64_t)row_bytes;
  if (kind == 0) return reinterpret_cast<const float *>(r)[d];
  if (kind == 1) return (float)r[d];
  return (float)((r[d >> 3] >> (d & 7)) & 1);
}
__global__ void tc_check_exact_kernel(const float *__restrict__ rows, uint64_t count, int *flag) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    float x = rows[i];
    if (__bfloat162float(__float2bfloat16_rn(x)) != x) {
      *flag = 1;
      return;
    }
  }
}

// One thread per (row, 16-byte unit of the packed K axis). side 0 = query operand [hi|hi|lo], 1 = row operand
// [hi|lo|hi]. The unit is written where SWIZZLE_128B puts it: byte r*128 + ((u ^ (r & 7)) * 16) of the 16 KB tile.
// fold != 0 (L2): the row operand is scaled by -2 (exact in bf16) and one more k-chunk carries the squared norms,
// each split into three bf16 parts, against ones on the other side:
//     A extra = [qn0 qn1 qn2 1 1 1 0 ...],  B extra = [1 1 1 rn0 rn1 rn2 0 ...]
// so the accumulator IS ||q||^2 + ||x||^2 - 2 q.x and the epilogue only compares. Padding / empty rows get a huge
// finite norm (3e38) and can never pass.
__global__ void tc_pack_kernel(const uint8_t *__restrict__ rows, int kind, uint32_t row_bytes, uint64_t n_rows, uint64_t n_pad,
                               uint32_t padded_dim, int side, uint32_t nseg, uint32_t kchunks, int fold,
                               const float *__restrict__ norms, uint8_t *__restrict__ tiles) {
  const uint64_t units_per_row = (uint64_t)kchunks * 8;
  const uint64_t total = n_pad * units_per_row;
  const uint32_t fold_unit = (kchunks - 1) * 8;   // first unit of the extra chunk
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = i / units_per_row;
    const uint32_t U = (uint32_t)(i % units_per_row);
    __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = __float2bfloat16_rn(0.f);
    if (fold && U >= fold_unit) {
      if (U == fold_unit) {
        float nrm = norms[row];
        if (!(nrm < 3.0e38f)) nrm = 3.0e38f;
        const __nv_bfloat16 p0 = __float2bfloat16_rn(nrm);
        const float r1 = nrm - __bfloat162float(p0);
        const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
        const __nv_bfloat16 one = __float2bfloat16_rn(1.0f);
        if (side == 0) {
          v[0] = p0; v[1] = p1; v[2] = p2; v[3] = one; v[4] = one; v[5] = one;
        } else {
          v[0] = one; v[1] = one; v[2] = one; v[3] = p0; v[4] = p1; v[5] = p2;
        }
      }
    } else {
      const uint32_t kappa = U * 8;
      const uint32_t seg = kappa / padded_dim, d0 = kappa % padded_dim;
      if (row < n_rows && seg < nseg) {
        float src[8];
#pragma unroll
        for (int j = 0; j < 8; j++) src[j] = tc_elem(rows, kind, row, row_bytes, d0 + j);
        const bool want_lo = side == 0 ? seg == 2 : seg == 1;
        const float scale = (fold && side == 1) ? -2.0f : 1.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float x = src[j];
          const __nv_bfloat16 hi = __float2bfloat16_rn(x);
          v[j] = __float2bfloat16_rn(scale * (want_lo ? x - __bfloat162float(hi) : __bfloat162float(hi)));
        }
      }
    }
    const uint64_t tile = row / TC_TILE;
    const uint32_t r = (uint32_t)(row % TC_TILE), c = U / 8, u = U % 8;
    uint8_t *dst = tiles + (tile * kchunks + c) * (uint64_t)TC_TILE_BYTES + r * 128 + ((u ^ (r & 7)) * 16);
    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(v);
  }
}

// Integer kinds on tcgen05 kind::i8 (u8 x u8 -> s32, exact): one K element per byte (uint8 L2) or per bit (Hamming, 0 / 1),
// 128 elements per 128-byte swizzle row -- half the bytes of the bf16 image, and an MMA covers K = 32. Products of
// unsigned bytes only add up, so the row norm rides along as h = (M - ||x||^2) >> 1 >= 0 (M: a bound of all row norms)
// in a few extra K slots: h = 255 t + r is spread as bytes over `wslots` slots that meet 255 on the query side and one
// slot that meets 1. The accumulator is then q.x + h, and ||q||^2 + ||x||^2 - 2 q.x = ||q||^2 + M - 2 acc - (parity of
// M - ||x||^2): one subtraction per thread and a max tree per eight columns decide "nothing here", like the float path.
// Padding / empty rows: zero bytes, h = 0, integer norm 0x3f000000 (they never become candidates).
__global__ void tc_pack_i8_kernel(const uint8_t *__restrict__ rows, int kind, uint32_t row_bytes, uint64_t n_rows, uint64_t n_pad,
                                  uint32_t kdim, uint32_t kchunks, int side, int m_bound, uint32_t wslots,
                                  const float *__restrict__ norms, int32_t *__restrict__ norms_i, uint8_t *__restrict__ tiles) {
  const uint64_t units_per_row = (uint64_t)kchunks * 8;
  const uint64_t total = n_pad * units_per_row;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = i / units_per_row;
    const uint32_t U = (uint32_t)(i % units_per_row);
    const float nrm = norms[row];
    const bool live = row < n_rows && nrm < 1.0e9f;
    const int rn = live ? (int)nrm : 0x3f000000;
    const uint32_t h = live ? (uint32_t)(m_bound - rn) >> 1 : 0u;
    const uint32_t t = h / 255u, r = h % 255u;
    uint8_t v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint32_t d = U * 16 + j;
      uint32_t x = 0;
      if (d < kdim) {
        if (row < n_rows) {
          const uint8_t *rp = rows + row * (uint64_t)row_bytes;
          x = kind == 1 ? rp[d] : ((rp[d >> 3] >> (d & 7)) & 1u);
        }
      } else if (d < kdim + wslots) {
        const uint32_t s = d - kdim;
        x = side == 0 ? 255u : (t > 255u * s ? (t - 255u * s < 255u ? t - 255u * s : 255u) : 0u);
      } else if (d == kdim + wslots) {
        x = side == 0 ? 1u : r;
      }
      v[j] = (uint8_t)x;
    }
    if (U == 0 && norms_i) norms_i[row] = rn;
    const uint64_t tile = row / TC_TILE;
    const uint32_t rr = (uint32_t)(row % TC_TILE), c = U / 8, u = U % 8;
    uint8_t *dst = tiles + (tile * kchunks + c) * (uint64_t)TC_TILE_BYTES + rr * 128 + ((u ^ (rr & 7)) * 16);
    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(v);
  }
}

// largest finite squared norm of the rows (bounds the filter's error margin per query)
__global__ void tc_max_norm_kernel(const float *__restrict__ norms, uint64_t n, float *out) {
  float m = 0.f;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    float v = norms[i];
    if (v < 3.0e38f && v > m) m = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));   // non-negative floats order like ints
}

// squared norms (plain fp32: they only feed the filter); +inf marks padding and empty slots
__global__ void tc_norms_kernel(const uint8_t *__restrict__ rows, int kind, uint32_t row_bytes, uint64_t n_rows, uint64_t n_pad,
                                uint32_t padded_dim, uint32_t first_id, const uint8_t *__restrict__ valid, float *__restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n_pad; r += nwarps) {
    float s = 0.f;
    bool ok = r < n_rows && !(valid && valid[first_id + r] == 0);
    if (ok)
      for (uint32_t i = lane; i < padded_dim; i += 32) {
        float x = tc_elem(rows, kind, r, row_bytes, i);
        s = fmaf(x, x, s);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[r] = ok ? s : __int_as_float(0x7f800000);
  }
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// one lane of the (fully active) warp, chosen by the hardware: code under this predicate is known to the compiler to run in
// a single thread, so tcgen05.mma / cp.async.bulk take their descriptors through one R2UR each instead of a
// value-uniformity loop per instruction (VOTEU / ELECT / R2UR.BROADCAST / BRA.U.ANY around every UTCHMMA made the
// issuing thread, not the tensor pipe, the limit: 170-200 cycles per 92-cycle MMA)
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// SWIZZLE_128B, K-major operand tile (rows of 128 bytes, 8-row groups 1024 bytes apart): cute::UMMA::SmemDescriptor
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major): 1
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_tmem_ld(uint32_t taddr, uint32_t (&v)[32]) { tc_tmem_ld32(taddr, v); }
__device__ __forceinline__ void tc_tmem_ld(uint32_t taddr, uint32_t (&v)[16]) { tc_tmem_ld16(taddr, v); }
// extremes of a block of N values as trees (log depth instead of a chain)
template <int N>
__device__ __forceinline__ float tc_fmin(const uint32_t (&v)[N]) {
  float m[N / 2];
#pragma unroll
  for (int i = 0; i < N / 2; i++) m[i] = fminf(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
#pragma unroll
  for (int w = N / 4; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; i++) m[i] = fminf(m[i], m[i + w]);
  return m[0];
}
template <int N>
__device__ __forceinline__ float tc_fmax(const uint32_t (&v)[N]) {
  float m[N / 2];
#pragma unroll
  for (int i = 0; i < N / 2; i++) m[i] = fmaxf(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
#pragma unroll
  for (int w = N / 4; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; i++) m[i] = fmaxf(m[i], m[i + w]);
  return m[0];
}
template <int N>
__device__ __forceinline__ int tc_imax(const uint32_t (&v)[N]) {
  int m[N / 2];
#pragma unroll
  for (int i = 0; i < N / 2; i++) m[i] = max((int)v[2 * i], (int)v[2 * i + 1]);
#pragma unroll
  for (int w = N / 4; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; i++) m[i] = max(m[i], m[i + w]);
  return m[0];
}
__device__ __forceinline__ void tc_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- the filter kernel ------------------------------------------------------------------------------------
struct TcArgs {
  const uint8_t *a_tiles;   // packed query operand  [q_tiles][kchunks][16 KB]
  const uint8_t *b_tiles;   // packed row operand    [r_tiles][kchunks][16 KB]
  const float *a_norms;     // [q_tiles * 128]
  const float *b_norms;     // [r_tiles * 128]
  uint32_t nq;
  uint64_t n_rows;
  uint32_t kchunks;
  uint32_t k;
  int mode;                 // 0 L2, 1 dot (normalised kinds), 2 cosine, 3 integer L2 / Hamming on kind::i8 (accumulator = q.x, exact)
  float rel_margin;         // error bound of the bf16 product, relative to ||q||^2+||x||^2 (L2) or absolute (similarities)
  float max_row_norm;       // MODE 0: largest ||x||^2, the margin of a query is 2 * rel_margin * (||q||^2 + max_row_norm)
  uint32_t stages;          // ring depth (2..TC_STAGES)
  int exclude_self;
  uint32_t self_base;       // row index (0-based) of query 0 when queries are stored rows
  uint32_t qtiles;
  uint32_t nsplit;
  uint64_t tiles_per_split;
  uint32_t qgroups;         // query tiles per CTA: 2 (256 queries share every row tile) or 1 when two query operands do not fit
  int stream;               // 1: the query operand does not fit shared memory and streams through the ring with the rows (long K axis)
  const int32_t *b_norms_i; // MODE 3 (kind::i8): integer squared norms / popcounts of the rows
  int i8_m;                 // ... and their bound M (see tc_pack_i8_kernel)
  int debug_skip;           // development: 1 = the epilogue only releases the accumulators (timing of the MMA side alone)
  uint32_t cap;             // entries of one (query, split) buffer: a multiple of 32, <= TC_CAND_MAX
  uint2 *cand;              // [nq][nsplit][cap] (score bits, row index 0-based)
  uint32_t *cand_n;         // [nq][nsplit]; 0xffffffff = overflow
};

// Compaction of ONE query's buffer by the whole warp (`buf`, `cnt` are warp-uniform: broadcast from the owning lane).
// The K-th smallest score of the buffer becomes the query's threshold; entries within `margin` of it stay (everything else
// can never be within the margin of a later, smaller threshold). Returns the new count; *thr_out is left alone when the
// buffer holds fewer than K entries.
__device__ __noinline__ uint32_t tc_compact(uint2 *buf, uint32_t cnt, uint32_t K, float margin, float *thr_out, int lane) {
  constexpr int PER = TC_CAND_MAX / 32;
  uint32_t o[PER], r[PER];
  __syncwarp();   // the owner's appends are ordered before the other lanes' reads
#pragma unroll
  for (int i = 0; i < PER; i++) {
    const uint32_t idx = (uint32_t)i * 32u + (uint32_t)lane;
    o[i] = 0xffffffffu;
    r[i] = 0u;
    if (idx < cnt) {
      const uint2 e = __ldcg(buf + idx);
      o[i] = ord_of_float(__uint_as_float(e.x));
      r[i] = e.y;
    }
  }
  uint32_t limit = 0xfffffffeu;   // fewer than K entries: keep them all
  if (cnt >= K) {
    // the K-th smallest ordered value, bit by bit: the largest v with fewer than K entries below it
    uint32_t ans = 0;
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      const uint32_t c = ans | (1u << b);
      uint32_t below = 0;
#pragma unroll
      for (int i = 0; i < PER; i++) below += o[i] < c ? 1u : 0u;
      below = __reduce_add_sync(0xffffffffu, below);
      if (below < K) ans = c;
    }
    const float kth = float_of_ord(ans);
    *thr_out = kth;
    limit = ord_of_float(kth + margin);
  }
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < PER; i++) mine += o[i] <= limit ? 1u : 0u;
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  uint32_t pos = incl - mine;
  __syncwarp();   // every lane has its entries in registers before anything is overwritten
#pragma unroll
  for (int i = 0; i < PER; i++)
    if (o[i] <= limit) {
      __stcg(buf + pos, make_uint2(__float_as_uint(float_of_ord(o[i])), r[i]));
      pos++;
    }
  return __shfl_sync(0xffffffffu, incl, 31);
}

// Eight columns of one query against its threshold snapshot; out of line on purpose: the column loop of the epilogue must
// stay small (eight warps in two phases share the instruction cache; with the appends inlined sixteen times the kernel was
// instruction-fetch bound: 58 % I-cache hit rate, 6 "no instruction" stall cycles per issue).
template <int MODE>
__device__ __noinline__ uint32_t tc_append8(uint4 lo, uint4 hi, float thr, float margin, float qn, uint32_t rbase,
                                             const float *__restrict__ b_norms, uint2 *buf, uint32_t cnt, int m_bound) {
  const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  if (MODE == 3) {
    // v[] are q.x + h (exact integers, h = (M - ||x||^2) >> 1): with the rows' integer norms (`b_norms` carries them here;
    // `qn` the bits of ||q||^2 as an integer, `margin` unchanged, M in the sign-free upper half of `rbase`'s partner `m_bound`)
    const float lim = thr + margin;
    const int4 *rn4 = reinterpret_cast<const int4 *>(reinterpret_cast<const int32_t *>(b_norms) + rbase);
    const int4 r0 = __ldg(rn4), r1 = __ldg(rn4 + 1);
    const int rn[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const int qn_i = __float_as_int(qn);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int h = (m_bound - rn[j]) >> 1;
      const float sc = __int2float_rn((rn[j] + qn_i) - 2 * ((int)v[j] - h));
      if (rn[j] >= 0x3f000000) continue;   // padding / empty row
      if (sc <= lim) {
        __stcg(buf + cnt, make_uint2(__float_as_uint(sc), rbase + j));
        cnt++;
      }
    }
  } else if (MODE == 0) {
    const float lim = thr + margin;   // the accumulator already is ||q||^2 + ||x||^2 - 2 q.x
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (__uint_as_float(v[j]) <= lim) {
        __stcg(buf + cnt, make_uint2(v[j], rbase + j));
        cnt++;
      }
  } else {
    const float4 *rn4 = reinterpret_cast<const float4 *>(b_norms + rbase);
    const float4 r0 = __ldg(rn4), r1 = __ldg(rn4 + 1);
    const float rns[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float rn = rns[j];   // +inf for padding and empty slots: never passes
      const float dot = __uint_as_float(v[j]);
      float score;
      if (MODE == 1) score = rn < 3.0e38f ? -dot : __int_as_float(0x7fc00000);   // NaN never passes
      else score = rn < 3.0e38f ? -dot * rsqrtf(qn * rn) : __int_as_float(0x7fc00000);
      if (score - margin <= thr) {
        __stcg(buf + cnt, make_uint2(__float_as_uint(score), rbase + j));
        cnt++;
      }
    }
  }
  return cnt;
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_tc_filter_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles want 1024-byte alignment
  __shared__ __align__(8) uint64_t bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_a, bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // qtile: a.qgroups query tiles of 128. Resident query operand: the CTAs of a wave are different query tiles of one split
  // (they read the same row tiles at about the same time). Streamed query operand: a wave is a few query tiles times all
  // the splits, so that the query tiles re-read for every row tile (the wave's working set) stay in L2.
  const uint32_t qtile = a.stream ? blockIdx.x / a.nsplit : blockIdx.x % a.qtiles;
  const uint32_t split = a.stream ? blockIdx.x % a.nsplit : blockIdx.x / a.qtiles;
  const uint64_t total_tiles = (a.n_rows + TC_TILE - 1) / TC_TILE;
  const uint64_t t_begin = (uint64_t)split * a.tiles_per_split;
  uint64_t t_end = t_begin + a.tiles_per_split;
  if (t_end > total_tiles) t_end = total_tiles;
  const uint64_t ntiles = t_end > t_begin ? t_end - t_begin : 0;

  // resident: [qgroups x kchunks x 16 KB query operand][a.stages x 16 KB ring of row tiles]
  // streamed: a.stages x (qgroups + 1) x 16 KB, one k-chunk of every query tile and of the row tile per stage
  uint8_t *sA = smem;
  uint8_t *sB = a.stream ? smem : smem + (size_t)a.qgroups * a.kchunks * TC_TILE_BYTES;
  const uint32_t stage_bytes = a.stream ? (a.qgroups + 1) * TC_TILE_BYTES : TC_TILE_BYTES;
  const uint32_t b_off = a.stream ? a.qgroups * TC_TILE_BYTES : 0u;          // the row tile inside a stage

  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; i++) {
      tc_mbar_init(&bar_full[i], 1);
      tc_mbar_init(&bar_empty[i], 1);
    }
    tc_mbar_init(&bar_a, 1);
    for (int i = 0; i < 2; i++) {
      tc_mbar_init(&bar_tfull[i], 1);
      tc_mbar_init(&bar_tempty[i], 128 * TC_HALVES * a.qgroups);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {
    // 512 columns: two alternating sets of (one 128-column fp32 accumulator per query group)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == TC_EPI_WARPS) {
    // ===================== producer =====================
    if (ntiles && tc_elect_one()) {
      const uint8_t *a_src = a.a_tiles + (size_t)qtile * a.qgroups * a.kchunks * TC_TILE_BYTES;   // this CTA's query tiles are consecutive
      if (!a.stream) {
        tc_mbar_expect_tx(&bar_a, a.qgroups * a.kchunks * TC_TILE_BYTES);
        for (uint32_t c = 0; c < a.qgroups * a.kchunks; c++)
          tc_bulk_load(sA + (size_t)c * TC_TILE_BYTES, a_src + (size_t)c * TC_TILE_BYTES, TC_TILE_BYTES, &bar_a);
      }
      uint64_t it = 0;
      for (uint64_t t = 0; t < ntiles; t++) {
        for (uint32_t c = 0; c < a.kchunks; c++, it++) {
          const uint32_t st = (uint32_t)(it % a.stages);
          tc_mbar_wait(&bar_empty[st], (uint32_t)((it / a.stages) & 1) ^ 1u);
          tc_mbar_expect_tx(&bar_full[st], stage_bytes);
          uint8_t *dst = sB + (size_t)st * stage_bytes;
          if (a.stream)
            for (uint32_t qg = 0; qg < a.qgroups; qg++)
              tc_bulk_load(dst + (size_t)qg * TC_TILE_BYTES, a_src + ((size_t)qg * a.kchunks + c) * TC_TILE_BYTES, TC_TILE_BYTES,
                           &bar_full[st]);
          tc_bulk_load(dst + b_off, a.b_tiles + ((size_t)(t_begin + t) * a.kchunks + c) * TC_TILE_BYTES, TC_TILE_BYTES, &bar_full[st]);
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (ntiles && tc_elect_one()) {
      // kind::f16, A = B = BF16, D = F32, K-major both, N = 128, M = 128 (cute::UMMA::InstrDescriptor)
      // (MODE 3: kind::i8, A = B = unsigned 8-bit, D = S32, K = 32 per instruction: the same four 32-byte steps per chunk)
      const uint32_t idesc = MODE == 3 ? (2u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24)
                                       : (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t descA0 = tc_smem_desc(smem_u32(sA)), descB0 = tc_smem_desc(smem_u32(sB));
      if (!a.stream) {
        tc_mbar_wait(&bar_a, 0);
        tc_fence_after();
      }
      uint64_t it = 0;
      for (uint64_t t = 0; t < ntiles; t++) {
        const uint32_t acc = (uint32_t)(t & 1);
        tc_mbar_wait(&bar_tempty[acc], (uint32_t)((t >> 1) & 1) ^ 1u);
        tc_fence_after();
        for (uint32_t c = 0; c < a.kchunks; c++, it++) {
          const uint32_t st = (uint32_t)(it % a.stages);
          tc_mbar_wait(&bar_full[st], (uint32_t)((it / a.stages) & 1));
          tc_fence_after();
          for (uint32_t qg = 0; qg < a.qgroups; qg++) {   // every row tile in shared memory feeds all the CTA's query tiles
#pragma unroll
            for (uint32_t s = 0; s < TC_KCHUNK / 16; s++) {
              const uint32_t a_at = a.stream ? st * stage_bytes + qg * TC_TILE_BYTES : (qg * a.kchunks + c) * TC_TILE_BYTES;
              const uint64_t da = descA0 + (uint64_t)((a_at + s * 32) >> 4);
              const uint64_t db = descB0 + (uint64_t)((st * stage_bytes + b_off + s * 32) >> 4);
              if (MODE == 3) tc_mma_i8(tmem_base + (acc * 2 + qg) * 128, da, db, idesc, (c | s) != 0 ? 1u : 0u);
              else tc_mma_bf16(tmem_base + (acc * 2 + qg) * 128, da, db, idesc, (c | s) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&bar_empty[st]);   // frees the ring slot when these MMAs have read it
        }
        tc_commit(&bar_tfull[acc]);    // accumulator complete
      }
    }
  } else {
    // ===================== epilogue: one query per thread =====================
    const uint32_t g = ((uint32_t)warp >> 2) & 1u;     // query group: the CTA's first / second query tile
    const uint32_t half = (uint32_t)warp >> 3;         // which TC_COLS columns of every row tile
    const uint32_t qlane = (uint32_t)tid & 127u;       // TMEM lane == query of the tile (warp w reads lanes 32 (w % 4) ..)
    const uint32_t q = (qtile * a.qgroups + g) * TC_TILE + qlane;
    const bool q_ok = q < a.nq && g < a.qgroups;
    const float qn = g < a.qgroups ? a.a_norms[(size_t)(qtile * a.qgroups + g) * TC_TILE + qlane] : 0.f;
    // K-th smallest approximate score so far; "nothing yet" is a large finite number so that the huge scores of
    // padding / empty rows (3e38) never pass. When the queries are stored rows the row itself (score ~ 0) is one of
    // the K = k + 1 smallest and is dropped by the re-evaluation.
    const uint32_t K = a.k + (a.exclude_self ? 1u : 0u);
    float thr = q_ok ? 1.0e37f : -__int_as_float(0x7f800000);
    uint32_t cnt = 0;
    bool bad = false;
    const size_t list = ((size_t)(q_ok ? q : 0u) * a.nsplit + split) * TC_HALVES + half;
    uint2 *mybuf = a.cand + list * a.cap;
    const float m2 = 2.0f * a.rel_margin;
    const float margin = (MODE == 0 || MODE == 3) ? m2 * (qn + a.max_row_norm) : m2;   // L2: per-query bound of twice the error
    const int qn_i = MODE == 3 ? __float2int_rn(qn) : 0;
    const int t_i = qn_i + a.i8_m - 1;
    const uint32_t room = a.cap - TC_COLS;            // one tile can add TC_COLS entries
    for (uint64_t t = 0; t < ntiles && g < a.qgroups; t++) {
      const uint32_t acc = (uint32_t)(t & 1);
      tc_mbar_wait(&bar_tfull[acc], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (a.debug_skip) {
        tc_fence_before();
        tc_mbar_arrive(&bar_tempty[acc]);
        continue;
      }
      const uint64_t row0 = (t_begin + t) * TC_TILE + half * TC_COLS;
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (acc * 2 + g) * 128 + half * TC_COLS;
      uint32_t v[2][TC_BLK];
      tc_tmem_ld(taddr, v[0]);
#pragma unroll 1
      for (uint32_t cbp = 0; cbp < 2; cbp++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t cb = cbp * 2 + h;
          tc_tmem_wait_ld();
          if (cb + 1 < 4) tc_tmem_ld(taddr + (cb + 1) * TC_BLK, v[(h + 1) & 1]);   // the next block in flight under this one
          const uint32_t(&blk)[TC_BLK] = v[h];
          // Most blocks of 32 columns, and most groups of eight inside the others, hold nothing for any of the warp's 32
          // queries: a min (max) tree and one vote decide that (the tree over 32 values has the instruction-level
          // parallelism the four dependent group tests lack). The threshold is a snapshot (it only shrinks): the tests
          // admit a superset.
          if (MODE == 0) {
            if (!__any_sync(0xffffffffu, tc_fmin<TC_BLK>(blk) <= thr + margin)) continue;
          } else if (MODE == 1) {
            if (!__any_sync(0xffffffffu, -tc_fmax<TC_BLK>(blk) - margin <= thr)) continue;
          } else if (MODE == 3) {
            if (!__any_sync(0xffffffffu, __int2float_rn(t_i - 2 * tc_imax<TC_BLK>(blk)) <= thr + margin)) continue;
          }
#pragma unroll
          for (int c8 = 0; c8 < TC_BLK / 8; c8++) {
            bool maybe = true;   // cosine scales every column by its row norm: no cheap bound
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = blk[c8 * 8 + j];
            if (MODE == 3) {
              // acc = q.x + h: the score of a column is >= ||q||^2 + M - 1 - 2 acc, so the largest of eight accumulators decides
              const int mx = max(max(max((int)w[0], (int)w[1]), max((int)w[2], (int)w[3])),
                                 max(max((int)w[4], (int)w[5]), max((int)w[6], (int)w[7])));
              maybe = __any_sync(0xffffffffu, __int2float_rn(t_i - 2 * mx) <= thr + margin);
            } else if (MODE == 0) {
              const float m = fminf(fminf(fminf(__uint_as_float(blk[c8 * 8]), __uint_as_float(blk[c8 * 8 + 1])),
                                          fminf(__uint_as_float(blk[c8 * 8 + 2]), __uint_as_float(blk[c8 * 8 + 3]))),
                                    fminf(fminf(__uint_as_float(blk[c8 * 8 + 4]), __uint_as_float(blk[c8 * 8 + 5])),
                                          fminf(__uint_as_float(blk[c8 * 8 + 6]), __uint_as_float(blk[c8 * 8 + 7]))));
              maybe = __any_sync(0xffffffffu, m <= thr + margin);
            } else if (MODE == 1) {
              const float m = fmaxf(fmaxf(fmaxf(__uint_as_float(blk[c8 * 8]), __uint_as_float(blk[c8 * 8 + 1])),
                                          fmaxf(__uint_as_float(blk[c8 * 8 + 2]), __uint_as_float(blk[c8 * 8 + 3]))),
                                    fmaxf(fmaxf(__uint_as_float(blk[c8 * 8 + 4]), __uint_as_float(blk[c8 * 8 + 5])),
                                          fmaxf(__uint_as_float(blk[c8 * 8 + 6]), __uint_as_float(blk[c8 * 8 + 7]))));
              maybe = __any_sync(0xffffffffu, -m - margin <= thr);   // score = -dot
            }
            if (maybe)
              cnt = tc_append8<MODE>(make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]), thr, margin,
                                     MODE == 3 ? __int_as_float(qn_i) : qn, (uint32_t)(row0 + cb * TC_BLK + c8 * 8),
                                     MODE == 3 ? reinterpret_cast<const float *>(a.b_norms_i) : a.b_norms, mybuf, cnt, a.i8_m);
          }
        }
      }
      tc_fence_before();
      tc_mbar_arrive(&bar_tempty[acc]);   // the accumulator is free again: the next MMAs run under the compaction below
      // buffers that could not take another tile are compacted by the warp, one query at a time
      uint32_t full = __ballot_sync(0xffffffffu, cnt > room);
      while (full) {
        const int src = __ffs(full) - 1;
        full &= full - 1;
        const uint64_t bp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)mybuf, src);
        const uint32_t bc = __shfl_sync(0xffffffffu, cnt, src);
        const float bm = __shfl_sync(0xffffffffu, margin, src);
        float nthr = 0.f;
        const uint32_t kept = tc_compact(reinterpret_cast<uint2 *>((uintptr_t)bp), bc, K, bm, &nthr, lane);
        if (lane == src) {
          cnt = kept;
          thr = nthr;
          if (kept > room) {   // K plus the rows inside the margin do not fit: the batch goes to the exact scan
            bad = true;
            cnt = 0;
          }
        }
      }
    }
    // final compaction: what stays is the candidate list
    {
      uint32_t todo = __ballot_sync(0xffffffffu, cnt > K);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint64_t bp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)mybuf, src);
        const uint32_t bc = __shfl_sync(0xffffffffu, cnt, src);
        const float bm = __shfl_sync(0xffffffffu, margin, src);
        float nthr = 0.f;
        const uint32_t kept = tc_compact(reinterpret_cast<uint2 *>((uintptr_t)bp), bc, K, bm, &nthr, lane);
        if (lane == src) cnt = kept;
      }
    }
    if (q_ok) a.cand_n[list] = bad ? 0xffffffffu : cnt;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- exact re-evaluation of the candidates + top-k (one warp per query) -----------------------------------
struct RerankArgs {
  const uint8_t *queries;   // prepared rows
  const uint8_t *rows;
  uint32_t row_bytes, chunks;
  uint32_t nq, k, nsplit;
  uint32_t first_row_id;
  const uint32_t *id_map;
  float radius;
  int dtype;
  const uint2 *cand;        // (score bits, row) pairs: only the rows are used here
  const uint32_t *cand_n;
  uint32_t cap;
  int exclude_self;         // drop the candidate whose row is the query's own (queries are stored rows)
  uint32_t self_base;
  uint32_t *ids;
  float *dists;
  uint32_t *counts;
  uint32_t *overflowed;     // incremented once per query whose candidate list overflowed
};

template <int ACC, int G>
__global__ void __launch_bounds__(256) knn_tc_rerank_kernel(const RerankArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * 8 + warp;
  if (q >= a.nq) return;
  uint64_t *top = reinterpret_cast<uint64_t *>(smem) + (size_t)warp * a.k;
  uint32_t n = 0;
  constexpr int R = 32 / G;
  const int gl = lane % G, grp = lane / G;
  const uint8_t *qptr = a.queries + (size_t)q * a.row_bytes;
  bool bad = false;
  for (uint32_t s = 0; s < a.nsplit; s++) {
    uint32_t c = a.cand_n[(size_t)q * a.nsplit + s];
    if (c == 0xffffffffu) {
      bad = true;
      c = 0;
    }
    const uint2 *list = a.cand + ((size_t)q * a.nsplit + s) * a.cap;
    const uint32_t skip = a.exclude_self ? a.self_base + q : 0xffffffffu;
    for (uint32_t i0 = 0; i0 < c; i0 += R) {
      const uint32_t i = i0 + grp;
      const uint32_t row = i < c ? list[i].y : 0u;
      const bool act = i < c && row != skip;
      const float d = group_distance_gmem<ACC, G>(qptr, act ? a.rows + (size_t)row * a.row_bytes : qptr, a.chunks, gl, a.dtype);
      uint64_t key = KEY_NONE;
      if (act && gl == 0 && (a.radius < 0.f || d <= a.radius))
        key = make_key(d, a.id_map ? a.id_map[row] : a.first_row_id + row);
      uint32_t mm = __ballot_sync(0xffffffffu, key != KEY_NONE);
      while (mm) {
        int src = __ffs(mm) - 1;
        mm &= mm - 1;
        uint64_t kk = shfl_u64(key, src);
        sorted_insert_u64(top, n, a.k, kk, lane);
      }
    }
  }
  for (uint32_t i = lane; i < a.k; i += 32) {
    const bool ok = i < n;
    a.ids[(size_t)q * a.k + i] = ok ? key_id(top[i]) : 0u;
    a.dists[(size_t)q * a.k + i] = ok ? key_dist(top[i]) : 0.f;
  }
  if (lane == 0) {
    a.counts[q] = n;
    if (bad) atomicAdd(a.overflowed, 1u);
  }
}

template <int ACC>
static cudaError_t launch_rerank(int group, const RerankArgs &a, cudaStream_t stream) {
  const unsigned grid = (a.nq + 7) / 8;
  const size_t smem = (size_t)8 * a.k * 8;
  switch (group) {
    case 1: knn_tc_rerank_kernel<ACC, 1><<<grid, 256, smem, stream>>>(a); break;
    case 2: knn_tc_rerank_kernel<ACC, 2><<<grid, 256, smem, stream>>>(a); break;
    case 4: knn_tc_rerank_kernel<ACC, 4><<<grid, 256, smem, stream>>>(a); break;
    case 8: knn_tc_rerank_kernel<ACC, 8><<<grid, 256, smem, stream>>>(a); break;
    case 16: knn_tc_rerank_kernel<ACC, 16><<<grid, 256, smem, stream>>>(a); break;
    case 32: knn_tc_rerank_kernel<ACC, 32><<<grid, 256, smem, stream>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- host side -------------------------------------------------------------------------------------------
static int tc_kind(const ngtgpu_index *ix) {   // tc_elem's layout code
  return ix->object_type == NGTGPU_OBJECT_FLOAT ? 0 : ix->acc_kind == ACC_U8_HAM ? 2 : 1;
}
static uint32_t tc_kdim(const ngtgpu_index *ix) {   // K elements per stored row (one per bit for Hamming)
  return ix->acc_kind == ACC_U8_HAM ? ix->padded_dim * 8 : ix->padded_dim;
}
static uint32_t tc_kchunks(const ngtgpu_index *ix, uint32_t nseg, int fold) {
  return (nseg * tc_kdim(ix) + TC_KCHUNK - 1) / TC_KCHUNK + (fold ? 1 : 0);
}
// kind::i8: one byte per element, + the slots that carry (M - norm) / 2 (tc_pack_i8_kernel): wslots of weight 255, one of weight 1
static uint32_t tc_i8_wslots(int m_bound) { return (uint32_t)(((uint64_t)(m_bound / 2) / 255 + 254) / 255); }
static uint32_t tc_kchunks_i8(const ngtgpu_index *ix, uint32_t wslots) { return (tc_kdim(ix) + wslots + 1 + 127) / 128; }

// rows are padded (zeros, norm +inf) to a multiple of `pad_rows`
static int tc_pack(ngtgpu_index *ix, const uint8_t *d_rows, uint64_t n_rows, uint32_t pad_rows, int side, uint32_t nseg, int fold,
                   uint32_t first_id, const uint8_t *d_valid, uint8_t *tiles, float *norms, cudaStream_t stream,
                   int i8 = 0, int32_t *norms_i = nullptr, bool norms_done = false) {
  const uint32_t kchunks = i8 ? tc_kchunks_i8(ix, (uint32_t)ix->tc_i8_w) : tc_kchunks(ix, nseg, fold);
  const uint64_t n_pad = (n_rows + pad_rows - 1) / pad_rows * pad_rows;
  uint64_t total = n_pad * kchunks * 8;
  unsigned blocks = (unsigned)((total + 255) / 256 > (uint64_t)ix->sm_count * 64 ? (uint64_t)ix->sm_count * 64 : (total + 255) / 256);
  if (!norms_done)
    tc_norms_kernel<<<ix->sm_count * 8, 256, 0, stream>>>(d_rows, tc_kind(ix), ix->row_bytes, n_rows, n_pad, tc_kdim(ix), first_id, d_valid, norms);
  CUDA_TRY(cudaGetLastError());
  if (i8)
    tc_pack_i8_kernel<<<blocks, 256, 0, stream>>>(d_rows, tc_kind(ix), ix->row_bytes, n_rows, n_pad, tc_kdim(ix), kchunks, side, ix->tc_i8_m,
                                                  (uint32_t)ix->tc_i8_w, norms, norms_i, tiles);
  else
    tc_pack_kernel<<<blocks, 256, 0, stream>>>(d_rows, tc_kind(ix), ix->row_bytes, n_rows, n_pad, tc_kdim(ix), side, nseg, kchunks, fold, norms, tiles);
  CUDA_TRY(cudaGetLastError());
  ix->launches += 2;
  return NGTGPU_OK;
}

static bool all_bf16_exact(ngtgpu_index *ix, const float *d_rows, uint64_t count, int *d_flag, cudaStream_t stream, int *rc) {
  *rc = NGTGPU_OK;
  int h = 0;
  if (cudaMemsetAsync(d_flag, 0, sizeof(int), stream) != cudaSuccess) { *rc = NGTGPU_ERR_CUDA; return false; }
  tc_check_exact_kernel<<<ix->sm_count * 8, 256, 0, stream>>>(d_rows, count, d_flag);
  ix->launches++;
  if (cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess) {
    ngtgpu_set_error("tensor-core kNN: exactness check failed");
    *rc = NGTGPU_ERR_CUDA;
    return false;
  }
  return h == 0;
}

// Returns NGTGPU_OK with *used = 1 when the tensor-core path produced the results, *used = 0 when the shape is
// outside what it supports (the caller then runs the CUDA-core scan).
int ngtgpu_scan_topk_tc(ngtgpu_index *ix, const ScanParams &p, cudaStream_t stream, int *used) {
  *used = 0;
  const bool integer_kind = ix->object_type != NGTGPU_OBJECT_FLOAT;   // uint8 L2 / Hamming: exact in bf16 x bf16 -> fp32
  if (p.k == 0 || p.k > TC_MAX_K || p.approx) return NGTGPU_OK;
  if (p.d_id_map != nullptr) return NGTGPU_OK;                          // pivot tables stay on the CUDA-core path
  if (p.n_rows < 32768 || p.nq < 1024) return NGTGPU_OK;                // too little work to amortise packing
  if (p.d_rows != ix->d_objects + ix->row_bytes || p.n_rows != ix->n) return NGTGPU_OK;   // only the whole repository is cached
  if (!ix->tc_enabled) return NGTGPU_OK;

  int *d_flag = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_TC_MISC, 256, (void **)&d_flag));
  int rc = NGTGPU_OK;
  // ---- row operand: packed once per set_objects, kept with the index
  if (!ix->tc_rows_valid) {
    const uint8_t *rows = p.d_rows;
    bool exact = integer_kind || all_bf16_exact(ix, reinterpret_cast<const float *>(rows), p.n_rows * ix->padded_dim, d_flag, stream, &rc);
    if (rc != NGTGPU_OK) return rc;
    uint32_t nseg = exact ? 1 : 3;
    // integer kinds: u8 x u8 -> s32 on tcgen05 kind::i8 (NGTGPU_TC_I8=0: the bf16 image instead, for comparison)
    const char *i8_env = getenv("NGTGPU_TC_I8");
    const int i8 = integer_kind && !(i8_env && atoi(i8_env) == 0) ? 1 : 0;
    const int fold = i8 ? 0 : (ix->acc_kind == ACC_F_L2 || integer_kind) ? 1 : 0;   // L2 (and Hamming = L2 of bits): norms ride in the GEMM
    const uint64_t n_tiles = (p.n_rows + TC_TILE - 1) / TC_TILE;
    if (ix->d_tc_tiles) cudaFree(ix->d_tc_tiles);
    if (ix->d_tc_norms) cudaFree(ix->d_tc_norms);
    ix->d_tc_tiles = nullptr;
    ix->d_tc_norms = nullptr;
    // float norms, the largest of them, and (kind::i8) the integer norms behind
    CUDA_TRY(cudaMalloc(&ix->d_tc_norms, (n_tiles * TC_TILE + 64) * sizeof(float) + (i8 ? n_tiles * TC_TILE * sizeof(int32_t) : 0)));
    float *d_max = ix->d_tc_norms + n_tiles * TC_TILE;
    if (i8) {
      // the norms and their bound come first: the bound decides how many K slots the rows' images carry
      tc_norms_kernel<<<ix->sm_count * 8, 256, 0, stream>>>(rows, tc_kind(ix), ix->row_bytes, p.n_rows, n_tiles * TC_TILE, tc_kdim(ix),
                                                            p.first_row_id, p.d_valid, ix->d_tc_norms);
      CUDA_TRY(cudaMemsetAsync(d_max, 0, sizeof(float), stream));
      tc_max_norm_kernel<<<ix->sm_count * 4, 256, 0, stream>>>(ix->d_tc_norms, n_tiles * TC_TILE, d_max);
      ix->launches += 2;
      CUDA_TRY(cudaMemcpyAsync(&ix->tc_max_norm, d_max, sizeof(float), cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      if (!(ix->tc_max_norm < 1.0e9f)) return NGTGPU_OK;   // (norms beyond 2^30: the CUDA-core scan)
      ix->tc_i8_m = (int)ix->tc_max_norm + 2;              // >= every truncated norm
      ix->tc_i8_w = (int)tc_i8_wslots(ix->tc_i8_m);
    }
    const uint32_t kchunks = i8 ? tc_kchunks_i8(ix, (uint32_t)ix->tc_i8_w) : tc_kchunks(ix, nseg, fold);
    CUDA_TRY(cudaMalloc(&ix->d_tc_tiles, n_tiles * kchunks * (size_t)TC_TILE_BYTES));
    NGTGPU_TRY(tc_pack(ix, rows, p.n_rows, TC_TILE, 1, nseg, fold, p.first_row_id, p.d_valid, ix->d_tc_tiles, ix->d_tc_norms, stream, i8,
                       i8 ? reinterpret_cast<int32_t *>(ix->d_tc_norms + n_tiles * TC_TILE + 64) : nullptr, i8 != 0));
    if (!i8) {
      CUDA_TRY(cudaMemsetAsync(d_max, 0, sizeof(float), stream));
      tc_max_norm_kernel<<<ix->sm_count * 4, 256, 0, stream>>>(ix->d_tc_norms, n_tiles * TC_TILE, d_max);
      ix->launches++;
      CUDA_TRY(cudaMemcpyAsync(&ix->tc_max_norm, d_max, sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    ix->tc_nseg = nseg;
    ix->tc_kchunks = kchunks;
    ix->tc_fold = fold;
    ix->tc_i8 = i8;
    ix->tc_rows_valid = true;
  }
  uint32_t nseg = ix->tc_nseg;
  // ---- query operand
  const uint8_t *qrows = p.d_queries;
  if (nseg == 1 && !integer_kind) {
    bool qexact = all_bf16_exact(ix, reinterpret_cast<const float *>(qrows), (uint64_t)p.nq * ix->padded_dim, d_flag, stream, &rc);
    if (rc != NGTGPU_OK) return rc;
    if (!qexact) return NGTGPU_OK;   // rows are bf16-exact but the queries are not: leave it to the CUDA-core scan
  }
  const uint32_t kchunks = ix->tc_kchunks;
  // two query tiles per CTA: every row tile read from L2 feeds 256 queries (the kernel is bound by that stream). The query
  // operand is resident when both tiles and a ring of two stages fit shared memory (K axis of <= 6 chunks: 128-d split
  // floats, 384-d bf16-exact data); longer K axes (960-d, > 128-d split floats) stream it through the ring with the rows.
  const uint32_t qgroups = 2u;
  int stream_a = (2 * (size_t)kchunks + 2) * TC_TILE_BYTES + 1024 <= 225 * 1024 ? 0 : 1;
  if (const char *env = getenv("NGTGPU_TC_STREAM")) stream_a = atoi(env) != 0 || stream_a;   // development knob: force streaming
  const uint32_t qtiles128 = ((p.nq + TC_TILE * qgroups - 1) / (TC_TILE * qgroups)) * qgroups;   // 128-query tiles, padded to whole CTAs
  const uint32_t qtiles = qtiles128 / qgroups;                                                   // CTAs per split
  uint8_t *a_tiles = nullptr;
  float *a_norms = nullptr;
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_TC_QUERY, (size_t)qtiles128 * kchunks * TC_TILE_BYTES + (size_t)qtiles128 * TC_TILE * 8, (void **)&a_tiles));
  a_norms = reinterpret_cast<float *>(a_tiles + (size_t)qtiles128 * kchunks * TC_TILE_BYTES);
  NGTGPU_TRY(tc_pack(ix, qrows, p.nq, TC_TILE * qgroups, 0, nseg, ix->tc_fold, 0, nullptr, a_tiles, a_norms, stream, ix->tc_i8, nullptr));

  const uint64_t total_tiles_for_norms = (p.n_rows + TC_TILE - 1) / TC_TILE;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.a_tiles = a_tiles;
  a.b_tiles = ix->d_tc_tiles;
  a.a_norms = a_norms;
  a.b_norms = ix->d_tc_norms;
  a.nq = p.nq;
  a.n_rows = p.n_rows;
  a.kchunks = kchunks;
  a.k = p.k;
  a.mode = ix->tc_i8 ? 3 : (ix->acc_kind == ACC_F_L2 || integer_kind) ? 0 : ix->acc_kind == ACC_F_DOT ? 1 : 2;
  a.b_norms_i = ix->tc_i8 ? reinterpret_cast<const int32_t *>(ix->d_tc_norms + total_tiles_for_norms * TC_TILE + 64) : nullptr;
  a.i8_m = ix->tc_i8_m;
  // split floats: the dropped lo x lo products (2^-16) and the bf16 rounding of the lo parts, plus fp32 accumulation over
  // the K axis, which grows with its length
  a.rel_margin = nseg == 1 ? 4.0e-6f : std::max(1.0e-4f, 6.0e-8f * (float)(kchunks * TC_KCHUNK));
  a.max_row_norm = ix->tc_max_norm;
  a.exclude_self = p.exclude_self;
  a.self_base = p.self_base - p.first_row_id;
  a.qtiles = qtiles;
  a.qgroups = qgroups;
  a.stream = stream_a;
  const uint64_t total_tiles = (p.n_rows + TC_TILE - 1) / TC_TILE;
  uint64_t want = ((uint64_t)ix->sm_count + qtiles - 1) / qtiles;
  if (want > 8) want = 8;
  if (stream_a) want = 8;   // a wave = sm_count / 8 query-tile pairs x 8 splits: the pairs' operands stay in L2 between row tiles
  if (want > total_tiles) want = total_tiles;
  if (want < 1) want = 1;
  a.tiles_per_split = (total_tiles + want - 1) / want;
  a.nsplit = (uint32_t)((total_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
  // (score, row) buffers: room for K entries, the rows inside the margin and one more block of 32 columns
  const uint32_t K = p.k + (p.exclude_self ? 1u : 0u);
  uint32_t cap = (K + TC_COLS + 96 + 31u) & ~31u;
  if (cap > TC_CAND_MAX) cap = TC_CAND_MAX;
  a.cap = cap;
  a.debug_skip = getenv("NGTGPU_TC_SKIP_EPILOGUE") ? 1 : 0;
  uint8_t *cand_raw = nullptr;
  const size_t n_lists = (size_t)p.nq * a.nsplit * TC_HALVES;
  const size_t list_bytes = n_lists * cap * sizeof(uint2);
  NGTGPU_TRY(ngtgpu_scratch(ix, SCR_TC_CAND, list_bytes + (n_lists + 4) * 4, (void **)&cand_raw));
  a.cand = reinterpret_cast<uint2 *>(cand_raw);
  a.cand_n = reinterpret_cast<uint32_t *>(cand_raw + list_bytes);
  uint32_t *d_over = a.cand_n + n_lists;
  CUDA_TRY(cudaMemsetAsync(d_over, 0, 4, stream));

  // ring depth: as many 16 KB stages as fit beside the resident query operand (2..TC_STAGES)
  const size_t fixed_smem = (stream_a ? 0 : (size_t)qgroups * kchunks * TC_TILE_BYTES) + 1024;
  const size_t stage_bytes = stream_a ? (size_t)(qgroups + 1) * TC_TILE_BYTES : TC_TILE_BYTES;
  uint32_t stages = TC_STAGES;
  if (const char *env = getenv("NGTGPU_TC_STAGES")) stages = std::max(2, std::min(TC_STAGES, atoi(env)));   // development knob
  while (stages > 2 && fixed_smem + (size_t)stages * stage_bytes > 225 * 1024) stages--;
  if (fixed_smem + (size_t)stages * stage_bytes > 225 * 1024) return NGTGPU_OK;
  a.stages = stages;
  const size_t smem = fixed_smem + (size_t)stages * stage_bytes;
  if (a.mode == 3) {
    CUDA_TRY(cudaFuncSetAttribute(knn_tc_filter_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_tc_filter_kernel<3><<<qtiles * a.nsplit, TC_THREADS, smem, stream>>>(a);
  } else if (a.mode == 0) {
    CUDA_TRY(cudaFuncSetAttribute(knn_tc_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_tc_filter_kernel<0><<<qtiles * a.nsplit, TC_THREADS, smem, stream>>>(a);
  } else if (a.mode == 1) {
    CUDA_TRY(cudaFuncSetAttribute(knn_tc_filter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_tc_filter_kernel<1><<<qtiles * a.nsplit, TC_THREADS, smem, stream>>>(a);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(knn_tc_filter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_tc_filter_kernel<2><<<qtiles * a.nsplit, TC_THREADS, smem, stream>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  ix->launches++;

  RerankArgs r;
  memset(&r, 0, sizeof(r));
  r.queries = p.d_queries;
  r.rows = p.d_rows;
  r.row_bytes = ix->row_bytes;
  r.chunks = ix->chunks;
  r.nq = p.nq;
  r.k = p.k;
  r.nsplit = a.nsplit * TC_HALVES;
  r.first_row_id = p.first_row_id;
  r.id_map = nullptr;
  r.radius = p.radius;
  r.dtype = ix->distance_type;
  r.cand = a.cand;
  r.cand_n = a.cand_n;
  r.cap = a.cap;
  r.exclude_self = a.exclude_self;
  r.self_base = a.self_base;
  r.ids = p.d_ids;
  r.dists = p.d_dists;
  r.counts = p.d_counts;
  r.overflowed = d_over;
  cudaError_t e;
  switch (ix->acc_kind) {
    case ACC_F_L2: e = launch_rerank<ACC_F_L2>((int)ix->group, r, stream); break;
    case ACC_F_DOT: e = launch_rerank<ACC_F_DOT>((int)ix->group, r, stream); break;
    case ACC_U8_L2: e = launch_rerank<ACC_U8_L2>((int)ix->group, r, stream); break;
    case ACC_U8_HAM: e = launch_rerank<ACC_U8_HAM>((int)ix->group, r, stream); break;
    default: e = launch_rerank<ACC_F_COS>((int)ix->group, r, stream); break;
  }
  if (e != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("tensor-core kNN rerank launch: ") + cudaGetErrorString(e));
  ix->launches++;
  uint32_t h_over = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_over, d_over, 4, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (h_over) return NGTGPU_OK;   // a candidate list overflowed: the caller re-runs the batch on the exact CUDA-core scan
  *used = 1;
  return NGTGPU_OK;
}

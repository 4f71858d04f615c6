// search_fast.cuh -- the traversal kernel for the common case, written for instruction economy.
//
// Same algorithm, same state and same results as search_kernel (search.cuh; NeighborhoodGraph::search,
// lib/NGT/Graph.cpp:398-495 / 499-638), restricted to what the headline workloads use so that the hot loop is a
// few hundred instructions instead of a few thousand (ncu source attribution of the general kernel: 5.4k
// warp-instructions per expansion, a third of them address arithmetic around the row copies):
//
//   rows of <= 512 bytes (<= 32 chunks), edge cap <= 128 (head table), epsilon >= 0 (set semantics), k <= 128,
//   <= 128 seeds, visited hash in a slab of global memory (both on-chip tiers). Everything else runs search_kernel.
//
// One CTA (4 warps) per query, persistent grid. Per round:
//   control  warp 0 (its state lives in shared memory between rounds): merge the previous round's keys (only those
//            within the exploration radius were published), pop the front of the unchecked set (sorted registers),
//            make the popped node's head-table row available in shared memory (already there when the pop was
//            predicted -- 63 % of the pops), stage the row of the node expected next.
//   filter   edges dealt round-robin to the warps, one per thread, from shared memory; one 32-byte bucket read of the
//            visited hash (a single 256-bit load); every warp keeps the new ids of its own edges (no CTA-wide
//            compaction) and prefetches their rows towards L2.
//   rows     every warp copies and evaluates its own new ids in groups of four rows. A group is copied with cp.async
//            (ids by one 128-bit shared load, then the copies back to back, L2 evict-first), groups go through a ring
//            of buffers with one cp.async group each, so copies of later groups are in flight while a group is
//            evaluated. Distances: eight lanes per row, four rows per step, packed FADD2 / FFMA2 (same summation
//            order as every other kernel: chunk c on lane c mod 32 of group_fold<ACC, 32>), the scalar tail (sqrt,
//            key, publish) once per eight steps with one row per lane.
// Without a seed list the seeds are the nearest pivots of the seed table: from seed_select_kernel (below), or selected
// by this kernel itself when seed fusion is switched on.
#pragma once
#include "search.cuh"

// Warps per query (CTA), W: 4 (8 CTAs per SM) for every row width; 2 (16 CTAs per SM: twice the queries in flight) is
// offered for rows of <= 128 bytes with <= 64 edges and seeds per round, where a round moves a quarter of the bytes and
// its latency chain, not the bandwidth, sets the pace. 3 warps x 10 CTAs/SM measured the same as 4 x 8 on 512-byte rows.
#define FAST_WARPS 4
#define FAST_KL_MAX 4     // result keys per lane of the control warp when k > 32 (k <= 128)
#define FAST_STAGE_PER_WARP 4096u
#ifndef FAST_INT_LD
#define FAST_INT_LD 4u    // lanes per row for integer rows of <= 128 bytes in a distance step (8 rows per step)
#endif
__host__ __device__ constexpr uint32_t fast_stage_per_warp(int ch, int w) { return (ch == 1 && w <= 2) ? 2048u : 4096u; }

// ---- visited hash: the bucket in one 256-bit load -------------------------------------------------------
// Slots of a bucket are taken in increasing order (hash_insert tries slot i only after slot i - 1 was seen occupied),
// so the occupied slots are a prefix: the id is present iff some word equals it (a min over the xors), the bucket is
// full iff its last word is taken, and the first free slot is found by a three-step binary search.
__device__ __forceinline__ void hash_load256(const uint32_t *hash, uint32_t b, uint32_t (&v)[8]) {
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(hash + (size_t)b * 8));
}
// `v` holds the home bucket of `nid` already (hash_load256 issued earlier, so that several probes of one thread overlap)
__device__ __forceinline__ bool hash_lookup256_loaded(const uint32_t *hash, uint32_t bucket_bits, uint32_t nid, uint32_t b, uint32_t (&v)[8],
                                                      BucketProbe &bp) {
  const uint32_t bmask = (1u << bucket_bits) - 1u;
  for (;;) {
    const uint32_t x = min(min(min(v[0] ^ nid, v[1] ^ nid), min(v[2] ^ nid, v[3] ^ nid)),
                           min(min(v[4] ^ nid, v[5] ^ nid), min(v[6] ^ nid, v[7] ^ nid)));
    if (x == 0u) return true;   // visited
    if (v[7] == 0u) {
      uint32_t s0 = v[3] != 0u ? 4u : 0u;
      const uint32_t m1 = s0 ? v[5] : v[1];
      s0 += m1 != 0u ? 2u : 0u;
      const uint32_t lo = (s0 & 2u) ? v[2] : v[0], hi = (s0 & 2u) ? v[6] : v[4];
      const uint32_t m0 = (s0 & 4u) ? hi : lo;
      s0 += m0 != 0u ? 1u : 0u;
      bp.bucket = b;
      bp.slot = s0;
      return false;
    }
    b = (b + 1) & bmask;
    hash_load256(hash, b, v);
  }
}
__device__ __forceinline__ bool hash_lookup256(const uint32_t *hash, uint32_t bucket_bits, uint32_t nid, BucketProbe &bp) {
  const uint32_t bmask = (1u << bucket_bits) - 1u;
  uint32_t b = (nid * 2654435761u) >> (32 - bucket_bits);
  for (;;) {
    uint32_t v[8];
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(hash + (size_t)b * 8));
    const uint32_t x = min(min(min(v[0] ^ nid, v[1] ^ nid), min(v[2] ^ nid, v[3] ^ nid)),
                           min(min(v[4] ^ nid, v[5] ^ nid), min(v[6] ^ nid, v[7] ^ nid)));
    if (x == 0u) return true;   // visited
    if (v[7] == 0u) {
      uint32_t s0 = v[3] != 0u ? 4u : 0u;
      const uint32_t m1 = s0 ? v[5] : v[1];
      s0 += m1 != 0u ? 2u : 0u;
      const uint32_t lo = (s0 & 2u) ? v[2] : v[0], hi = (s0 & 2u) ? v[6] : v[4];
      const uint32_t m0 = (s0 & 4u) ? hi : lo;
      s0 += m0 != 0u ? 1u : 0u;
      bp.bucket = b;
      bp.slot = s0;
      return false;
    }
    b = (b + 1) & bmask;
  }
}

// 16-byte copy with zero fill (src_bytes 0 or 16) and an L2 eviction policy: rows are touched once per query and must
// not push the visited-hash slabs and head-table rows out of L2
__device__ __forceinline__ void cp_async_s16z_hint(uint32_t smem_addr, const void *gsrc, uint32_t src_bytes, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(smem_addr), "l"(gsrc), "r"(src_bytes),
               "l"(policy)
               : "memory");
}

// base + 16 * off16 in one IMAD.WIDE (the base is the kernel parameter: an aligned register pair)
__device__ __forceinline__ const uint8_t *addr16(uint32_t off16, const uint8_t *base) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(r) : "r"(off16), "l"((uint64_t)(uintptr_t)base));
  return reinterpret_cast<const uint8_t *>((uintptr_t)r);
}

template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- result list of the control warp: KL keys per lane, sorted; position p lives on lane p / KL, slot p % KL -----------
// KL == 1 serves k <= 32 (one key per lane: every headline search); KL == 4 serves k <= 128 -- the searches of the
// construction loop and of refineANNG, whose k is the edge count (Index.h:815-837, GraphReconstructor.h:852).
template <int KL>
__device__ __forceinline__ uint64_t res_kth(const uint64_t (&res)[KL], uint32_t idx) {
  uint64_t v = res[0];
  if (KL > 1) {
    const uint32_t j = idx % KL;
#pragma unroll
    for (int m = 1; m < KL; m++)
      if (j == (uint32_t)m) v = res[m];
  }
  return shfl_u64(v, (int)(idx / KL));
}
// insert a key that is not in the list; entries past position k - 1 fall off
template <int KL>
__device__ __forceinline__ void res_insert(uint64_t (&res)[KL], uint64_t kk, uint32_t k, int lane) {
  uint32_t pos = 0;
#pragma unroll
  for (int m = 0; m < KL; m++) pos += __popc(__ballot_sync(0xffffffffu, res[m] < kk));
  const uint64_t up = shfl_up_u64(res[KL - 1], 1);
#pragma unroll
  for (int m = KL - 1; m >= 0; m--) {
    const uint32_t idx = (uint32_t)lane * KL + m;
    const uint64_t prev = m > 0 ? res[m > 0 ? m - 1 : 0] : up;
    if (idx == pos) res[m] = kk;
    else if (idx > pos) res[m] = prev;
    if (idx >= k) res[m] = KEY_NONE;
  }
}

// W == 1 (opt-in, ngtgpu_index_set_fast_shape): the whole round runs in one warp (32 CTAs per SM, nobody waiting at a
// barrier behind the control chain): a round of up to 64 edges is filtered in two passes whose bucket reads are issued
// together, CTA barriers become warp barriers, and the unsorted back of the unchecked set lives in a per-CTA slab of
// global memory next to the visited hash so that 32 CTAs fit the SM's shared memory. Measured slower than W == 2 on
// 128-byte rows (3.6 vs 3.2 ms per 10k batch, 3.0 vs 2.9 ms per 10k at batch 40k): twice the slabs fall out of L2.
template <int ACC, int CH, int W, int KL>
__global__ void __launch_bounds__(W * 32, 32 / W) search_fast_kernel(const SearchArgs a) {
  constexpr int P = W <= 2 ? 2 : 1;                            // filter passes = edges per thread and round (the second
                                                               // one only when the round has more than 32 * W edges)
  constexpr int EDG = W == 1 ? 64 : SEARCH_HEAD;               // edges (or seeds) of one round
  constexpr int CK = W == 1 ? 64 : SEARCH_CMAX;                // keys one round can publish
#define FAST_SYNC()                 \
  {                                 \
    if (W == 1) __syncwarp();       \
    else __syncthreads();           \
  }
  constexpr uint32_t SPW = fast_stage_per_warp(CH, W);         // staging ring of one warp
  constexpr uint32_t SROW = 128u * CH;                         // staging stride of a row
  // lanes per row in a distance step: eight (four rows per step; the float summation order needs it), or FAST_INT_LD
  // for integer rows of <= 128 bytes, whose sums are exact in any order: 32 / LD rows per step, so the per-step
  // overhead (waits, syncs, issue of the next group, loop) is paid once per 8 or 16 rows instead of once per 4
  constexpr uint32_t LD = (CH == 1 && (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM)) ? FAST_INT_LD : 8u;
  constexpr uint32_t RPS = 32u / LD;                           // rows per step = rows of one group
  constexpr uint32_t CPL = 8u * CH / LD;                       // 16-byte chunks of a row per lane
  constexpr uint32_t GBYTES = RPS * SROW;                      // one group
  constexpr int NB = (int)(SPW / GBYTES);                      // ring depth per warp: 2, 4 or 8
  static_assert(NB >= 2, "the staging ring must hold two groups");
  constexpr uint32_t RPI = 4u / CH;                            // rows per copy instruction
  constexpr uint32_t LPR = 8u * CH;                            // lanes per row in a copy instruction
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(16) uint32_t s_wids[W][32 * P + 16];  // new ids per warp (+ padding up to a whole group)
  __shared__ uint32_t s_wcnt[W], s_wval[W];
  __shared__ uint64_t s_cand_keys[CK];
  __shared__ __align__(16) uint32_t s_edges[2][EDG];   // edge lists of the round / of the expected next node
  __shared__ uint32_t s_key_n;
  __shared__ int s_state;         // 0 run, 1 finished, 2 overflow
  __shared__ uint32_t s_query;
  __shared__ uint32_t s_take;     // edges to filter this round
  __shared__ uint32_t s_buf;      // which half of s_edges holds them
  __shared__ int s_seeding;
  __shared__ float s_er;
  // control-warp state between rounds: it lives here while the rows are copied and evaluated, so that the row loop
  // has the register file to itself (lane constants stay in registers instead of being recomputed every step)
  __shared__ uint64_t s_res[32 * KL], s_front[32];   // s_res[32 * m + lane]: slot m of the lane
  __shared__ uint64_t s_T;
  __shared__ uint32_t s_ctl[12];   // fn, qsize, res_n, visited_n, st_dist, st_edge, st_exp, pref_id, buf, flags, radius, er

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  uint8_t *stage = smem_raw;                                                   // 4 x SPW
  uint64_t *queue = W == 1 ? a.queue_slabs + (size_t)blockIdx.x * a.queue_cap : reinterpret_cast<uint64_t *>(smem_raw + W * SPW);
  uint32_t *hash = a.hash_slabs + ((size_t)blockIdx.x << a.hash_bits);
  const uint32_t bucket_bits = a.hash_bits - 3;
  const uint32_t take_head = a.edge_cap < SEARCH_HEAD ? a.edge_cap : SEARCH_HEAD;

  // this lane's constant part of the row copies and reads
  const uint32_t wstage_s = (uint32_t)__cvta_generic_to_shared(stage + (size_t)warp * SPW);
  const uint32_t cp_row = (uint32_t)lane / LPR;                   // row inside one copy instruction
  const uint32_t cp_chunk = (uint32_t)lane % LPR;
  // LD < 8: chunk c of row r of a group sits at position (c + LD * r) % 8 of the row's 128 bytes, so that the 16-byte
  // reads of a quarter-warp (8 / LD rows x LD lanes) fall into eight distinct bank groups
  const uint32_t cp_pos = LD == 8u ? cp_chunk : ((cp_chunk + LD * cp_row) & 7u);
  const uint32_t cp_dst = wstage_s + cp_row * SROW + cp_pos * 16u;
  const uint32_t rb16 = a.row_bytes >> 4;                         // a row in 16-byte units (tables up to 64 GB)
  const uint32_t cp_size = cp_chunk < a.chunks ? 16u : 0u;        // chunks past the row's end are zero-filled
  const uint32_t rr = (uint32_t)lane / LD;                        // row inside a distance step
  const uint32_t ll = (uint32_t)lane % LD;                        // lane inside the row
  const uint32_t rd = wstage_s + rr * SROW + (LD == 8u ? ll : ((ll + LD * rr) & 7u)) * 16u;   // its first chunk
  uint64_t row_policy;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(row_policy));
#define FAST_ROWCP(dst, src, bytes) cp_async_s16z_hint(dst, src, bytes, row_policy)

  for (;;) {
    // ---- next query (dynamic scheduling over a persistent grid)
    if (tid == 0) {
      uint32_t w = atomicAdd(a.work_counter, 1u);
      const uint32_t total = a.query_list ? *a.query_list_count : a.nq;   // later tiers: the previous tier's overflow list
      s_query = w < total ? (a.query_list ? a.query_list[w] : w) : 0xffffffffu;
      s_state = 0;
      s_key_n = 0;
    }
    FAST_SYNC();
    const uint32_t q = s_query;
    if (q == 0xffffffffu) break;
    {
      uint4 *h4 = reinterpret_cast<uint4 *>(hash);
      for (uint32_t i = tid; i < (1u << a.hash_bits) / 4; i += (W * 32)) h4[i] = zero16();
    }
    const uint8_t *qrow = a.queries + (size_t)q * a.row_bytes;
    uint4 q8[CPL];   // lane (rr, ll) holds query chunks ll, ll + LD, ...
#pragma unroll
    for (int m = 0; m < (int)CPL; m++) {
      const uint32_t c = ll + m * LD;
      q8[m] = c < a.chunks ? ldg16(qrow + (size_t)c * 16) : zero16();
    }
    float qn = 0.f;
    if (ACC == ACC_F_COS) {
      // query norm^2 in the engine's summation order (lane c owns chunk c, xor butterfly)
      const uint4 v = (uint32_t)lane < a.chunks ? ldg16(qrow + (size_t)lane * 16) : zero16();
      float a0 = __uint_as_float(v.x), a1 = __uint_as_float(v.y), a2 = __uint_as_float(v.z), a3 = __uint_as_float(v.w);
      qn = fmaf(a0, a0, qn);
      qn = fmaf(a1, a1, qn);
      qn = fmaf(a2, a2, qn);
      qn = fmaf(a3, a3, qn);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
    }
    FAST_SYNC();   // the slab is zero before anybody probes it

    if (a.seeds == nullptr) {
      // ---- seeds: the n_seeds nearest pivots of the seed table (the same selection as seed_select_kernel, without
      // its launch): every warp scans a quarter of the table, four rows per step read straight from L1/L2, and keeps
      // its k smallest keys sorted one per lane; warp 0 merges the lists into the first round's edge list
      uint64_t wres = KEY_NONE, wthr = KEY_NONE;
      uint4 qp[CH];   // eight lanes per pivot row here whatever LD is: lane j of a row holds query chunks j, j + 8, ...
#pragma unroll
      for (int m = 0; m < CH; m++) {
        const uint32_t c = ((uint32_t)lane & 7u) + m * 8;
        qp[m] = c < a.chunks ? ldg16(qrow + (size_t)c * 16) : zero16();
      }
      for (uint32_t p0 = 4u * (uint32_t)warp; p0 < a.n_pivots; p0 += 4u * W) {
        const uint32_t row = p0 + ((uint32_t)lane >> 3);
        const bool valid = row < a.n_pivots;
        const uint8_t *rp = a.pivots + (size_t)(valid ? row : 0u) * a.row_bytes;
        Sums p[CH];
#pragma unroll
        for (int m = 0; m < CH; m++) {
          const uint32_t c = ((uint32_t)lane & 7u) + m * 8;
          p[m] = zero_sums();
          acc_chunk_packed<ACC>(p[m], qp[m], c < a.chunks ? ldg16(rp + (size_t)c * 16) : zero16());
          lane_total<ACC>(p[m]);
        }
        Sums tot = p[0];
        if (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM) {
          if (CH == 2) tot.u = p[0].u + p[1].u;
          if (CH == 4) tot.u = (p[0].u + p[2].u) + (p[1].u + p[3].u);
          tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 4);
          tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 2);
          tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 1);
        } else {
          if (CH == 2) tot.f0 = p[0].f0 + p[1].f0;
          if (CH == 4) tot.f0 = (p[0].f0 + p[2].f0) + (p[1].f0 + p[3].f0);
          tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 4);
          tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 2);
          tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 1);
          if (ACC == ACC_F_COS) {
            if (CH == 2) tot.f1 = p[0].f1 + p[1].f1;
            if (CH == 4) tot.f1 = (p[0].f1 + p[2].f1) + (p[1].f1 + p[3].f1);
            tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 4);
            tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 2);
            tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 1);
          }
        }
        uint64_t key = KEY_NONE;
        if (((uint32_t)lane & 7u) == 0u && valid) key = make_key(finish_distance<ACC>(a.dtype, tot, qn), __ldg(a.pivot_ids + row));
        uint32_t m = __ballot_sync(0xffffffffu, key < wthr);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const uint64_t kk = shfl_u64(key, src);
          if (kk >= wthr) continue;
          const uint32_t pos = __popc(__ballot_sync(0xffffffffu, wres < kk));
          const uint64_t up = shfl_up_u64(wres, 1);
          if ((uint32_t)lane == pos) wres = kk;
          else if ((uint32_t)lane > pos) wres = up;
          if ((uint32_t)lane >= a.n_seeds) wres = KEY_NONE;
          wthr = shfl_u64(wres, (int)a.n_seeds - 1);
        }
      }
      s_cand_keys[warp * 32 + lane] = wres;
      FAST_SYNC();
      if (warp == 0) {
        uint64_t mres = KEY_NONE, mthr = KEY_NONE;
        for (int w = 0; w < W; w++) {
          const uint64_t key = s_cand_keys[w * 32 + lane];
          uint32_t m = __ballot_sync(0xffffffffu, key < mthr);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t kk = shfl_u64(key, src);
            if (kk >= mthr) continue;
            const uint32_t pos = __popc(__ballot_sync(0xffffffffu, mres < kk));
            const uint64_t up = shfl_up_u64(mres, 1);
            if ((uint32_t)lane == pos) mres = kk;
            else if ((uint32_t)lane > pos) mres = up;
            if ((uint32_t)lane >= a.n_seeds) mres = KEY_NONE;
            mthr = shfl_u64(mres, (int)a.n_seeds - 1);
          }
        }
        if ((uint32_t)lane < a.n_seeds) {
          const uint32_t sid = mres != KEY_NONE ? key_id(mres) : 0u;
          s_edges[0][lane] = sid;
          a.seeds_out[(size_t)q * a.n_seeds + lane] = sid;
        }
        __syncwarp();
      }
    }

    // ---- control-warp state (kept in shared memory between rounds)
    if (warp == 0) {
#pragma unroll
      for (int m = 0; m < KL; m++) s_res[32 * m + lane] = KEY_NONE;     // result list: lane i holds the keys of positions i * KL .. (k <= 32 * KL)
      s_front[lane] = KEY_NONE;
      if (lane == 0) s_T = KEY_NONE;
      if (lane < 10) s_ctl[lane] = lane == 9 ? 1u : 0u;   // flags: bit 0 seeding, bit 1 seeds taken, bit 2 head round, bit 3 a round ran
      if (lane == 10) s_ctl[10] = __float_as_uint(a.radius);
      if (lane == 11) s_ctl[11] = __float_as_uint(a.coef * a.radius);
      __syncwarp();
    }

    for (;;) {
      // The head-table row of the node expected to be popped next was staged during the round that just ended: the warps
      // that wait for the control warp pull the visited-hash buckets of its edges towards L2 meanwhile (a hint only: when
      // the prediction fails, 37 % of the rounds, the lines are simply not used). uint8 128-byte rows 3.28 -> 3.19 ms per
      // 10k batch, 512-byte float rows 3.39 -> 3.37 ms.
      if (W > 1 && warp != 0) {
        const uint32_t pid = s_ctl[7];
        if (pid) {
          const uint32_t *pe = s_edges[(s_buf ^ 1u) & 1u];
          for (uint32_t e = (uint32_t)(warp - 1) * 32u + (uint32_t)lane; e < take_head; e += (uint32_t)(W - 1) * 32u) {
            const uint32_t nid = pe[e];
            if (nid != 0u && nid <= a.n)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(hash + (size_t)((nid * 2654435761u) >> (32 - bucket_bits)) * 8));
          }
        }
      }
      // ================= control (warp 0) =================
      if (warp == 0) {
        uint64_t res[KL];
#pragma unroll
        for (int m = 0; m < KL; m++) res[m] = s_res[32 * m + lane];
        Unchecked U;
        U.front = s_front[lane];
        U.T = s_T;
        U.fn = s_ctl[0];
        U.qsize = s_ctl[1];
        U.queue = queue;
        U.cap = a.queue_cap;
        uint32_t res_n = s_ctl[2], visited_n = s_ctl[3], st_dist = s_ctl[4], st_edge = s_ctl[5], st_exp = s_ctl[6];
        uint32_t pref_id = s_ctl[7], buf = s_ctl[8];   // s_edges[buf ^ 1] holds the head row of node pref_id (0: nothing)
        const uint32_t flags = s_ctl[9];
        bool seeding = (flags & 1u) != 0, seeds_taken = (flags & 2u) != 0, head_round = (flags & 4u) != 0;
        const bool rounds_done = (flags & 8u) != 0;
        float radius = __uint_as_float(s_ctl[10]);
        float er = __uint_as_float(s_ctl[11]);
        uint32_t cand_n = 0;
        bool overflow = false, finished = false;
        if (rounds_done) {
#pragma unroll
          for (int w = 0; w < W; w++) cand_n += s_wcnt[w];
          if (head_round) {
#pragma unroll
            for (int w = 0; w < W; w++) st_edge += s_wval[w];
          }
        }
        visited_n += cand_n;
        st_dist += cand_n;
        const uint32_t key_n = s_key_n;
        if (key_n) {
          for (uint32_t j0 = 0; j0 < key_n; j0 += 32) {
            const uint64_t key = j0 + lane < key_n ? s_cand_keys[j0 + lane] : KEY_NONE;   // KEY_NONE compares false (NaN)
            uint32_t m = __ballot_sync(0xffffffffu, key_dist(key) <= radius);
            while (m) {
              const int src = __ffs(m) - 1;
              m &= m - 1;
              const uint64_t kk = shfl_u64(key, src);
              if (key_dist(kk) > radius) continue;   // the radius shrank meanwhile
              res_insert<KL>(res, kk, a.k, lane);
              if (res_n < a.k) res_n++;
              if (!seeding && res_n >= a.k) radius = key_dist(res_kth<KL>(res, a.k - 1));
            }
          }
          if (!seeding) er = a.coef * radius;
          for (uint32_t j0 = 0; j0 < key_n && !overflow; j0 += 32) {
            const uint64_t key = j0 + lane < key_n ? s_cand_keys[j0 + lane] : KEY_NONE;
            const bool acc = key != KEY_NONE && (seeding || key_dist(key) <= er);
            const bool low = acc && key < U.T;
            const uint32_t bm = __ballot_sync(0xffffffffu, acc && !low);
            uint32_t fm = __ballot_sync(0xffffffffu, low);
            const uint32_t cnt = __popc(bm);
            if (cnt) {
              if (U.qsize + cnt > U.cap) back_compact(U, er, lane);
              if (U.qsize + cnt > U.cap) {
                overflow = true;
                break;
              }
              if (acc && !low) U.queue[U.qsize + __popc(bm & lanemask_lt())] = key;
              U.qsize += cnt;
              __syncwarp();
            }
            while (fm) {
              const int src = __ffs(fm) - 1;
              fm &= fm - 1;
              if (!unchecked_insert(U, shfl_u64(key, src), er, lane, nullptr, a.edge_cap)) {
                overflow = true;
                break;
              }
            }
          }
        }
        head_round = false;
        uint32_t take = 0;
        if (!overflow) {
          if (!seeds_taken) {
            // the seed list is the first round's edge list (setupDistances/setupSeeds, Graph.cpp:243-394)
            seeds_taken = true;
            take = a.n_seeds;
            if (a.seeds) {   // (else the kernel selected them above: they are in s_edges[0] already)
              const uint32_t *sp = a.seeds + (size_t)q * a.n_seeds;
              for (uint32_t i = lane; i < take; i += 32) s_edges[buf][i] = __ldg(sp + i);
            }
          } else {
            if (seeding) {
              // setupSeeds: radius from the seeds once k of them are within it (Graph.cpp:349-351)
              seeding = false;
              if (res_n >= a.k) radius = key_dist(res_kth<KL>(res, a.k - 1));
              er = a.coef * radius;
            }
            if (U.fn == 0 && U.qsize != 0 && key_dist(U.T) <= er) {
              front_refill(U, er, lane);
              const uint64_t mine = U.front;
              if (mine != KEY_NONE) prefetch_head_row(a.head, key_id(mine), a.edge_cap);
            }
            const uint64_t best = shfl_u64(U.front, 0);
            if (U.fn == 0 || !(key_dist(best) <= er)) {   // Graph.cpp:430-435
              finished = true;
            } else {
              U.front = shfl_down_u64(U.front, 1);
              if (lane == 31) U.front = KEY_NONE;
              U.fn--;
              const uint32_t t = key_id(best);
              st_exp++;
              take = take_head;
              head_round = true;
              buf ^= 1u;   // the half that may already hold t's row
              const bool staged = t == pref_id;
              if (!staged && (uint32_t)lane * 4u < take)
                cp_async_row16(&s_edges[buf][lane * 4], a.head + (size_t)t * SEARCH_HEAD + lane * 4);
              // the node most likely to be popped next is the front's new first key: stage its row in the other half
              // while this round's rows are in flight
              const uint64_t nxt = shfl_u64(U.front, 0);
              pref_id = 0;
              if (!staged) cp_async_commit_wait_all();
              if (U.fn != 0 && key_dist(nxt) <= er) {
                pref_id = key_id(nxt);
                if ((uint32_t)lane * 4u < take)
                  cp_async_row16(&s_edges[buf ^ 1u][lane * 4], a.head + (size_t)pref_id * SEARCH_HEAD + lane * 4);
              }
            }
          }
          if (!finished && visited_n + take > a.hash_limit) overflow = true;
        }
#pragma unroll
        for (int m = 0; m < KL; m++) s_res[32 * m + lane] = res[m];
        s_front[lane] = U.front;
        if (lane == 0) {
          s_T = U.T;
          s_ctl[0] = U.fn;
          s_ctl[1] = U.qsize;
          s_ctl[2] = res_n;
          s_ctl[3] = visited_n;
          s_ctl[4] = st_dist;
          s_ctl[5] = st_edge;
          s_ctl[6] = st_exp;
          s_ctl[7] = pref_id;
          s_ctl[8] = buf;
          s_ctl[9] = (seeding ? 1u : 0u) | (seeds_taken ? 2u : 0u) | (head_round ? 4u : 0u) | 8u;
          s_ctl[10] = __float_as_uint(radius);
          s_ctl[11] = __float_as_uint(er);
          s_key_n = 0;
          s_take = take;
          s_buf = buf;
          s_seeding = seeding ? 1 : 0;
          s_er = seeding ? __int_as_float(0x7f800000) : er;
          if (overflow) s_state = 2;
          else if (finished) s_state = 1;
        }
      }
      FAST_SYNC();  // (A) the round is published
      if (s_state != 0) break;

      // ================= filter: one edge per thread (and pass) =================
      const bool seeding_round = s_seeding != 0;
      uint32_t pend_id[P];
      BucketProbe bp[P];
      uint32_t cn_w = 0, val_w = 0;
      if (W == 1 && !seeding_round && s_take > 32u) {
        // two edges per thread: the bucket reads of both passes are issued before either is looked at
        uint32_t nid2[P], b2[P], v2[P][8];
        bool valid2[P];
#pragma unroll
        for (int ps = 0; ps < P; ps++) {
          pend_id[ps] = 0;
          bp[ps].bucket = 0;
          bp[ps].slot = 0;
          const uint32_t e = (uint32_t)W * (uint32_t)(lane + 32 * ps) + (uint32_t)warp;
          nid2[ps] = e < s_take ? s_edges[s_buf][e] : 0u;
          valid2[ps] = nid2[ps] != 0u && nid2[ps] <= a.n;
          b2[ps] = (nid2[ps] * 2654435761u) >> (32 - bucket_bits);
          if (valid2[ps]) hash_load256(hash, b2[ps], v2[ps]);
        }
#pragma unroll
        for (int ps = 0; ps < P; ps++) {
          bool isnew = false;
          if (valid2[ps]) {
            isnew = !hash_lookup256_loaded(hash, bucket_bits, nid2[ps], b2[ps], v2[ps], bp[ps]);
            if (isnew) {
              pend_id[ps] = nid2[ps];
              const uint8_t *rp = a.objects + (size_t)nid2[ps] * a.row_bytes;
#pragma unroll
              for (int o = 0; o < CH; o++)
                if (o == 0 || (uint32_t)o * 128u < a.row_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + o * 128));
            }
          }
          const uint32_t m = __ballot_sync(0xffffffffu, isnew);
          const uint32_t mv = __ballot_sync(0xffffffffu, valid2[ps]);
          if (isnew) s_wids[warp][cn_w + __popc(m & lanemask_lt())] = nid2[ps];
          cn_w += (uint32_t)__popc(m);
          val_w += (uint32_t)__popc(mv);
        }
      } else
#pragma unroll
      for (int ps = 0; ps < P; ps++) {
        pend_id[ps] = 0;
        bp[ps].bucket = 0;
        bp[ps].slot = 0;
        if (ps > 0 && (uint32_t)(32 * W * ps) >= s_take) continue;
        if (ps > 0 && seeding_round) __syncwarp();   // a repeated seed id must see the insertion of the first pass
        const uint32_t e = (uint32_t)W * (uint32_t)(lane + 32 * ps) + (uint32_t)warp;   // edges dealt round-robin: the warps get equal shares
        const uint32_t nid = e < s_take ? s_edges[s_buf][e] : 0u;
        const bool valid = nid != 0u && nid <= a.n;
        bool isnew = false;
        if (valid) {
          isnew = !hash_lookup256(hash, bucket_bits, nid, bp[ps]);
          if (isnew) {
            // seed lists may repeat an id: insert at once so that the second copy is seen
            if (seeding_round) isnew = hash_insert(hash, bucket_bits, nid, bp[ps]);
            else pend_id[ps] = nid;
            // the row is copied to shared memory a few hundred cycles from now: start it towards L2
            if (isnew) {
              const uint8_t *rp = a.objects + (size_t)nid * a.row_bytes;
              // (one cp.async.bulk.prefetch.L2 per row instead of a prefetch per line measured slower: 4.22 vs 4.10 ms)
#pragma unroll
              for (int o = 0; o < CH; o++)
                if (o == 0 || (uint32_t)o * 128u < a.row_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + o * 128));
            }
          }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, isnew);
        const uint32_t mv = __ballot_sync(0xffffffffu, valid);
        if (isnew) s_wids[warp][cn_w + __popc(m & lanemask_lt())] = nid;
        cn_w += (uint32_t)__popc(m);
        val_w += (uint32_t)__popc(mv);
      }
      // the last group (four rows, or RPS) is filled up with row 0 (the all-zero dummy object, an L2 hit): its copies need no
      // predicates and its distances are dropped
      if ((uint32_t)lane < RPS - 1u) s_wids[warp][cn_w + lane] = 0u;
      if (lane == 0) {
        s_wcnt[warp] = cn_w;
        s_wval[warp] = val_w;
      }
      __syncwarp();
      const uint32_t cn = cn_w;           // this warp's rows
      const uint32_t *cand_ids = s_wids[warp];

      // ================= rows: copy and evaluate, group g of the round on warp g % 4 =================
      {
        const float er_pub = s_er;
        const uint32_t ngw = (cn + RPS - 1u) / RPS;                           // groups of RPS rows of this warp
#define FAST_GROUP_C0(t) (RPS * (t))
        // issue group t of this warp into ring slot t % NB
#define FAST_ISSUE(t)                                                                                   \
  {                                                                                                     \
    const uint32_t _t = (t);                                                                            \
    if (_t < ngw) {                                                                                     \
      const uint32_t _c0 = FAST_GROUP_C0(_t);                                                           \
      const uint32_t _dst = cp_dst + (_t % NB) * GBYTES;                                                \
      if (CH == 4) {                                                                                    \
        const uint4 _ids = *reinterpret_cast<const uint4 *>(&cand_ids[_c0]);                            \
        FAST_ROWCP(_dst, addr16(_ids.x * rb16 + cp_chunk, a.objects), cp_size);                           \
        FAST_ROWCP(_dst + SROW, addr16(_ids.y * rb16 + cp_chunk, a.objects), cp_size);                    \
        FAST_ROWCP(_dst + 2 * SROW, addr16(_ids.z * rb16 + cp_chunk, a.objects), cp_size);                \
        FAST_ROWCP(_dst + 3 * SROW, addr16(_ids.w * rb16 + cp_chunk, a.objects), cp_size);                \
      } else {                                                                                          \
        constexpr int _NI = (int)(RPS / RPI);   /* copy instructions of one group: RPI rows each */       \
        uint32_t _id[_NI];                                                                              \
        _Pragma("unroll") for (int _i = 0; _i < _NI; _i++) _id[_i] = cand_ids[_c0 + _i * RPI + cp_row]; \
        _Pragma("unroll") for (int _i = 0; _i < _NI; _i++)                                              \
          FAST_ROWCP(_dst + _i * RPI * SROW, addr16(_id[_i] * rb16 + cp_chunk, a.objects), cp_size);      \
      }                                                                                                 \
    }                                                                                                   \
    asm volatile("cp.async.commit_group;" ::: "memory");                                                \
  }
#pragma unroll
        for (int b = 0; b < NB; b++) FAST_ISSUE((uint32_t)b)
#pragma unroll
        for (int ps = 0; ps < P; ps++)
          if (pend_id[ps]) hash_insert(hash, bucket_bits, pend_id[ps], bp[ps]);   // under the row copies in flight
        float tot0 = 0.f, tot1 = 0.f;   // this lane's row of the current block of LD steps
        uint32_t totu = 0;
        for (uint32_t t = 0; t < ngw; t++) {
          cp_async_wait_group<NB - 1>();
          __syncwarp();
          const uint32_t slot = (t % NB) * GBYTES;
          if (LD == 8u) {
            const uint32_t ra = rd + slot;
            Sums p[CH];
#pragma unroll
            for (int m = 0; m < CH; m++) {
              p[m] = zero_sums();
              acc_chunk_packed<ACC>(p[m], q8[m], lds16(ra + m * 128));
              lane_total<ACC>(p[m]);
            }
            // chunks j, j + 8, j + 16, j + 24 of a row sit on lanes j, j + 8, ... of group_fold<ACC, 32>: the xor-16 and
            // xor-8 levels of its butterfly are these local adds, the remaining three levels are shuffles
            Sums tot = p[0];
            if (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM) {
              if (CH == 2) tot.u = p[0].u + p[1].u;
              if (CH == 4) tot.u = (p[0].u + p[2].u) + (p[1].u + p[3].u);
              tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 4);
              tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 2);
              tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 1);
              if ((uint32_t)(lane & 7) == (t & 7u)) totu = tot.u;
            } else {
              if (CH == 2) tot.f0 = p[0].f0 + p[1].f0;
              if (CH == 4) tot.f0 = (p[0].f0 + p[2].f0) + (p[1].f0 + p[3].f0);
              tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 4);
              tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 2);
              tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 1);
              if ((uint32_t)(lane & 7) == (t & 7u)) tot0 = tot.f0;
              if (ACC == ACC_F_COS) {
                if (CH == 2) tot.f1 = p[0].f1 + p[1].f1;
                if (CH == 4) tot.f1 = (p[0].f1 + p[2].f1) + (p[1].f1 + p[3].f1);
                tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 4);
                tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 2);
                tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 1);
                if ((uint32_t)(lane & 7) == (t & 7u)) tot1 = tot.f1;
              }
            }
          } else {
            // integer rows of <= 128 bytes, LD lanes per row: lane (rr, ll) takes chunks ll, ll + LD, ... of row rr from
            // their swizzled positions; integer sums are exact in any order
            const uint32_t base = wstage_s + slot + rr * SROW;
            const uint32_t p0 = (ll + LD * rr) & 7u;
            Sums acc = zero_sums();
#pragma unroll
            for (int m = 0; m < (int)CPL; m++) acc_chunk<ACC>(acc, q8[m], lds16(base + ((p0 + LD * m) & 7u) * 16u));
#pragma unroll
            for (uint32_t o = LD / 2; o > 0; o >>= 1) acc.u += __shfl_xor_sync(0xffffffffu, acc.u, o);
            if (ll == t % LD) totu = acc.u;
          }
          __syncwarp();
          FAST_ISSUE(t + NB)
          if (t % LD == LD - 1u || t + 1 == ngw) {
            // the scalar tail for up to 32 rows at once: lane (rr, s) owns row rr of step (t - t % LD) + s
            const uint32_t ts = (t - t % LD) + ll;
            const uint32_t j = FAST_GROUP_C0(ts) + rr;
            const bool owner = ts <= t && j < cn;
            float d = 0.f;
            if (owner) {
              Sums s = zero_sums();
              s.f0 = tot0;
              s.f1 = tot1;
              s.u = totu;
              d = finish_distance<ACC>(a.dtype, s, qn);
            }
            const bool pass = owner && d <= er_pub;
            const uint32_t pm = __ballot_sync(0xffffffffu, pass);
            if (pm) {
              uint32_t base = 0;
              if (lane == 0)
                asm volatile("atom.shared.add.u32 %0, [%1], %2;"
                             : "=r"(base)
                             : "r"((uint32_t)__cvta_generic_to_shared(&s_key_n)), "r"((uint32_t)__popc(pm))
                             : "memory");
              base = __shfl_sync(0xffffffffu, base, 0);
              if (pass) s_cand_keys[base + __popc(pm & lanemask_lt())] = make_key(d, cand_ids[j]);
            }
          }
        }
        cp_async_wait_group<0>();   // also the staged head-table row of the expected next node (warp 0)
#undef FAST_ISSUE
#undef FAST_GROUP_C0
      }
      FAST_SYNC();  // (C) keys are published; buffers may be overwritten
    }

    // ---- write the outcome
    const int state = s_state;
    if (warp == 0) {
      if (state == 1) {
        const uint32_t res_n = s_ctl[2];
#pragma unroll
        for (int m = 0; m < KL; m++) {
          const uint32_t p = (uint32_t)lane + 32u * m;   // position p: lane p / KL, slot p % KL (coalesced stores)
          if (p < a.k) {
            const uint64_t res = s_res[32 * (p % KL) + p / KL];
            const bool ok = p < res_n;
            if (a.keys_out) {
              a.keys_out[(size_t)q * a.k + p] = ok ? res + a.id_offset : KEY_NONE;
            } else {
              a.ids[(size_t)q * a.k + p] = ok ? key_id(res) : 0u;
              a.dists[(size_t)q * a.k + p] = ok ? key_dist(res) : 0.f;
            }
          }
        }
        if (lane == 0) {
          a.counts[q] = res_n;
          if (a.stats) {
            a.stats[(size_t)q * 3 + 0] = s_ctl[4];
            a.stats[(size_t)q * 3 + 1] = s_ctl[5];
            a.stats[(size_t)q * 3 + 2] = s_ctl[6];
          }
        }
      } else if (lane == 0) {
        a.counts[q] = 0xffffffffu;
        const uint32_t slot = atomicAdd(a.overflow_count, 1u);
        a.overflow_list[slot] = q;
      }
    }
    FAST_SYNC();  // s_query / s_state are rewritten by thread 0 next
  }
#undef FAST_SYNC
}

// ---- seed selection: the n_seeds nearest pivots of the seed table, one warp per query -----------------------
// Stands where GraphAndTreeIndex::getSeedsFromTree stands (lib/NGT/Index.h:1524-1567: a DVP-tree descent to one
// leaf). The table is small (hundreds of rows, L1/L2 resident), so every warp walks all of it: four rows per step,
// eight lanes per row like the traversal, exact distances, the k smallest (distance, id) keys kept sorted one per lane.
struct SeedArgs {
  const uint8_t *queries;     // prepared rows, nq x row_bytes
  const uint8_t *pivots;      // n_pivots x row_bytes
  const uint32_t *pivot_ids;
  uint32_t nq, n_pivots, row_bytes, chunks, k;
  int dtype;
  uint32_t *seeds;            // nq x k, ascending by (distance, id)
};

template <int ACC, int CH>
__global__ void __launch_bounds__(128) seed_select_kernel(const SeedArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * 4u + (threadIdx.x >> 5);
  if (q >= a.nq) return;
  const uint32_t rr = (uint32_t)lane >> 3, j = (uint32_t)lane & 7u;
  const uint8_t *qrow = a.queries + (size_t)q * a.row_bytes;
  uint4 q8[CH];
#pragma unroll
  for (int m = 0; m < CH; m++) {
    const uint32_t c = j + m * 8;
    q8[m] = c < a.chunks ? ldg16(qrow + (size_t)c * 16) : zero16();
  }
  float qn = 0.f;
  if (ACC == ACC_F_COS) {
    const uint4 v = (uint32_t)lane < a.chunks ? ldg16(qrow + (size_t)lane * 16) : zero16();
    float a0 = __uint_as_float(v.x), a1 = __uint_as_float(v.y), a2 = __uint_as_float(v.z), a3 = __uint_as_float(v.w);
    qn = fmaf(a0, a0, qn);
    qn = fmaf(a1, a1, qn);
    qn = fmaf(a2, a2, qn);
    qn = fmaf(a3, a3, qn);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
  }
  uint64_t res = KEY_NONE, thr = KEY_NONE;   // lane i: the i-th smallest key so far; thr: the k-th
#pragma unroll 2
  for (uint32_t p0 = 0; p0 < a.n_pivots; p0 += 4) {
    const uint32_t row = p0 + rr;
    const bool valid = row < a.n_pivots;
    const uint8_t *rp = a.pivots + (size_t)(valid ? row : 0u) * a.row_bytes;
    Sums p[CH];
#pragma unroll
    for (int m = 0; m < CH; m++) {
      const uint32_t c = j + m * 8;
      p[m] = zero_sums();
      acc_chunk_packed<ACC>(p[m], q8[m], c < a.chunks ? ldg16(rp + (size_t)c * 16) : zero16());
      lane_total<ACC>(p[m]);
    }
    Sums tot = p[0];
    if (ACC == ACC_U8_L2 || ACC == ACC_U8_HAM) {
      if (CH == 2) tot.u = p[0].u + p[1].u;
      if (CH == 4) tot.u = (p[0].u + p[2].u) + (p[1].u + p[3].u);
      tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 4);
      tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 2);
      tot.u += __shfl_xor_sync(0xffffffffu, tot.u, 1);
    } else {
      if (CH == 2) tot.f0 = p[0].f0 + p[1].f0;
      if (CH == 4) tot.f0 = (p[0].f0 + p[2].f0) + (p[1].f0 + p[3].f0);
      tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 4);
      tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 2);
      tot.f0 += __shfl_xor_sync(0xffffffffu, tot.f0, 1);
      if (ACC == ACC_F_COS) {
        if (CH == 2) tot.f1 = p[0].f1 + p[1].f1;
        if (CH == 4) tot.f1 = (p[0].f1 + p[2].f1) + (p[1].f1 + p[3].f1);
        tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 4);
        tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 2);
        tot.f1 += __shfl_xor_sync(0xffffffffu, tot.f1, 1);
      }
    }
    uint64_t key = KEY_NONE;
    if (j == 0 && valid) key = make_key(finish_distance<ACC>(a.dtype, tot, qn), __ldg(a.pivot_ids + row));
    uint32_t m = __ballot_sync(0xffffffffu, key < thr);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t kk = shfl_u64(key, src);
      if (kk >= thr) continue;
      const uint32_t pos = __popc(__ballot_sync(0xffffffffu, res < kk));
      const uint64_t up = shfl_up_u64(res, 1);
      if ((uint32_t)lane == pos) res = kk;
      else if ((uint32_t)lane > pos) res = up;
      if ((uint32_t)lane >= a.k) res = KEY_NONE;
      thr = shfl_u64(res, (int)a.k - 1);
    }
  }
  if ((uint32_t)lane < a.k) a.seeds[(size_t)q * a.k + lane] = res != KEY_NONE ? key_id(res) : 0u;
}

template <int ACC, int CH>
static cudaError_t seed_one(const SeedArgs &a, cudaStream_t stream) {
  seed_select_kernel<ACC, CH><<<(a.nq + 3) / 4, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

template <int ACC>
cudaError_t seed_select_dispatch(const SeedArgs &a, cudaStream_t stream) {
  const int ch = a.chunks <= 8 ? 1 : a.chunks <= 16 ? 2 : 4;
  if (ch == 1) return seed_one<ACC, 1>(a, stream);
  if (ch == 2) return seed_one<ACC, 2>(a, stream);
  return seed_one<ACC, 4>(a, stream);
}

// op == 0: launch, op == 1: occupancy query
template <int ACC, int CH, int W, int KL = 1>
static cudaError_t fast_one(const SearchArgs &a, unsigned grid, size_t smem, cudaStream_t stream, int op, int *blocks) {
  cudaError_t e = cudaFuncSetAttribute(search_fast_kernel<ACC, CH, W, KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(search_fast_kernel<ACC, CH, W, KL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return e;
  if (op == 1) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, search_fast_kernel<ACC, CH, W, KL>, W * 32, smem);
  search_fast_kernel<ACC, CH, W, KL><<<grid, W * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

// ch: chunks per lane of a row's eight lanes (1, 2, 4); warps per query: 4 or 2 (1 when asked for); result lists of
// 33..128 keys (four per lane of the control warp) come with four warps per query
template <int ACC>
cudaError_t search_fast_dispatch(const SearchArgs &a, int ch, int warps, unsigned grid, size_t smem, cudaStream_t stream, int op,
                                 int *blocks) {
  if (a.k > 32u * FAST_KL_MAX) return cudaErrorInvalidValue;
  if (a.k > 32u) {
    if (warps != FAST_WARPS) return cudaErrorInvalidValue;
    if (ch == 1) return fast_one<ACC, 1, 4, FAST_KL_MAX>(a, grid, smem, stream, op, blocks);
    if (ch == 2) return fast_one<ACC, 2, 4, FAST_KL_MAX>(a, grid, smem, stream, op, blocks);
    if (ch == 4) return fast_one<ACC, 4, 4, FAST_KL_MAX>(a, grid, smem, stream, op, blocks);
    return cudaErrorInvalidValue;
  }
  if (ch == 1 && warps == 1) return fast_one<ACC, 1, 1>(a, grid, smem, stream, op, blocks);
  if (ch == 1 && warps == 2) return fast_one<ACC, 1, 2>(a, grid, smem, stream, op, blocks);
  if (ch == 2 && warps == 2) return fast_one<ACC, 2, 2>(a, grid, smem, stream, op, blocks);
  if (ch == 4 && warps == 2) return fast_one<ACC, 4, 2>(a, grid, smem, stream, op, blocks);
  if (ch == 1) return fast_one<ACC, 1, 4>(a, grid, smem, stream, op, blocks);
  if (ch == 2) return fast_one<ACC, 2, 4>(a, grid, smem, stream, op, blocks);
  if (ch == 4) return fast_one<ACC, 4, 4>(a, grid, smem, stream, op, blocks);
  return cudaErrorInvalidValue;
}

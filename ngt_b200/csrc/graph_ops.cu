// graph_ops.cu -- adjacency-list surgery of ONNG construction on the device (no distance arithmetic here).
//
// ngtgpu_graph_adjust_paths restates GraphReconstructor::adjustPathsEffectively
// (lib/NGT/GraphReconstructor.h:197-386), the "shortcut reduction" step GraphOptimizer::execute runs after
// reconstructGraph (lib/NGT/GraphOptimizer.h:279-292): an edge src->dst is dropped when a two-hop path
// src->path->dst with both hops shorter than the edge already exists in the graph being rebuilt.
//
// The reference rebuilds the graph edge by edge in the order (rank of the edge in its list, source id):
//   for rank = 0, 1, ...: for id = 1..n: edge e = (id, rank) is inserted unless
//       (edges kept so far of id) + (edges of id not yet looked at) > minNoOfEdges          (:318)
//       and some candidate (path, dst) of e has  id->path  and  path->dst  in the rebuilt graph  (:330)
// where the candidates of e = (id -> dst) are the pairs with path in N(id), dst in N(path),
// d(id,path) < d(id,dst) and d(path,dst) < d(id,dst) (:255-272). Whether path->dst "is in the rebuilt graph"
// at that moment depends only on its own decision and on its position in the same order:
// it was looked at earlier iff (rank of dst in N(path), path) < (rank, id) lexicographically. id->path always was
// (its distance is smaller, so its rank is). Hence:
//   phase 1  one CTA per source node enumerates N(path) for every path in N(src) (coalesced reads), finds dst in
//            N(src) by binary search in a shared-memory copy sorted by id, and records for edge (src, rank of dst)
//            the pair (rank of path, edge index of path->dst) -- but only if path->dst can precede the edge at all;
//            pairs of the same sweep (equal rank, smaller id) are flagged. Two passes: count, prefix sum, fill.
//   phase 2  one launch (or a few) per rank: every edge of that rank whose same-sweep dependencies are settled
//            decides itself; edges waiting for a lower id of the same sweep are retried until none is left
//            (the lowest undecided id never waits, so this terminates).
// The result is the reference's graph bit for bit (tests/golden/adjust_paths.npz pins it).
#include <cub/cub.cuh>
#include <mutex>
#include <atomic>
#include <vector>

#include "ngtgpu_internal.cuh"

namespace {

constexpr int AP_THREADS = 128;
constexpr uint32_t AP_SMEM_DEG = 2048;   // lists up to this length are staged in shared memory

struct Cand {
  uint32_t p_flag;   // rank of `path` in N(src) | same-sweep flag << 31
  uint32_t e2;       // edge index of path -> dst
};

__global__ void edge_source_kernel(const uint64_t *__restrict__ row_ptr, uint64_t n, uint32_t *__restrict__ esrc,
                                   uint32_t *__restrict__ erank, uint32_t *__restrict__ deg_hist, uint32_t hist_cap) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t id = warp; id <= n; id += warps) {
    const uint64_t b = row_ptr[id], e = row_ptr[id + 1];
    for (uint64_t i = b + lane; i < e; i += 32) {
      esrc[i] = (uint32_t)id;
      erank[i] = (uint32_t)(i - b);
    }
    if (lane == 0) {
      const uint64_t d = e - b;
      atomicAdd(&deg_hist[d < hist_cap ? (uint32_t)d : hist_cap], 1u);
    }
  }
}

__global__ void make_sort_keys_kernel(const uint32_t *__restrict__ esrc, const uint32_t *__restrict__ col, uint64_t nnz,
                                      uint64_t *__restrict__ keys) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * blockDim.x)
    keys[i] = ((uint64_t)esrc[i] << 32) | col[i];
}

__global__ void low_words_kernel(const uint64_t *__restrict__ keys, uint64_t nnz, uint32_t *__restrict__ out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = (uint32_t)keys[i];
}

__global__ void degree_kernel(const uint64_t *__restrict__ row_ptr, uint64_t n, uint32_t *__restrict__ deg,
                              uint32_t *__restrict__ node) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
    deg[i] = (uint32_t)(row_ptr[i + 1] - row_ptr[i]);
    node[i] = (uint32_t)i;
  }
}

// position of `id` in the ascending array s[0..len), or len when absent
__device__ __forceinline__ uint32_t find_id(const uint32_t *s, uint32_t len, uint32_t id) {
  uint32_t lo = 0, hi = len;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (s[mid] < id) lo = mid + 1;
    else hi = mid;
  }
  return (lo < len && s[lo] == id) ? lo : len;
}

// Phase 1. FILL == false: cand_count[e] = number of candidates of edge e. FILL == true: write them at cand_ptr[e].
template <bool FILL>
__global__ void __launch_bounds__(AP_THREADS) candidates_kernel(const uint64_t *__restrict__ row_ptr,
                                                                const uint32_t *__restrict__ col,
                                                                const float *__restrict__ dist,
                                                                const uint32_t *__restrict__ sid,     // col sorted by id per row
                                                                const uint32_t *__restrict__ srank,   // rank of each sorted entry
                                                                uint64_t n, uint32_t *__restrict__ cand_count,
                                                                const uint64_t *__restrict__ cand_ptr,
                                                                Cand *__restrict__ cands, uint32_t *__restrict__ work) {
  __shared__ uint32_t s_sid[AP_SMEM_DEG];
  __shared__ uint32_t s_srank[AP_SMEM_DEG];
  __shared__ float s_dist[AP_SMEM_DEG];
  __shared__ uint32_t s_cnt[AP_SMEM_DEG];
  __shared__ uint32_t s_src;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    if (tid == 0) s_src = atomicAdd(work, 1u) + 1u;   // ids 1..n
    __syncthreads();
    const uint32_t src = s_src;
    if (src > n) break;
    const uint64_t sb = row_ptr[src];
    const uint32_t sdeg = (uint32_t)(row_ptr[src + 1] - sb);
    const bool in_smem = sdeg <= AP_SMEM_DEG;
    const uint32_t *q_sid = sid + sb, *q_srank = srank + sb;
    const float *q_dist = dist + sb;
    uint32_t *q_cnt = cand_count + sb;
    if (in_smem) {
      for (uint32_t i = tid; i < sdeg; i += AP_THREADS) {
        s_sid[i] = sid[sb + i];
        s_srank[i] = srank[sb + i];
        s_dist[i] = dist[sb + i];
        s_cnt[i] = 0;
      }
      q_sid = s_sid;
      q_srank = s_srank;
      q_dist = s_dist;
      q_cnt = s_cnt;
    } else if (FILL) {
      for (uint32_t i = tid; i < sdeg; i += AP_THREADS) q_cnt[i] = 0;   // reused as the fill cursor
    }
    __syncthreads();
    for (uint32_t sni = warp; sni < sdeg; sni += AP_THREADS / 32) {
      const uint32_t path = col[sb + sni];
      const float d1 = in_smem ? s_dist[sni] : dist[sb + sni];
      const uint64_t pb = row_ptr[path];
      const uint32_t pdeg = (uint32_t)(row_ptr[(uint64_t)path + 1] - pb);
      for (uint32_t pni = lane; pni < pdeg; pni += 32) {
        const uint32_t dst = col[pb + pni];
        const float d2 = dist[pb + pni];
        const uint32_t pos = find_id(q_sid, sdeg, dst);
        if (pos == sdeg) continue;
        const uint32_t r = q_srank[pos];
        const float d = q_dist[r];
        if (!(d1 < d && d2 < d)) continue;                       // GraphReconstructor.h:255-257
        // can path->dst be in the rebuilt graph when (src, r) is looked at?
        const bool earlier = pni < r;
        const bool same_sweep = pni == r && path < src;
        if (!earlier && !same_sweep) continue;
        const uint32_t slot = atomicAdd(&q_cnt[r], 1u);
        if (FILL) {
          Cand c;
          c.p_flag = sni | (same_sweep ? 0x80000000u : 0u);
          c.e2 = (uint32_t)(pb + pni);
          cands[cand_ptr[sb + r] + slot] = c;
        }
      }
    }
    __syncthreads();
    if (!FILL && in_smem)
      for (uint32_t i = tid; i < sdeg; i += AP_THREADS) cand_count[sb + i] = s_cnt[i];
    __syncthreads();
  }
}

__device__ __forceinline__ uint8_t load_status(const uint8_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return (uint8_t)v;
}

// Phase 2, one sweep: the edges of rank `rank` of the first n_active nodes of `order` (degree > rank).
// status: 0 undecided, 1 kept, 2 removed.
__global__ void sweep_kernel(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ order, uint32_t n_active,
                             uint32_t rank, const uint32_t *__restrict__ cand_count, const uint64_t *__restrict__ cand_ptr,
                             const Cand *__restrict__ cands, uint8_t *status, uint32_t *kept_count, uint32_t min_edges,
                             uint32_t *pending, unsigned long long *removed) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_active) return;
  const uint32_t src = order[i];
  const uint64_t sb = row_ptr[src];
  const uint32_t sdeg = (uint32_t)(row_ptr[(uint64_t)src + 1] - sb);
  const uint64_t e = sb + rank;
  if (load_status(status + e) != 0) return;
  const uint32_t nc = cand_count[e];
  bool remove = false, wait = false;
  // (edges kept so far) + (edges not yet looked at, this one included) > minNoOfEdges   GraphReconstructor.h:318
  if (nc != 0 && (uint64_t)kept_count[src] + sdeg - rank > min_edges) {
    for (int attempt = 0; attempt < 4 && !remove; attempt++) {
      wait = false;
      const Cand *c = cands + cand_ptr[e];
      for (uint32_t j = 0; j < nc; j++) {
        const Cand cj = c[j];
        if (status[sb + (cj.p_flag & 0x7fffffffu)] != 1) continue;   // src->path was removed (decided in an earlier sweep)
        const uint8_t s2 = load_status(status + cj.e2);
        if (s2 == 1) {
          remove = true;
          break;
        }
        if (s2 == 0 && (cj.p_flag & 0x80000000u)) wait = true;        // a lower id of this sweep has not decided yet
      }
      if (!wait) break;
    }
  }
  if (remove) {
    status[e] = 2;
    atomicAdd(removed, 1ull);
  } else if (wait) {
    *pending = 1;
  } else {
    kept_count[src] += 1;
    __threadfence();
    status[e] = 1;
  }
}

__global__ void keep_mask_kernel(const uint8_t *__restrict__ status, uint64_t nnz, uint8_t *__restrict__ keep) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * blockDim.x)
    keep[i] = status[i] == 1 ? 1 : 0;
}

// Temporary device buffers of one call: handed back, whichever way the function returns, to a small per-device cache of
// blocks instead of cudaFree -- the construction loop calls these functions once per batch of 200 objects, and a dozen
// cudaMalloc / cudaFree pairs per call (each a driver round trip with an implicit device synchronisation) were most of
// its time (19 ms per batch of which ~2 ms on the GPU). A block goes back to the cache only after the device is idle,
// which is the guarantee cudaFree gave.
std::atomic<uint64_t> g_batches_merged{0}, g_batches_sorted{0}, g_blocks_malloc{0}, g_blocks_cached{0};

struct BlockCache {
  struct Block {
    void *p;
    size_t bytes;
    int dev;
  };
  std::mutex mu;
  std::vector<Block> free_blocks;
  size_t cached_bytes = 0;
  static constexpr size_t kMaxCachedBytes = (size_t)8 << 30;
  static constexpr size_t kMaxBlocks = 96;
  void *take(size_t bytes, int dev, size_t *got) {
    std::lock_guard<std::mutex> lock(mu);
    int best = -1;
    for (size_t i = 0; i < free_blocks.size(); i++) {
      const Block &b = free_blocks[i];
      if (b.dev != dev || b.bytes < bytes || b.bytes > 2 * bytes + ((size_t)1 << 20)) continue;
      if (best < 0 || b.bytes < free_blocks[best].bytes) best = (int)i;
    }
    if (best < 0) return nullptr;
    Block b = free_blocks[best];
    free_blocks.erase(free_blocks.begin() + best);
    cached_bytes -= b.bytes;
    *got = b.bytes;
    return b.p;
  }
  void give(void *p, size_t bytes, int dev) {
    std::lock_guard<std::mutex> lock(mu);
    free_blocks.push_back(Block{p, bytes, dev});
    cached_bytes += bytes;
    while (!free_blocks.empty() && (cached_bytes > kMaxCachedBytes || free_blocks.size() > kMaxBlocks)) {   // oldest first
      cudaFree(free_blocks.front().p);
      cached_bytes -= free_blocks.front().bytes;
      free_blocks.erase(free_blocks.begin());
    }
  }
  // everything back to the driver (a failed cudaMalloc retries after this)
  void trim() {
    std::lock_guard<std::mutex> lock(mu);
    for (Block &b : free_blocks) cudaFree(b.p);
    free_blocks.clear();
    cached_bytes = 0;
  }
};
BlockCache &block_cache() {
  static BlockCache *c = new BlockCache();   // leaked on purpose: no cudaFree after the driver has shut down
  return *c;
}

struct DeviceBuffers {
  struct Held {
    void *p;
    size_t bytes;
  };
  std::vector<Held> held;
  int dev = -1;
  cudaStream_t stream;   // the caller's stream: blocks are zeroed on it before they are handed out
  explicit DeviceBuffers(cudaStream_t s) : stream(s) {}
  ~DeviceBuffers() {
    if (held.empty()) return;
    cudaDeviceSynchronize();
    for (Held &h : held) block_cache().give(h.p, h.bytes, dev);
  }
  template <typename T>
  cudaError_t alloc(T **out, size_t count) {
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    size_t bytes = (count ? count * sizeof(T) : sizeof(T));
    bytes = (bytes + 511) & ~(size_t)511;
    size_t got = 0;
    void *p = block_cache().take(bytes, dev, &got);
    cudaError_t e = cudaSuccess;
    if (p) g_blocks_cached++;
    else g_blocks_malloc++;
    if (!p) {
      got = bytes;
      e = cudaMalloc(&p, bytes);
      if (e != cudaSuccess) {   // the cache may be what fills the device
        cudaGetLastError();
        cudaDeviceSynchronize();
        block_cache().trim();
        e = cudaMalloc(&p, bytes);
      }
    }
    if (e == cudaSuccess) {
      held.push_back(Held{p, got});
      // (development: NGTGPU_POISON=1 fills with 0xff instead, which shows every dependence on that)
      static const bool poison = getenv("NGTGPU_POISON") != nullptr;
      // a block of the cache holds what its last user left, and a fresh one is zero only by the driver's habit: every
      // block starts zeroed (the callers' counters and sparse tables count on it)
      e = cudaMemsetAsync(p, poison ? 0xff : 0, got, stream);
    } else {
      p = nullptr;
    }
    *out = static_cast<T *>(p);
    return e;
  }
  void release(void *p) {
    for (size_t i = 0; i < held.size(); i++)
      if (held[i].p == p) {
        cudaDeviceSynchronize();
        block_cache().give(p, held[i].bytes, dev);
        held.erase(held.begin() + i);
        return;
      }
  }
};

}  // namespace

extern "C" int ngtgpu_graph_adjust_paths(uint64_t n, const uint64_t *d_row_ptr, const uint32_t *d_col, const float *d_dist,
                                         uint32_t min_edges, uint8_t *d_keep, uint64_t *stats, void *stream_) {
  if (!d_row_ptr || !d_keep) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_adjust_paths: null buffer");
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  uint64_t nnz = 0;
  CUDA_TRY(cudaMemcpyAsync(&nnz, d_row_ptr + n + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
  if (nnz == 0) return NGTGPU_OK;
  if (!d_col || !d_dist) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_adjust_paths: null buffer");
  if (nnz >= (1ull << 31) || n >= 0xfffffffeull)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_adjust_paths: more than 2^31 edges are not supported");
  const unsigned grid = (unsigned)sms * 8;
  uint64_t launches = 0;
  DeviceBuffers mem(stream);

  // ---- edge -> (source, rank); degree histogram (for the number of active nodes per sweep)
  const uint32_t hist_cap = 1u << 20;
  uint32_t *esrc, *erank, *hist;
  CUDA_TRY(mem.alloc(&esrc, nnz));
  CUDA_TRY(mem.alloc(&erank, nnz));
  CUDA_TRY(mem.alloc(&hist, (size_t)hist_cap + 1));
  CUDA_TRY(cudaMemsetAsync(hist, 0, ((size_t)hist_cap + 1) * 4, stream));
  edge_source_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, n, esrc, erank, hist, hist_cap);
  launches++;
  std::vector<uint32_t> h_hist((size_t)hist_cap + 1);
  CUDA_TRY(cudaMemcpyAsync(h_hist.data(), hist, h_hist.size() * 4, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (h_hist[hist_cap]) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_adjust_paths: a node has 2^20 edges or more");
  uint32_t max_deg = 0;
  for (uint32_t d = 0; d < hist_cap; d++)
    if (h_hist[d]) max_deg = d;
  mem.release(hist);

  // ---- every list once more, sorted by id (with the rank each entry has in the distance order)
  uint32_t *sid, *srank;
  CUDA_TRY(mem.alloc(&sid, nnz));
  CUDA_TRY(mem.alloc(&srank, nnz));
  {
    uint64_t *k_in, *k_out;
    CUDA_TRY(mem.alloc(&k_in, nnz));
    CUDA_TRY(mem.alloc(&k_out, nnz));
    make_sort_keys_kernel<<<grid, 256, 0, stream>>>(esrc, d_col, nnz, k_in);
    int hi_bits = 1;
    while (hi_bits < 32 && (n >> hi_bits) != 0) hi_bits++;
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, erank, srank, (int64_t)nnz, 0, 32 + hi_bits,
                                             stream));
    uint8_t *tmp;
    CUDA_TRY(mem.alloc(&tmp, tmp_bytes));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, erank, srank, (int64_t)nnz, 0, 32 + hi_bits, stream));
    low_words_kernel<<<grid, 256, 0, stream>>>(k_out, nnz, sid);
    launches += 4;
    CUDA_TRY(cudaStreamSynchronize(stream));
    mem.release(tmp);
    mem.release(k_in);
    mem.release(k_out);
  }
  mem.release(esrc);
  mem.release(erank);

  // ---- phase 1: candidates per edge (count, prefix sum, fill)
  uint32_t *cand_count, *work;
  uint64_t *cand_ptr;
  CUDA_TRY(mem.alloc(&cand_count, nnz + 1));
  CUDA_TRY(mem.alloc(&cand_ptr, nnz + 1));
  CUDA_TRY(mem.alloc(&work, 4));
  CUDA_TRY(cudaMemsetAsync(cand_count, 0, (nnz + 1) * 4, stream));
  CUDA_TRY(cudaMemsetAsync(work, 0, 16, stream));
  candidates_kernel<false><<<grid, AP_THREADS, 0, stream>>>(d_row_ptr, d_col, d_dist, sid, srank, n, cand_count, nullptr,
                                                            nullptr, work);
  launches++;
  {
    size_t tmp_bytes = 0;
    // (64-bit sums of 32-bit counts: the initial value's type drives the accumulator)
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, tmp_bytes, cand_count, cand_ptr, cub::Sum(), (uint64_t)0,
                                            (int64_t)nnz + 1, stream));
    uint8_t *tmp;
    CUDA_TRY(mem.alloc(&tmp, tmp_bytes));
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, cand_count, cand_ptr, cub::Sum(), (uint64_t)0, (int64_t)nnz + 1,
                                            stream));
    launches++;
    CUDA_TRY(cudaStreamSynchronize(stream));
    mem.release(tmp);
  }
  uint64_t n_cand = 0;
  CUDA_TRY(cudaMemcpyAsync(&n_cand, cand_ptr + nnz, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  Cand *cands;
  CUDA_TRY(mem.alloc(&cands, (size_t)n_cand));
  CUDA_TRY(cudaMemsetAsync(work, 0, 16, stream));
  candidates_kernel<true><<<grid, AP_THREADS, 0, stream>>>(d_row_ptr, d_col, d_dist, sid, srank, n, cand_count, cand_ptr,
                                                           cands, work);
  launches++;
  CUDA_TRY(cudaStreamSynchronize(stream));
  mem.release(sid);
  mem.release(srank);
  // the long lists (> AP_SMEM_DEG) used cand_count as their fill cursor: it ends equal to the count again

  // ---- phase 2: sweeps over ranks; nodes ordered by degree (descending) so the active ones are a prefix
  uint32_t *deg, *node, *deg_sorted, *order;
  CUDA_TRY(mem.alloc(&deg, n + 1));
  CUDA_TRY(mem.alloc(&node, n + 1));
  CUDA_TRY(mem.alloc(&deg_sorted, n + 1));
  CUDA_TRY(mem.alloc(&order, n + 1));
  degree_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, n, deg, node);
  {
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, deg, deg_sorted, node, order, (int64_t)n + 1, 0, 21,
                                                       stream));
    uint8_t *tmp;
    CUDA_TRY(mem.alloc(&tmp, tmp_bytes));
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, deg, deg_sorted, node, order, (int64_t)n + 1, 0, 21,
                                                       stream));
    launches += 2;
    CUDA_TRY(cudaStreamSynchronize(stream));
    mem.release(tmp);
  }
  uint8_t *status;
  uint32_t *kept_count, *pending;
  unsigned long long *removed;
  CUDA_TRY(mem.alloc(&status, nnz));
  CUDA_TRY(mem.alloc(&kept_count, n + 1));
  CUDA_TRY(mem.alloc(&pending, 4));
  CUDA_TRY(mem.alloc(&removed, 2));
  CUDA_TRY(cudaMemsetAsync(status, 0, nnz, stream));
  CUDA_TRY(cudaMemsetAsync(kept_count, 0, (n + 1) * 4, stream));
  CUDA_TRY(cudaMemsetAsync(removed, 0, 16, stream));
  // active[r] = number of nodes with degree > r
  std::vector<uint32_t> active(max_deg + 1, 0);
  {
    uint64_t acc = 0;
    for (uint32_t d = max_deg; d >= 1; d--) {
      acc += h_hist[d];
      active[d - 1] = (uint32_t)acc;
    }
  }
  uint32_t *h_pending = nullptr;
  CUDA_TRY(cudaMallocHost(&h_pending, 4));
  uint64_t sweeps = 0;
  int rc = NGTGPU_OK;
  for (uint32_t r = 0; r < max_deg && rc == NGTGPU_OK; r++) {
    const uint32_t na = active[r];
    for (uint64_t iter = 0;; iter++) {
      cudaMemsetAsync(pending, 0, 4, stream);
      sweep_kernel<<<(na + 255) / 256, 256, 0, stream>>>(d_row_ptr, order, na, r, cand_count, cand_ptr, cands, status,
                                                         kept_count, min_edges, pending, removed);
      launches++;
      sweeps++;
      cudaMemcpyAsync(h_pending, pending, 4, cudaMemcpyDeviceToHost, stream);
      cudaError_t e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) {
        ngtgpu_set_error(std::string("ngtgpu_graph_adjust_paths: sweep: ") + cudaGetErrorString(e));
        rc = NGTGPU_ERR_CUDA;
        break;
      }
      if (*h_pending == 0) break;
      if (iter > (uint64_t)na + 8) {
        ngtgpu_set_error("ngtgpu_graph_adjust_paths: a sweep did not settle");
        rc = NGTGPU_ERR_STATE;
        break;
      }
    }
  }
  cudaFreeHost(h_pending);
  if (rc != NGTGPU_OK) return rc;
  keep_mask_kernel<<<grid, 256, 0, stream>>>(status, nnz, d_keep);
  launches++;
  unsigned long long h_removed = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_removed, removed, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  CUDA_TRY(cudaGetLastError());
  if (stats) {
    stats[0] = n_cand;
    stats[1] = h_removed;
    stats[2] = sweeps;
    stats[3] = launches;
  }
  return NGTGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ngtgpu_graph_reconstruct -- GraphReconstructor::reconstructGraph (lib/NGT/GraphReconstructor.h:425-561), the first
// step of ONNG construction: node i keeps its first `outgoing` edges (all of them when it has fewer, :447-456), every
// node then receives the reverse of the first `incoming` edges of every other node, lists are sorted by
// (distance, id) and an entry whose id repeats the previous entry's id is dropped (:519-533).
// On the device: emit (source, distance, target) triples, two stable radix sorts (by (distance, target), then by
// source), flag the survivors, compact.
namespace {

__global__ void emit_edges_kernel(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                  const float *__restrict__ dist, uint64_t n, uint32_t outgoing, uint32_t incoming,
                                  uint32_t *__restrict__ out_src, uint64_t *__restrict__ out_key,
                                  unsigned long long *__restrict__ counter) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t id = warp; id <= n; id += warps) {
    const uint64_t b = row_ptr[id];
    const uint32_t deg = (uint32_t)(row_ptr[id + 1] - b);
    const bool keep_all = deg < outgoing;
    for (uint32_t r0 = 0; r0 < deg; r0 += 32) {
      const uint32_t r = r0 + lane;
      const bool have = r < deg;
      const uint32_t t = have ? col[b + r] : 0u;
      const uint32_t dord = have ? ord_of_float(dist[b + r]) : 0u;
      const bool fwd = have && outgoing > 0 && (r < outgoing || keep_all);
      const bool rev = have && r < incoming;
      const uint32_t mf = __ballot_sync(0xffffffffu, fwd), mr = __ballot_sync(0xffffffffu, rev);
      unsigned long long base = 0;
      if (lane == 0 && (mf | mr)) base = atomicAdd(counter, (unsigned long long)(__popc(mf) + __popc(mr)));
      base = shfl_u64(base, 0);
      const uint32_t below = (1u << lane) - 1u;
      if (fwd) {
        const uint64_t p = base + __popc(mf & below);
        out_src[p] = (uint32_t)id;
        out_key[p] = ((uint64_t)dord << 32) | t;
      }
      if (rev) {
        const uint64_t p = base + __popc(mf) + __popc(mr & below);
        out_src[p] = t;
        out_key[p] = ((uint64_t)dord << 32) | (uint32_t)id;
      }
    }
  }
}

__global__ void flag_unique_kernel(const uint32_t *__restrict__ src, const uint64_t *__restrict__ key, uint64_t m,
                                   uint8_t *__restrict__ flag, uint32_t *__restrict__ deg) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    const bool keep = i == 0 || src[i] != src[i - 1] || (uint32_t)key[i] != (uint32_t)key[i - 1];
    flag[i] = keep ? 1 : 0;
    if (keep) atomicAdd(&deg[src[i]], 1u);
  }
}

__global__ void unpack_edges_kernel(const uint64_t *__restrict__ key, uint64_t m, uint32_t *__restrict__ col,
                                    float *__restrict__ dist) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    col[i] = (uint32_t)key[i];
    dist[i] = float_of_ord((uint32_t)(key[i] >> 32));
  }
}

}  // namespace

namespace {

// Triples (source, (distance << 32 | target)) in src_a / key_a [m] -> CSR sorted by (source, distance, target) with
// repeated (source, target) neighbours dropped. src_b / key_b are scratch of the same size; counter: 2 device words.
int csr_from_triples(DeviceBuffers &mem, uint64_t n, uint32_t *src_a, uint32_t *src_b, uint64_t *key_a, uint64_t *key_b,
                     unsigned long long *counter, uint64_t m, uint64_t capacity, uint64_t *d_out_row_ptr,
                     uint32_t *d_out_col, float *d_out_dist, uint64_t *out_nnz, unsigned grid, cudaStream_t stream) {
  int hi_bits = 1;
  while (hi_bits < 32 && (n >> hi_bits) != 0) hi_bits++;
  {
    size_t t1 = 0, t2 = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, key_a, key_b, src_a, src_b, (int64_t)m, 0, 64, stream));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t2, src_b, src_a, key_b, key_a, (int64_t)m, 0, hi_bits, stream));
    uint8_t *tmp;
    CUDA_TRY(mem.alloc(&tmp, t1 > t2 ? t1 : t2));
    // by (distance, target) ...
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, t1, key_a, key_b, src_a, src_b, (int64_t)m, 0, 64, stream));
    // ... then, stably, by source
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, t2, src_b, src_a, key_b, key_a, (int64_t)m, 0, hi_bits, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    mem.release(tmp);
  }
  uint8_t *flag;
  uint32_t *deg;
  CUDA_TRY(mem.alloc(&flag, (size_t)m));
  CUDA_TRY(mem.alloc(&deg, n + 2));
  CUDA_TRY(cudaMemsetAsync(deg, 0, (n + 2) * 4, stream));
  CUDA_TRY(cudaMemsetAsync(d_out_row_ptr, 0, sizeof(uint64_t), stream));
  flag_unique_kernel<<<grid, 256, 0, stream>>>(src_a, key_a, m, flag, deg);
  // surviving keys -> key_b, their number -> counter; row_ptr[1 + id] = sum of deg[0..id]
  size_t t1 = 0, t2 = 0;
  CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, t1, key_a, flag, key_b, counter, (int64_t)m, stream));
  CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, t2, deg, d_out_row_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));
  uint8_t *tmp;
  CUDA_TRY(mem.alloc(&tmp, t1 > t2 ? t1 : t2));
  CUDA_TRY(cub::DeviceSelect::Flagged(tmp, t1, key_a, flag, key_b, counter, (int64_t)m, stream));
  CUDA_TRY(cub::DeviceScan::InclusiveScan(tmp, t2, deg, d_out_row_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));
  unsigned long long kept = 0;
  CUDA_TRY(cudaMemcpyAsync(&kept, counter, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  mem.release(tmp);
  mem.release(flag);
  mem.release(deg);
  if (kept > capacity)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "graph output capacity " + std::to_string(capacity) + " < " + std::to_string(kept) + " edges");
  unpack_edges_kernel<<<grid, 256, 0, stream>>>(key_b, kept, d_out_col, d_out_dist);
  CUDA_TRY(cudaStreamSynchronize(stream));
  CUDA_TRY(cudaGetLastError());
  *out_nnz = kept;
  return NGTGPU_OK;
}

}  // namespace

extern "C" int ngtgpu_graph_reconstruct(uint64_t n, const uint64_t *d_row_ptr, const uint32_t *d_col, const float *d_dist,
                                        uint32_t outgoing, uint32_t incoming, uint64_t capacity, uint64_t *d_out_row_ptr,
                                        uint32_t *d_out_col, float *d_out_dist, uint64_t *out_nnz, void *stream_) {
  if (!d_row_ptr || !d_out_row_ptr || !out_nnz) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_reconstruct: null buffer");
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)sms * 8;
  uint64_t nnz = 0;
  CUDA_TRY(cudaMemcpyAsync(&nnz, d_row_ptr + n + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  *out_nnz = 0;
  CUDA_TRY(cudaMemsetAsync(d_out_row_ptr, 0, (n + 2) * sizeof(uint64_t), stream));
  if (nnz == 0) return NGTGPU_OK;
  if (!d_col || !d_dist || !d_out_col || !d_out_dist) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_reconstruct: null buffer");
  if (n >= 0xfffffffeull) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_reconstruct: too many nodes");
  DeviceBuffers mem(stream);
  uint32_t *src_a, *src_b;
  uint64_t *key_a, *key_b;
  unsigned long long *counter;
  const uint64_t cap2 = 2 * nnz;
  CUDA_TRY(mem.alloc(&src_a, cap2));
  CUDA_TRY(mem.alloc(&src_b, cap2));
  CUDA_TRY(mem.alloc(&key_a, cap2));
  CUDA_TRY(mem.alloc(&key_b, cap2));
  CUDA_TRY(mem.alloc(&counter, 2));
  CUDA_TRY(cudaMemsetAsync(counter, 0, 16, stream));
  emit_edges_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, d_col, d_dist, n, outgoing, incoming, src_a, key_a, counter);
  unsigned long long m = 0;
  CUDA_TRY(cudaMemcpyAsync(&m, counter, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (m == 0) return NGTGPU_OK;
  return csr_from_triples(mem, n, src_a, src_b, key_a, key_b, counter, m, capacity, d_out_row_ptr, d_out_col, d_out_dist,
                          out_nnz, grid, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// ngtgpu_index_refine_anng -- GraphReconstructor::refineANNG (lib/NGT/GraphReconstructor.h:814-924; C API
// ngt_refine_anng): batch by batch, every object is searched for in the current graph (size = searched_edges, the given
// epsilon / edge size), the results (the object itself excluded) are merged into its list (sort, repeated ids
// dropped, :869-888), and -- unless a kNN graph was asked for (no_of_edges != 0) -- every result gets the reverse edge
// (addEdge without identity check, :893-901). The next batch searches the refined graph. With no_of_edges > 0 the
// lists are finally cut to that length (:904-917). The batch search is the traversal kernel; the merges are radix
// sorts on the device.
namespace {

__global__ void refine_emit_kernel(const uint32_t *__restrict__ ids, const float *__restrict__ dists,
                                   const uint32_t *__restrict__ counts, uint32_t first_id, uint32_t count, uint32_t k,
                                   const uint8_t *__restrict__ valid, int reverse, uint32_t *__restrict__ out_src,
                                   uint64_t *__restrict__ out_key, unsigned long long *__restrict__ counter) {
  const uint64_t total = (uint64_t)count * k;
  for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < total; x += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t q = (uint32_t)(x / k), r = (uint32_t)(x % k);
    const uint32_t id = first_id + q;
    if (valid && valid[id] == 0) continue;          // objectRepository.isEmpty(id)
    const uint32_t c = counts[q];
    if (c == 0xffffffffu || r >= c) continue;
    const uint32_t t = ids[x];
    if (t == id || t == 0) continue;                // :872
    const uint32_t dord = ord_of_float(dists[x]);
    const unsigned long long p = atomicAdd(counter, reverse ? 2ull : 1ull);
    out_src[p] = id;
    out_key[p] = ((uint64_t)dord << 32) | t;
    if (reverse) {
      out_src[p + 1] = t;
      out_key[p + 1] = ((uint64_t)dord << 32) | id;
    }
  }
}

__global__ void truncate_lists_kernel(const uint64_t *__restrict__ row_ptr, uint64_t n, uint32_t keep, uint32_t *deg) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t d = row_ptr[i + 1] - row_ptr[i];
    deg[i] = (uint32_t)(d < keep ? d : keep);
  }
}

__global__ void gather_truncated_kernel(const uint64_t *__restrict__ old_ptr, const uint64_t *__restrict__ new_ptr, uint64_t n,
                                        const uint32_t *__restrict__ col, const float *__restrict__ dist,
                                        uint32_t *__restrict__ out_col, float *__restrict__ out_dist) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t id = warp; id <= n; id += warps) {
    const uint64_t ob = old_ptr[id], nb = new_ptr[id], d = new_ptr[id + 1] - nb;
    for (uint64_t i = lane; i < d; i += 32) {
      out_col[nb + i] = col[ob + i];
      out_dist[nb + i] = dist[ob + i];
    }
  }
}

}  // namespace

// A [n x k] neighbour table (the exhaustive kNN pass, ngtgpu_index_knn_graph) -> CSR. symmetric != 0 adds the reverse
// of every edge: the ANNG that insertANNGNode's out-edges + reverse edges (lib/NGT/Graph.h:611-626) converge to.
extern "C" int ngtgpu_graph_from_knn_table(uint64_t n, const uint32_t *d_ids, const float *d_dists, const uint32_t *d_counts,
                                           uint32_t k, const uint8_t *d_valid, int symmetric, uint64_t capacity,
                                           uint64_t *d_out_row_ptr, uint32_t *d_out_col, float *d_out_dist, uint64_t *out_nnz,
                                           void *stream_) {
  if (!d_ids || !d_dists || !d_counts || !d_out_row_ptr || !d_out_col || !d_out_dist || !out_nnz)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_from_knn_table: null buffer");
  if (n >= 0xfffffffeull) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_from_knn_table: too many nodes");
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)sms * 8;
  *out_nnz = 0;
  CUDA_TRY(cudaMemsetAsync(d_out_row_ptr, 0, (n + 2) * sizeof(uint64_t), stream));
  if (n == 0 || k == 0) return NGTGPU_OK;
  DeviceBuffers mem(stream);
  const uint64_t m_max = (uint64_t)n * k * (symmetric ? 2 : 1);
  uint32_t *src_a, *src_b;
  uint64_t *key_a, *key_b;
  unsigned long long *counter;
  CUDA_TRY(mem.alloc(&src_a, m_max));
  CUDA_TRY(mem.alloc(&src_b, m_max));
  CUDA_TRY(mem.alloc(&key_a, m_max));
  CUDA_TRY(mem.alloc(&key_b, m_max));
  CUDA_TRY(mem.alloc(&counter, 2));
  CUDA_TRY(cudaMemsetAsync(counter, 0, 16, stream));
  refine_emit_kernel<<<grid, 256, 0, stream>>>(d_ids, d_dists, d_counts, 1u, (uint32_t)n, k, d_valid, symmetric ? 1 : 0, src_a,
                                               key_a, counter);
  unsigned long long m = 0;
  CUDA_TRY(cudaMemcpyAsync(&m, counter, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (m == 0) return NGTGPU_OK;
  return csr_from_triples(mem, n, src_a, src_b, key_a, key_b, counter, m, capacity, d_out_row_ptr, d_out_col, d_out_dist,
                          out_nnz, grid, stream);
}

namespace {
__global__ void kept_degree_kernel(const uint64_t *__restrict__ row_ptr, uint64_t n, const uint8_t *__restrict__ keep,
                                   uint32_t *__restrict__ deg) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t id = warp; id <= n; id += warps) {
    uint32_t c = 0;
    for (uint64_t e = row_ptr[id] + lane; e < row_ptr[id + 1]; e += 32) c += keep[e] ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) deg[id] = c;
  }
}
__global__ void gather_kept_kernel(const uint64_t *__restrict__ row_ptr, const uint64_t *__restrict__ new_ptr, uint64_t n,
                                   const uint8_t *__restrict__ keep, const uint32_t *__restrict__ col,
                                   const float *__restrict__ dist, uint32_t *__restrict__ out_col, float *__restrict__ out_dist) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t id = warp; id <= n; id += warps) {
    uint64_t w = new_ptr[id];
    const uint64_t b = row_ptr[id], e = row_ptr[id + 1];
    for (uint64_t e0 = b; e0 < e; e0 += 32) {
      const uint64_t i = e0 + lane;
      const bool k = i < e && keep[i];
      const uint32_t m = __ballot_sync(0xffffffffu, k);
      if (k) {
        const uint64_t p = w + __popc(m & ((1u << lane) - 1u));
        out_col[p] = col[i];
        out_dist[p] = dist[i];
      }
      w += __popc(m);
    }
  }
}
}  // namespace

// The sub-graph of the edges with keep[e] != 0 (order inside the lists preserved). DEVICE buffers; out_col / out_dist
// need as many entries as edges are kept (the input's count always suffices).
extern "C" int ngtgpu_graph_select_edges(uint64_t n, const uint64_t *d_row_ptr, const uint32_t *d_col, const float *d_dist,
                                         const uint8_t *d_keep, uint64_t *d_out_row_ptr, uint32_t *d_out_col,
                                         float *d_out_dist, uint64_t *out_nnz, void *stream_) {
  if (!d_row_ptr || !d_out_row_ptr || !out_nnz) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_graph_select_edges: null buffer");
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)sms * 8;
  DeviceBuffers mem(stream);
  uint32_t *deg;
  CUDA_TRY(mem.alloc(&deg, n + 2));
  CUDA_TRY(cudaMemsetAsync(d_out_row_ptr, 0, 8, stream));
  kept_degree_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, n, d_keep, deg);
  size_t tb = 0;
  CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, tb, deg, d_out_row_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));
  uint8_t *tmp;
  CUDA_TRY(mem.alloc(&tmp, tb));
  CUDA_TRY(cub::DeviceScan::InclusiveScan(tmp, tb, deg, d_out_row_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));
  gather_kept_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, d_out_row_ptr, n, d_keep, d_col, d_dist, d_out_col, d_out_dist);
  CUDA_TRY(cudaMemcpyAsync(out_nnz, d_out_row_ptr + n + 1, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  CUDA_TRY(cudaGetLastError());
  return NGTGPU_OK;
}

int ngtgpu_search_prepared(ngtgpu_index *ix, const uint8_t *d_queries, uint32_t nq, const ngtgpu_search_params *params,
                           uint32_t n_seeds, uint32_t *d_ids, float *d_dists, uint32_t *d_counts, cudaStream_t stream);

extern "C" int ngtgpu_index_refine_anng(ngtgpu_index *ix, float epsilon, int32_t no_of_edges, int64_t edge_size,
                                        uint32_t searched_edges, uint64_t batch_size, uint32_t n_seeds, uint64_t capacity,
                                        uint64_t *d_row_ptr, uint32_t *d_col, float *d_dist, uint64_t *nnz_out) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!d_row_ptr || !d_col || !d_dist || !nnz_out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_refine_anng: null buffer");
  if (!ix->d_objects || ix->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_refine_anng: the index holds no objects");
  if (searched_edges == 0 || batch_size == 0) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_refine_anng: zero size");
  cudaStream_t stream = ix->stream;
  const uint64_t n = ix->n;
  const unsigned grid = (unsigned)ix->sm_count * 8;
  const uint32_t k = searched_edges;
  if (batch_size > n) batch_size = n;
  uint64_t nnz = 0;
  CUDA_TRY(cudaMemcpyAsync(&nnz, d_row_ptr + n + 1, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  DeviceBuffers mem(stream);
  uint32_t *r_ids, *r_counts, *t_col;
  float *r_dists, *t_dist;
  uint64_t *t_ptr;
  unsigned long long *counter;
  CUDA_TRY(mem.alloc(&r_ids, batch_size * k));
  CUDA_TRY(mem.alloc(&r_dists, batch_size * k));
  CUDA_TRY(mem.alloc(&r_counts, batch_size));
  CUDA_TRY(mem.alloc(&t_ptr, n + 2));
  CUDA_TRY(mem.alloc(&t_col, capacity));
  CUDA_TRY(mem.alloc(&t_dist, capacity));
  CUDA_TRY(mem.alloc(&counter, 2));
  ngtgpu_search_params sp;
  sp.size = k;
  sp.epsilon = epsilon;
  sp.radius = -1.0f;
  sp.edge_size = edge_size;
  for (uint64_t bid = 1; bid <= n; bid += batch_size) {
    const uint32_t count = (uint32_t)(n - bid + 1 < batch_size ? n - bid + 1 : batch_size);
    NGTGPU_TRY(ngtgpu_index_set_graph(ix, d_row_ptr, d_col, 1));
    NGTGPU_TRY(ngtgpu_search_prepared(ix, ix->d_objects + bid * ix->row_bytes, count, &sp, n_seeds, r_ids, r_dists, r_counts,
                                      stream));
    // old edges + new triples -> the refined graph
    const uint64_t m_max = nnz + 2ull * count * k;
    uint32_t *src_a, *src_b;
    uint64_t *key_a, *key_b;
    CUDA_TRY(mem.alloc(&src_a, m_max));
    CUDA_TRY(mem.alloc(&src_b, m_max));
    CUDA_TRY(mem.alloc(&key_a, m_max));
    CUDA_TRY(mem.alloc(&key_b, m_max));
    CUDA_TRY(cudaMemsetAsync(counter, 0, 16, stream));
    emit_edges_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, d_col, d_dist, n, 0xffffffffu, 0u, src_a, key_a, counter);
    refine_emit_kernel<<<grid, 256, 0, stream>>>(r_ids, r_dists, r_counts, (uint32_t)bid, count, k, ix->d_valid,
                                                 no_of_edges == 0 ? 1 : 0, src_a, key_a, counter);
    ix->launches += 2;
    unsigned long long m = 0;
    CUDA_TRY(cudaMemcpyAsync(&m, counter, 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (m) {
      uint64_t out_nnz = 0;
      NGTGPU_TRY(csr_from_triples(mem, n, src_a, src_b, key_a, key_b, counter, m, capacity, t_ptr, t_col, t_dist, &out_nnz, grid,
                                  stream));
      CUDA_TRY(cudaMemcpyAsync(d_row_ptr, t_ptr, (n + 2) * 8, cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(cudaMemcpyAsync(d_col, t_col, out_nnz * 4, cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(cudaMemcpyAsync(d_dist, t_dist, out_nnz * 4, cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      nnz = out_nnz;
    }
    mem.release(src_a);
    mem.release(src_b);
    mem.release(key_a);
    mem.release(key_b);
  }
  if (no_of_edges > 0) {   // prune to a kNN graph, GraphReconstructor.h:904-917
    uint32_t *deg;
    CUDA_TRY(mem.alloc(&deg, n + 2));
    truncate_lists_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, n, (uint32_t)no_of_edges, deg);
    size_t tb = 0;
    CUDA_TRY(cudaMemsetAsync(t_ptr, 0, 8, stream));
    CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, tb, deg, t_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));
    uint8_t *tmp;
    CUDA_TRY(mem.alloc(&tmp, tb));
    CUDA_TRY(cub::DeviceScan::InclusiveScan(tmp, tb, deg, t_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));
    gather_truncated_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, t_ptr, n, d_col, d_dist, t_col, t_dist);
    CUDA_TRY(cudaMemcpyAsync(&nnz, t_ptr + n + 1, 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaMemcpyAsync(d_row_ptr, t_ptr, (n + 2) * 8, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_col, t_col, nnz * 4, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_dist, t_dist, nnz * 4, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    ix->launches += 3;
  }
  NGTGPU_TRY(ngtgpu_index_set_graph(ix, d_row_ptr, d_col, 1));
  *nnz_out = nnz;
  return NGTGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ngtgpu_index_insert_batch -- one batch of the reference's ANNG construction loop (lib/NGT/Index.cpp:631-719,
// 721-792; Index.h:815-837; Graph.h:611-626, 845-886), which is also how objects are added to an existing index
// (ngt_insert_index + ngt_create_index):
//   search    every new object is searched for in the graph as it was BEFORE the batch (size = edgeSizeForCreation,
//             the creation epsilon; searched again with all edges when fewer results came back, Index.h:826-836)
//   a-13      insertMultipleSearchResults: object i of the batch also gets the distances to the objects j < i of the
//             same batch ("to imitate sequential insertion"), the list is sorted and cut to edgeSizeForCreation
//   a-15      insertANNGNode: the list becomes the node's edges and every listed node gets the reverse edge
// The searches are one launch of the traversal kernel, the in-batch distances one kernel (a warp per new object, the
// engine's exact distance), the insertion a radix-sort merge of the edge triples into the CSR.
namespace {

struct IntraArgs {
  const uint8_t *rows;        // the index's object table (row 0 = dummy)
  uint32_t row_bytes, chunks;
  uint32_t first_id, count, e;
  int dtype;
  const uint8_t *valid;
  const uint32_t *r_ids;      // [count x e] search results (ascending), nullable when the graph was empty
  const float *r_dists;
  const uint32_t *r_counts;
  uint32_t *out_ids;          // [count x e]
  float *out_dists;
  uint32_t *out_counts;
};

template <int ACC, int G>
__global__ void __launch_bounds__(256) intra_batch_kernel(const IntraArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t x = blockIdx.x * 8 + warp;
  if (x >= a.count) return;
  uint64_t *top = reinterpret_cast<uint64_t *>(smem) + (size_t)warp * a.e;
  const uint32_t id = a.first_id + x;
  uint32_t n = 0;
  const bool live = !(a.valid && a.valid[id] == 0);
  if (live && a.r_counts) {
    uint32_t c = a.r_counts[x];
    if (c == 0xffffffffu) c = 0;
    if (c > a.e) c = a.e;
    for (uint32_t i = lane; i < c; i += 32) top[i] = make_key(a.r_dists[(size_t)x * a.e + i], a.r_ids[(size_t)x * a.e + i]);
    n = c;
  }
  __syncwarp();
  if (live) {
    constexpr int R = 32 / G;
    const int gl = lane % G, grp = lane / G;
    const uint8_t *qptr = a.rows + (size_t)id * a.row_bytes;
    for (uint32_t j0 = 0; j0 < x; j0 += R) {   // the objects inserted before this one in the same batch
      const uint32_t j = j0 + grp;
      const uint32_t jid = a.first_id + j;
      const bool act = j < x && !(a.valid && a.valid[jid] == 0);
      const float d = group_distance_gmem<ACC, G>(qptr, act ? a.rows + (size_t)jid * a.row_bytes : qptr, a.chunks, gl, a.dtype);
      uint64_t key = KEY_NONE;
      if (act && gl == 0) key = make_key(d, jid);
      uint32_t mm = __ballot_sync(0xffffffffu, key != KEY_NONE);
      while (mm) {
        const int src = __ffs(mm) - 1;
        mm &= mm - 1;
        sorted_insert_u64(top, n, a.e, shfl_u64(key, src), lane);
      }
    }
  }
  for (uint32_t i = lane; i < a.e; i += 32) {
    const bool ok = i < n;
    a.out_ids[(size_t)x * a.e + i] = ok ? key_id(top[i]) : 0u;
    a.out_dists[(size_t)x * a.e + i] = ok ? key_dist(top[i]) : 0.f;
  }
  if (lane == 0) a.out_counts[x] = n;
}

template <int ACC>
cudaError_t launch_intra(int group, const IntraArgs &a, cudaStream_t stream) {
  const unsigned grid = (a.count + 7) / 8;
  const size_t smem = (size_t)8 * a.e * 8;
  switch (group) {
    case 1: intra_batch_kernel<ACC, 1><<<grid, 256, smem, stream>>>(a); break;
    case 2: intra_batch_kernel<ACC, 2><<<grid, 256, smem, stream>>>(a); break;
    case 4: intra_batch_kernel<ACC, 4><<<grid, 256, smem, stream>>>(a); break;
    case 8: intra_batch_kernel<ACC, 8><<<grid, 256, smem, stream>>>(a); break;
    case 16: intra_batch_kernel<ACC, 16><<<grid, 256, smem, stream>>>(a); break;
    case 32: intra_batch_kernel<ACC, 32><<<grid, 256, smem, stream>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// per query: take the second search (all edges) where the first one came back short (Index.h:826-836)
__global__ void pick_retry_kernel(uint32_t count, uint32_t e, uint32_t *ids, float *dists, uint32_t *counts,
                                  const uint32_t *ids2, const float *dists2, const uint32_t *counts2) {
  const uint64_t total = (uint64_t)count * e;
  for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < total; x += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t q = (uint32_t)(x / e);
    uint32_t c = counts[q];
    if (c == 0xffffffffu) c = 0;
    if (c < e) {
      ids[x] = ids2[x];
      dists[x] = dists2[x];
    }
  }
}
__global__ void pick_retry_counts_kernel(uint32_t count, uint32_t e, uint32_t *counts, const uint32_t *counts2) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  uint32_t c = counts[q];
  if (c == 0xffffffffu) c = 0;
  if (c < e) counts[q] = counts2[q];
}

}  // namespace

// ---- a batch's edge triples merged into the sorted lists of the graph ----------------------------------------
// The lists of the CSR are in (distance, target) order already and a batch touches a few thousand of them: instead of
// sorting every edge of the graph again (11 ms per batch of 200 at 1M objects), only the NEW triples are sorted by
// (source, distance, target); then one pass moves every list to its new place, a warp per list, and a list that gains
// edges is merged with its run of new triples by rank: an old entry moves up by the number of new keys below it, a new
// entry lands behind the old entries not above it (binary searches; hub lists of thousands of entries stay parallel).
// The full sort also drops an entry whose target equals the entry before it. That cannot happen when new nodes are
// inserted, but a caller may insert an id the graph already links: any such neighbourhood (and any list that does not
// come strictly ascending) raises a flag and the batch takes the full sort, which handles both.
namespace {

__global__ void list_degree_kernel(const uint64_t *__restrict__ row_ptr, uint64_t n, uint32_t *__restrict__ deg) {
  for (uint64_t id = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; id <= n; id += (uint64_t)gridDim.x * blockDim.x)
    deg[id] = (uint32_t)(row_ptr[id + 1] - row_ptr[id]);
}

// src sorted: every triple adds one to its source's degree, the first of a run records where the run starts
__global__ void run_degree_kernel(const uint32_t *__restrict__ src, uint64_t m, uint32_t *__restrict__ deg,
                                  uint32_t *__restrict__ run_start) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t s = src[i];
    atomicAdd(&deg[s], 1u);
    if (i == 0 || src[i - 1] != s) run_start[s] = (uint32_t)i;
  }
}

__device__ __forceinline__ uint64_t edge_key(const uint32_t *col, const float *dist, uint64_t e) {
  return ((uint64_t)ord_of_float(dist[e]) << 32) | col[e];
}

__global__ void merge_lists_kernel(const uint64_t *__restrict__ row_ptr, const uint64_t *__restrict__ new_ptr, uint64_t n,
                                   const uint32_t *__restrict__ col, const float *__restrict__ dist,
                                   const uint64_t *__restrict__ key, const uint32_t *__restrict__ run_start,
                                   uint32_t *__restrict__ out_col, float *__restrict__ out_dist, uint32_t *__restrict__ redo) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t id = warp; id <= n; id += warps) {
    const uint64_t ob = row_ptr[id], nb = new_ptr[id], d = row_ptr[id + 1] - ob;
    const uint64_t lb = (new_ptr[id + 1] - nb) - d;   // new triples of this list
    const uint64_t rs = lb ? run_start[id] : 0;
    for (uint64_t i = lane; i < d; i += 32) {
      const uint32_t t = col[ob + i];
      const float x = dist[ob + i];
      const uint64_t ka = ((uint64_t)ord_of_float(x) << 32) | t;
      if (i + 1 < d) {
        const uint32_t t1 = col[ob + i + 1];
        if (edge_key(col, dist, ob + i + 1) <= ka || t1 == t) *redo = 1u;
      }
      uint64_t lo = 0, hi = lb;   // new keys below this one
      while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (key[rs + mid] < ka) lo = mid + 1;
        else hi = mid;
      }
      out_col[nb + i + lo] = t;
      out_dist[nb + i + lo] = x;
    }
    for (uint64_t j = lane; j < lb; j += 32) {
      const uint64_t kb = key[rs + j];
      const uint32_t t = (uint32_t)kb;
      uint64_t lo = 0, hi = d;    // old keys not above this one
      while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (edge_key(col, dist, ob + mid) <= kb) lo = mid + 1;
        else hi = mid;
      }
      out_col[nb + j + lo] = t;
      out_dist[nb + j + lo] = float_of_ord((uint32_t)(kb >> 32));
      if ((lo > 0 && col[ob + lo - 1] == t) || (lo < d && col[ob + lo] == t) || (j > 0 && (uint32_t)key[rs + j - 1] == t)) *redo = 1u;
    }
  }
}

}  // namespace

extern "C" int ngtgpu_index_insert_batch(ngtgpu_index *ix, uint32_t first_id, uint32_t count, uint32_t edge_size_for_creation,
                                         float epsilon, int64_t edge_size, uint32_t n_seeds, uint32_t n_pivots,
                                         uint64_t pivot_seed, uint64_t capacity, uint64_t *d_row_ptr, uint32_t *d_col,
                                         float *d_dist, uint64_t *nnz_out) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!d_row_ptr || !d_col || !d_dist || !nnz_out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_insert_batch: null buffer");
  if (!ix->d_objects || ix->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_insert_batch: the index holds no objects");
  const uint32_t e = edge_size_for_creation;
  if (e == 0 || e > 1024) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_insert_batch: edge size for creation must be in [1, 1024]");
  if (first_id == 0 || count == 0 || (uint64_t)first_id + count - 1 > ix->n)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_insert_batch: ids out of range");
  cudaStream_t stream = ix->stream;
  const uint64_t n = ix->n;
  const unsigned grid = (unsigned)ix->sm_count * 8;
  uint64_t nnz = 0;
  CUDA_TRY(cudaMemcpyAsync(&nnz, d_row_ptr + n + 1, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  DeviceBuffers mem(stream);
  uint32_t *r_ids = nullptr, *r_counts = nullptr, *o_ids, *o_counts;
  float *r_dists = nullptr, *o_dists;
  CUDA_TRY(mem.alloc(&o_ids, (size_t)count * e));
  CUDA_TRY(mem.alloc(&o_dists, (size_t)count * e));
  CUDA_TRY(mem.alloc(&o_counts, count));
  const bool have_graph = nnz > 0 && first_id > 1;
  if (have_graph) {
    CUDA_TRY(mem.alloc(&r_ids, (size_t)count * e));
    CUDA_TRY(mem.alloc(&r_dists, (size_t)count * e));
    CUDA_TRY(mem.alloc(&r_counts, count));
    // (the previous batch left exactly this graph on the index: same caller buffers, same edge count)
    if (!(ix->graph_source == d_row_ptr && ix->nnz == nnz)) NGTGPU_TRY(ngtgpu_index_set_graph(ix, d_row_ptr, d_col, 1));
    if (n_pivots == 0) {   // SeedTypeFixedNodes (Index.h:1122-1127): ids 1..min(seedSize, nodes in the graph)
      std::vector<uint32_t> fixed(std::min<uint32_t>(n_seeds, first_id - 1));
      for (uint32_t i = 0; i < fixed.size(); i++) fixed[i] = i + 1;
      NGTGPU_TRY(ngtgpu_index_set_seed_table_ids(ix, fixed.data(), (uint32_t)fixed.size()));
    } else {
      NGTGPU_TRY(ngtgpu_index_build_seed_table_range(ix, n_pivots, pivot_seed, first_id - 1));
    }
    ngtgpu_search_params sp;
    sp.size = e;
    sp.epsilon = epsilon;
    sp.radius = -1.0f;
    sp.edge_size = edge_size;
    const uint8_t *q = ix->d_objects + (size_t)first_id * ix->row_bytes;
    NGTGPU_TRY(ngtgpu_search_prepared(ix, q, count, &sp, n_seeds, r_ids, r_dists, r_counts, stream));
    if (ngtgpu_effective_edge_size(ix, &sp) != 0x7fffffff) {
      // searched again without the edge cap where fewer than E results came back (and the graph has more nodes)
      std::vector<uint32_t> h_counts(count);
      CUDA_TRY(cudaMemcpyAsync(h_counts.data(), r_counts, (size_t)count * 4, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      bool shortfall = false;
      for (uint32_t c : h_counts)
        if (c != 0xffffffffu && c < e && c < first_id) shortfall = true;   // result.size() < repository.size()
      if (shortfall) {
        uint32_t *r2_ids, *r2_counts;
        float *r2_dists;
        CUDA_TRY(mem.alloc(&r2_ids, (size_t)count * e));
        CUDA_TRY(mem.alloc(&r2_dists, (size_t)count * e));
        CUDA_TRY(mem.alloc(&r2_counts, count));
        sp.edge_size = 0;
        NGTGPU_TRY(ngtgpu_search_prepared(ix, q, count, &sp, n_seeds, r2_ids, r2_dists, r2_counts, stream));
        pick_retry_kernel<<<grid, 256, 0, stream>>>(count, e, r_ids, r_dists, r_counts, r2_ids, r2_dists, r2_counts);
        pick_retry_counts_kernel<<<(count + 255) / 256, 256, 0, stream>>>(count, e, r_counts, r2_counts);
        ix->launches += 2;
        CUDA_TRY(cudaStreamSynchronize(stream));
        mem.release(r2_ids);
        mem.release(r2_dists);
        mem.release(r2_counts);
      }
    }
  }
  IntraArgs ia;
  ia.rows = ix->d_objects;
  ia.row_bytes = ix->row_bytes;
  ia.chunks = ix->chunks;
  ia.first_id = first_id;
  ia.count = count;
  ia.e = e;
  ia.dtype = ix->distance_type;
  ia.valid = ix->d_valid;
  ia.r_ids = r_ids;
  ia.r_dists = r_dists;
  ia.r_counts = r_counts;
  ia.out_ids = o_ids;
  ia.out_dists = o_dists;
  ia.out_counts = o_counts;
  cudaError_t ce;
  switch (ix->acc_kind) {
    case ACC_F_L2: ce = launch_intra<ACC_F_L2>((int)ix->group, ia, stream); break;
    case ACC_F_DOT: ce = launch_intra<ACC_F_DOT>((int)ix->group, ia, stream); break;
    case ACC_F_COS: ce = launch_intra<ACC_F_COS>((int)ix->group, ia, stream); break;
    case ACC_U8_L2: ce = launch_intra<ACC_U8_L2>((int)ix->group, ia, stream); break;
    default: ce = launch_intra<ACC_U8_HAM>((int)ix->group, ia, stream); break;
  }
  if (ce != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("in-batch distance kernel launch: ") + cudaGetErrorString(ce));
  ix->launches++;
  // (new node -> listed node) + (listed node -> new node) merged into the lists of the graph
  static const bool merge_off = getenv("NGTGPU_INSERT_MERGE") != nullptr && atoi(getenv("NGTGPU_INSERT_MERGE")) == 0;
  bool merged = false;
  if (nnz && !merge_off) {
    const uint64_t m_new_max = 2ull * count * e;
    uint32_t *ns_a, *ns_b, *deg, *unsorted;
    uint64_t *nk_a, *nk_b, *t_ptr;
    unsigned long long *cnt;
    CUDA_TRY(mem.alloc(&ns_a, m_new_max));
    CUDA_TRY(mem.alloc(&ns_b, m_new_max));
    CUDA_TRY(mem.alloc(&nk_a, m_new_max));
    CUDA_TRY(mem.alloc(&nk_b, m_new_max));
    CUDA_TRY(mem.alloc(&cnt, 2));
    CUDA_TRY(mem.alloc(&unsorted, 4));
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 16, stream));
    CUDA_TRY(cudaMemsetAsync(unsorted, 0, 16, stream));
    refine_emit_kernel<<<grid, 256, 0, stream>>>(o_ids, o_dists, o_counts, first_id, count, e, ix->d_valid, 1, ns_a, nk_a, cnt);
    ix->launches++;
    unsigned long long m_new = 0;
    CUDA_TRY(cudaMemcpyAsync(&m_new, cnt, 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (m_new == 0) {
      merged = true;   // nothing to add: the graph stays as it is
    } else {
      const uint64_t out_nnz = nnz + m_new;
      if (out_nnz > capacity)
        NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_insert_batch: graph capacity " + std::to_string(capacity) + " < " +
                                            std::to_string(out_nnz) + " edges");
      int hi_bits = 1;
      while (hi_bits < 32 && (n >> hi_bits) != 0) hi_bits++;
      size_t t1 = 0, t2 = 0, t3 = 0;
      uint32_t *run_start, *t_col;
      float *t_dist;
      CUDA_TRY(mem.alloc(&deg, n + 2));
      CUDA_TRY(mem.alloc(&run_start, n + 2));
      CUDA_TRY(mem.alloc(&t_ptr, n + 2));
      // (the caller's capacity, not this batch's edge count: the same size in every batch of a loop, so the block cache
      // serves it -- a size that grows by a few KB per batch would miss it every time)
      CUDA_TRY(mem.alloc(&t_col, (size_t)capacity));
      CUDA_TRY(mem.alloc(&t_dist, (size_t)capacity));
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, nk_a, nk_b, ns_a, ns_b, (int64_t)m_new_max, 0, 64, stream));
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t2, ns_b, ns_a, nk_b, nk_a, (int64_t)m_new_max, 0, hi_bits, stream));
      CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, t3, deg, t_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));
      uint8_t *tmp;
      CUDA_TRY(mem.alloc(&tmp, std::max(t1, std::max(t2, t3))));   // (sized for the largest batch: constant as well)
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, t1, nk_a, nk_b, ns_a, ns_b, (int64_t)m_new, 0, 64, stream));       // by (distance, target)
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, t2, ns_b, ns_a, nk_b, nk_a, (int64_t)m_new, 0, hi_bits, stream));  // then, stably, by source
      list_degree_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, n, deg);
      run_degree_kernel<<<grid, 256, 0, stream>>>(ns_a, m_new, deg, run_start);
      CUDA_TRY(cub::DeviceScan::InclusiveScan(tmp, t3, deg, t_ptr + 1, cub::Sum(), (int64_t)n + 1, stream));   // (t_ptr[0] is zero: a fresh block)
      merge_lists_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, t_ptr, n, d_col, d_dist, nk_a, run_start, t_col, t_dist, unsorted);
      ix->launches += 3;
      uint32_t bad = 0;
      CUDA_TRY(cudaMemcpyAsync(&bad, unsorted, 4, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      CUDA_TRY(cudaGetLastError());
      if (!bad) {
        CUDA_TRY(cudaMemcpyAsync(d_row_ptr, t_ptr, (n + 2) * 8, cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_col, t_col, out_nnz * 4, cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_dist, t_dist, out_nnz * 4, cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        nnz = out_nnz;
        merged = true;
      }
    }
    if (merged) g_batches_merged++;
  }
  if (!merged) {
  g_batches_sorted++;
  // the full sort: old edges + new triples (an empty graph, or lists that did not come in (distance, target) order)
  uint32_t *src_a, *src_b;
  uint64_t *key_a, *key_b;
  unsigned long long *counter;
  const uint64_t m_alloc = std::max<uint64_t>(capacity, nnz) + 2ull * count * e;   // the same in every batch of a loop (block cache)
  CUDA_TRY(mem.alloc(&src_a, m_alloc));
  CUDA_TRY(mem.alloc(&src_b, m_alloc));
  CUDA_TRY(mem.alloc(&key_a, m_alloc));
  CUDA_TRY(mem.alloc(&key_b, m_alloc));
  CUDA_TRY(mem.alloc(&counter, 2));
  CUDA_TRY(cudaMemsetAsync(counter, 0, 16, stream));
  if (nnz) emit_edges_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, d_col, d_dist, n, 0xffffffffu, 0u, src_a, key_a, counter);
  refine_emit_kernel<<<grid, 256, 0, stream>>>(o_ids, o_dists, o_counts, first_id, count, e, ix->d_valid, 1, src_a, key_a, counter);
  ix->launches += 2;
  unsigned long long m = 0;
  CUDA_TRY(cudaMemcpyAsync(&m, counter, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (m) {
    uint64_t *t_ptr;
    uint32_t *t_col;
    float *t_dist;
    CUDA_TRY(mem.alloc(&t_ptr, n + 2));
    CUDA_TRY(mem.alloc(&t_col, (size_t)m));
    CUDA_TRY(mem.alloc(&t_dist, (size_t)m));
    uint64_t out_nnz = 0;
    NGTGPU_TRY(csr_from_triples(mem, n, src_a, src_b, key_a, key_b, counter, m, m, t_ptr, t_col, t_dist, &out_nnz, grid, stream));
    if (out_nnz > capacity)
      NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_insert_batch: graph capacity " + std::to_string(capacity) + " < " +
                                          std::to_string(out_nnz) + " edges");
    CUDA_TRY(cudaMemcpyAsync(d_row_ptr, t_ptr, (n + 2) * 8, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_col, t_col, out_nnz * 4, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_dist, t_dist, out_nnz * 4, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    nnz = out_nnz;
  }
  }
  NGTGPU_TRY(ngtgpu_index_set_graph(ix, d_row_ptr, d_col, 1));
  ix->graph_source = d_row_ptr;
  *nnz_out = nnz;
  return NGTGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ngtgpu_index_build_onng -- the reference's ONNG recipe for the objects of one index, entirely on the device:
//   exact kNN table (the brute-force pass of Index.h:839-856, tensor cores where the shape allows)
//   -> kNN graph as CSR -> GraphReconstructor::reconstructGraph (outgoing / incoming, GraphReconstructor.h:425-561)
//   -> GraphReconstructor::adjustPathsEffectively (shortcut reduction, :197-386) -> the index's graph.
// What `ngt create -E knn` + `ngt reconstruct-graph -o outgoing -i incoming` produce from an exact neighbour table.
// graph_out (nullable): receives cudaMalloc'ed copies of the CSR with distances (free with ngtgpu_device_free).
extern "C" int ngtgpu_index_build_onng(ngtgpu_index *ix, uint32_t knn, uint32_t outgoing, uint32_t incoming,
                                       int shortcut_reduction, uint32_t min_edges, ngtgpu_graph_buffers *graph_out,
                                       double *seconds) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!ix->d_objects || ix->n < 2) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_build_onng: the index holds fewer than two objects");
  if (knn == 0) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_build_onng: knn is zero");
  cudaStream_t stream = ix->stream;
  const uint64_t n = ix->n;
  const uint32_t k = (uint32_t)std::min<uint64_t>(knn, n - 1);
  cudaEvent_t ev[4];
  for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
  struct EvGuard {
    cudaEvent_t *e;
    ~EvGuard() { for (int i = 0; i < 4; i++) cudaEventDestroy(e[i]); }
  } guard{ev};
  DeviceBuffers mem(stream);
  uint32_t *t_ids, *t_counts;
  float *t_dists;
  CUDA_TRY(mem.alloc(&t_ids, n * k));
  CUDA_TRY(mem.alloc(&t_dists, n * k));
  CUDA_TRY(mem.alloc(&t_counts, n));
  CUDA_TRY(cudaEventRecord(ev[0], stream));
  // query batches of three full waves of the tensor-core kernel's 256-query CTAs (no partly filled last wave)
  const uint64_t batch = (uint64_t)ix->sm_count * 256 * 3;
  for (uint64_t s = 0; s < n; s += batch) {
    const uint32_t m = (uint32_t)std::min<uint64_t>(batch, n - s);
    NGTGPU_TRY(ngtgpu_index_knn_graph(ix, k, (uint32_t)s + 1, m, t_ids + s * k, t_dists + s * k, t_counts + s, stream));
  }
  CUDA_TRY(cudaEventRecord(ev[1], stream));
  // kNN table -> CSR (lists ascending by (distance, id))
  uint64_t *a_rp, nnz_a = 0;
  uint32_t *a_col;
  float *a_dist;
  CUDA_TRY(mem.alloc(&a_rp, n + 2));
  CUDA_TRY(mem.alloc(&a_col, n * k));
  CUDA_TRY(mem.alloc(&a_dist, n * k));
  NGTGPU_TRY(ngtgpu_graph_from_knn_table(n, t_ids, t_dists, t_counts, k, ix->d_valid, 0, n * k, a_rp, a_col, a_dist, &nnz_a, stream));
  mem.release(t_ids);
  mem.release(t_dists);
  mem.release(t_counts);
  // reconstructGraph
  const uint64_t cap_b = std::max<uint64_t>(2 * nnz_a, 1);
  uint64_t *b_rp, nnz_b = 0;
  uint32_t *b_col;
  float *b_dist;
  CUDA_TRY(mem.alloc(&b_rp, n + 2));
  CUDA_TRY(mem.alloc(&b_col, cap_b));
  CUDA_TRY(mem.alloc(&b_dist, cap_b));
  NGTGPU_TRY(ngtgpu_graph_reconstruct(n, a_rp, a_col, a_dist, outgoing, incoming, cap_b, b_rp, b_col, b_dist, &nnz_b, stream));
  mem.release(a_rp);
  mem.release(a_col);
  mem.release(a_dist);
  CUDA_TRY(cudaEventRecord(ev[2], stream));
  uint64_t *g_rp = b_rp, nnz = nnz_b;
  uint32_t *g_col = b_col;
  float *g_dist = b_dist;
  if (shortcut_reduction && nnz_b) {
    uint8_t *keep;
    CUDA_TRY(mem.alloc(&keep, nnz_b));
    NGTGPU_TRY(ngtgpu_graph_adjust_paths(n, b_rp, b_col, b_dist, min_edges, keep, nullptr, stream));
    uint64_t *c_rp;
    uint32_t *c_col;
    float *c_dist;
    CUDA_TRY(mem.alloc(&c_rp, n + 2));
    CUDA_TRY(mem.alloc(&c_col, nnz_b));
    CUDA_TRY(mem.alloc(&c_dist, nnz_b));
    NGTGPU_TRY(ngtgpu_graph_select_edges(n, b_rp, b_col, b_dist, keep, c_rp, c_col, c_dist, &nnz, stream));
    mem.release(keep);
    mem.release(b_rp);
    mem.release(b_col);
    mem.release(b_dist);
    g_rp = c_rp, g_col = c_col, g_dist = c_dist;
  }
  CUDA_TRY(cudaEventRecord(ev[3], stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  NGTGPU_TRY(ngtgpu_index_set_graph(ix, g_rp, g_col, 1));
  if (seconds) {
    float ms = 0.f;
    for (int i = 0; i < 3; i++) {
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      seconds[i] = ms * 1e-3;   // kNN pass, reconstructGraph, path adjustment
    }
  }
  if (graph_out) {
    graph_out->n = n;
    graph_out->nnz = nnz;
    graph_out->row_ptr = nullptr, graph_out->col = nullptr, graph_out->dist = nullptr;
    CUDA_TRY(cudaMalloc(&graph_out->row_ptr, (n + 2) * 8));
    CUDA_TRY(cudaMalloc(&graph_out->col, std::max<uint64_t>(nnz, 1) * 4));
    CUDA_TRY(cudaMalloc(&graph_out->dist, std::max<uint64_t>(nnz, 1) * 4));
    CUDA_TRY(cudaMemcpy(graph_out->row_ptr, g_rp, (n + 2) * 8, cudaMemcpyDeviceToDevice));
    if (nnz) {
      CUDA_TRY(cudaMemcpy(graph_out->col, g_col, nnz * 4, cudaMemcpyDeviceToDevice));
      CUDA_TRY(cudaMemcpy(graph_out->dist, g_dist, nnz * 4, cudaMemcpyDeviceToDevice));
    }
    CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));   // (device-to-device copies do not wait on the host)
  }
  return NGTGPU_OK;
}

extern "C" int ngtgpu_construction_counters(uint64_t out[4]) {
  if (!out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_construction_counters: null buffer");
  out[0] = g_batches_merged, out[1] = g_batches_sorted, out[2] = g_blocks_malloc, out[3] = g_blocks_cached;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_device_copy(void *dst, const void *src, uint64_t bytes) {
  if (bytes) {
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));   // (device-to-device copies do not wait on the host)
  }
  return NGTGPU_OK;
}

extern "C" int ngtgpu_device_free(void *device_pointer) {
  if (device_pointer) CUDA_TRY(cudaFree(device_pointer));
  return NGTGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Distances among a handful of stored objects, the engine's exact distance: what NeighborhoodGraph::removeEdgesReliably
// (lib/NGT/Graph.cpp:641-864) asks the comparator for when it re-links the neighbours of a removed node.
namespace {
template <int ACC, int G>
__global__ void __launch_bounds__(256) pairwise_kernel(const uint8_t *__restrict__ rows, uint32_t row_bytes, uint32_t chunks, int dtype,
                                                       const uint32_t *__restrict__ ids, uint32_t m, float *__restrict__ out) {
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t total = (uint64_t)m * m;
  const uint64_t pair = warp * R + grp;
  const bool act = pair < total;
  const uint32_t i = act ? (uint32_t)(pair / m) : 0u, j = act ? (uint32_t)(pair % m) : 0u;
  const float d = group_distance_gmem<ACC, G>(rows + (size_t)ids[i] * row_bytes, rows + (size_t)ids[j] * row_bytes, chunks, gl, dtype);
  if (act && gl == 0) out[pair] = d;
}
template <int ACC>
cudaError_t launch_pairwise(int group, const uint8_t *rows, uint32_t row_bytes, uint32_t chunks, int dtype, const uint32_t *ids, uint32_t m,
                            float *out, cudaStream_t stream) {
  const uint64_t pairs = (uint64_t)m * m;
  const uint64_t warps = (pairs + (32 / group) - 1) / (32 / group);
  const unsigned grid = (unsigned)((warps + 7) / 8);
  switch (group) {
    case 1: pairwise_kernel<ACC, 1><<<grid, 256, 0, stream>>>(rows, row_bytes, chunks, dtype, ids, m, out); break;
    case 2: pairwise_kernel<ACC, 2><<<grid, 256, 0, stream>>>(rows, row_bytes, chunks, dtype, ids, m, out); break;
    case 4: pairwise_kernel<ACC, 4><<<grid, 256, 0, stream>>>(rows, row_bytes, chunks, dtype, ids, m, out); break;
    case 8: pairwise_kernel<ACC, 8><<<grid, 256, 0, stream>>>(rows, row_bytes, chunks, dtype, ids, m, out); break;
    case 16: pairwise_kernel<ACC, 16><<<grid, 256, 0, stream>>>(rows, row_bytes, chunks, dtype, ids, m, out); break;
    case 32: pairwise_kernel<ACC, 32><<<grid, 256, 0, stream>>>(rows, row_bytes, chunks, dtype, ids, m, out); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
}  // namespace

extern "C" int ngtgpu_index_pairwise_distances(ngtgpu_index *ix, const uint32_t *ids, uint32_t m, float *out) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (m == 0) return NGTGPU_OK;
  if (!ids || !out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_pairwise_distances: null buffer");
  if (!ix->d_objects) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_pairwise_distances: the index holds no objects");
  if (m > 4096) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_pairwise_distances: at most 4096 objects");
  for (uint32_t i = 0; i < m; i++)
    if (ids[i] == 0 || ids[i] > ix->n) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_pairwise_distances: id out of range");
  cudaStream_t stream = ix->stream;
  DeviceBuffers mem(stream);
  uint32_t *d_ids;
  float *d_out;
  CUDA_TRY(mem.alloc(&d_ids, m));
  CUDA_TRY(mem.alloc(&d_out, (size_t)m * m));
  CUDA_TRY(cudaMemcpyAsync(d_ids, ids, (size_t)m * 4, cudaMemcpyHostToDevice, stream));
  cudaError_t ce;
  switch (ix->acc_kind) {
    case ACC_F_L2: ce = launch_pairwise<ACC_F_L2>((int)ix->group, ix->d_objects, ix->row_bytes, ix->chunks, ix->distance_type, d_ids, m, d_out, stream); break;
    case ACC_F_DOT: ce = launch_pairwise<ACC_F_DOT>((int)ix->group, ix->d_objects, ix->row_bytes, ix->chunks, ix->distance_type, d_ids, m, d_out, stream); break;
    case ACC_F_COS: ce = launch_pairwise<ACC_F_COS>((int)ix->group, ix->d_objects, ix->row_bytes, ix->chunks, ix->distance_type, d_ids, m, d_out, stream); break;
    case ACC_U8_L2: ce = launch_pairwise<ACC_U8_L2>((int)ix->group, ix->d_objects, ix->row_bytes, ix->chunks, ix->distance_type, d_ids, m, d_out, stream); break;
    default: ce = launch_pairwise<ACC_U8_HAM>((int)ix->group, ix->d_objects, ix->row_bytes, ix->chunks, ix->distance_type, d_ids, m, d_out, stream); break;
  }
  if (ce != cudaSuccess) NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("pairwise distance kernel launch: ") + cudaGetErrorString(ce));
  ix->launches++;
  CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)m * m * 4, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return NGTGPU_OK;
}

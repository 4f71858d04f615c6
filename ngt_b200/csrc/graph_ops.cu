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

struct DeviceBuffers {   // frees what it allocated, whichever way the function returns
  std::vector<void *> ptrs;
  ~DeviceBuffers() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <typename T>
  cudaError_t alloc(T **out, size_t count) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, count ? count * sizeof(T) : sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = static_cast<T *>(p);
    return e;
  }
  void release(void *p) {
    for (size_t i = 0; i < ptrs.size(); i++)
      if (ptrs[i] == p) {
        cudaFree(p);
        ptrs.erase(ptrs.begin() + i);
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
  DeviceBuffers mem;

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

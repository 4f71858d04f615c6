// index.cu -- life cycle of the device-resident index, object/graph upload, query preparation.
//
// Replaces, for the HBM copy of an index, what GraphIndex's constructor and loadIndex do on the host
// in the reference (lib/NGT/Index.cpp:587-606, Index.h:665-695): objects become one row-major padded
// array [(n+1) x row_bytes] (row 0 = the dummy slot, ObjectSpace.h:357-400 padding is zeros), the
// adjacency lists one CSR. There is no CPU fallback anywhere in this library.
#include <cstring>
#include <mutex>

#include "ngtgpu_internal.cuh"

static thread_local std::string g_last_error;
void ngtgpu_set_error(const std::string &msg) { g_last_error = msg; }

extern "C" const char *ngtgpu_last_error(void) { return g_last_error.c_str(); }

extern "C" int ngtgpu_device_count(int *count) {
  if (!count) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_device_count: null argument");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess || c == 0) {
    *count = 0;
    cudaGetLastError();
    NGTGPU_FAIL(NGTGPU_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                          " (this engine has no CPU fallback)");
  }
  *count = c;
  return NGTGPU_OK;
}

int ngtgpu_check_device(ngtgpu_index *ix) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  CUDA_TRY(cudaSetDevice(ix->device));
  return NGTGPU_OK;
}

// the lane the calling thread holds (ngtgpu_lane_guard), and of which index; none: lane 0 of whatever index is asked
static thread_local ngtgpu_lane *tls_lane = nullptr;
static thread_local ngtgpu_index *tls_lane_owner = nullptr;

ngtgpu_lane_guard::ngtgpu_lane_guard(ngtgpu_index *index) : ix(index), lane(0), prev(tls_lane) {
  int got = -1;
  for (int l = 0; l < NGTGPU_LANES && got < 0; l++)
    if (ix->lane_mutex[l].try_lock()) got = l;
  if (got < 0) {
    ix->lane_mutex[0].lock();
    got = 0;
  }
  lane = got;
  if (lane > 0) {
    ngtgpu_lane &L = ix->lanes[lane - 1];
    if (!L.stream) {
      cudaSetDevice(ix->device);
      if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) {
        ngtgpu_set_error("cudaStreamCreate failed for a search lane");
        status = NGTGPU_ERR_CUDA;
      }
    }
    tls_lane = &L;
    tls_lane_owner = ix;
  } else {
    tls_lane = nullptr;
    tls_lane_owner = nullptr;
  }
}
ngtgpu_lane_guard::~ngtgpu_lane_guard() {
  tls_lane = prev;
  if (!prev) tls_lane_owner = nullptr;
  ix->lane_mutex[lane].unlock();
}
cudaStream_t ngtgpu_lane_guard::stream() const { return lane > 0 ? ix->lanes[lane - 1].stream : ix->stream; }

int ngtgpu_scratch(ngtgpu_index *ix, int slot, size_t bytes, void **out) {
  const bool mine = tls_lane && tls_lane_owner == ix;
  void **ptrs = mine ? tls_lane->d_scratch : ix->d_scratch;
  size_t *sizes = mine ? tls_lane->scratch_bytes : ix->scratch_bytes;
  cudaStream_t stream = mine ? tls_lane->stream : ix->stream;
  if (bytes == 0) bytes = 16;
  if (sizes[slot] < bytes) {
    if (ptrs[slot]) {
      // buffers may still be in use by work queued on the lane's stream
      CUDA_TRY(cudaStreamSynchronize(stream));
      CUDA_TRY(cudaFree(ptrs[slot]));
      ptrs[slot] = nullptr;
      sizes[slot] = 0;
    }
    size_t want = bytes + bytes / 4;
    want = (want + 255) & ~(size_t)255;
    CUDA_TRY(cudaMalloc(&ptrs[slot], want));
    sizes[slot] = want;
  }
  *out = ptrs[slot];
  return NGTGPU_OK;
}

static int acc_kind_of(int object_type, int distance_type) {
  if (object_type == NGTGPU_OBJECT_UINT8) {
    if (distance_type == NGTGPU_DISTANCE_L2) return ACC_U8_L2;
    if (distance_type == NGTGPU_DISTANCE_HAMMING) return ACC_U8_HAM;
    return -1;
  }
  if (object_type == NGTGPU_OBJECT_FLOAT) {
    switch (distance_type) {
      case NGTGPU_DISTANCE_L2: return ACC_F_L2;
      case NGTGPU_DISTANCE_ANGLE:
      case NGTGPU_DISTANCE_COSINE: return ACC_F_COS;
      case NGTGPU_DISTANCE_NORMALIZED_ANGLE:
      case NGTGPU_DISTANCE_NORMALIZED_COSINE:
      case NGTGPU_DISTANCE_NORMALIZED_L2: return ACC_F_DOT;
      default: return -1;
    }
  }
  return -1;
}

extern "C" int ngtgpu_index_create(ngtgpu_index **out, int device, int object_type, int distance_type,
                                   uint32_t dimension) {
  if (!out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_create: null output handle");
  *out = nullptr;
  int count = 0;
  NGTGPU_TRY(ngtgpu_device_count(&count));
  if (device < 0 || device >= count) NGTGPU_FAIL(NGTGPU_ERR_NO_DEVICE, "ngtgpu_index_create: no such device");
  int acc = acc_kind_of(object_type, distance_type);
  if (acc < 0)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID,
                "ngtgpu_index_create: unsupported object/distance type pair (supported: uint8 x {L2, Hamming}; "
                "float x {L2, Angle, Cosine, NormalizedAngle, NormalizedCosine, NormalizedL2})");
  if (dimension == 0) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_create: dimension is zero");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    NGTGPU_FAIL(NGTGPU_ERR_NO_DEVICE, std::string("device is ") + prop.name +
                                          ", this library is built for sm_100a only (no fallback path)");
  ngtgpu_index *ix = new ngtgpu_index();
  ix->device = device;
  ix->object_type = object_type;
  ix->distance_type = distance_type;
  ix->acc_kind = acc;
  ix->dim = dimension;
  ix->padded_dim = ((dimension - 1) / 16 + 1) * 16;  // ObjectSpace.h:249
  ix->elem_size = object_type == NGTGPU_OBJECT_UINT8 ? 1 : 4;
  ix->row_bytes = ix->padded_dim * ix->elem_size;
  ix->chunks = ix->row_bytes / 16;
  uint32_t g = 1;
  while (g < ix->chunks && g < 32) g <<= 1;
  ix->group = g;
  ix->normalizes = distance_type == NGTGPU_DISTANCE_NORMALIZED_ANGLE ||
                   distance_type == NGTGPU_DISTANCE_NORMALIZED_COSINE ||
                   distance_type == NGTGPU_DISTANCE_NORMALIZED_L2;
  ix->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ix;
    NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  *out = ix;
  return NGTGPU_OK;
}

static void free_graph(ngtgpu_index *ix) {
  if (ix->d_row_ptr) cudaFree(ix->d_row_ptr);
  if (ix->d_col) cudaFree(ix->d_col);
  if (ix->d_head) cudaFree(ix->d_head);
  ix->d_head = nullptr;
  ix->d_row_ptr = nullptr;
  ix->d_col = nullptr;
  ix->nnz = 0;
  ix->row_ptr_cap = ix->col_cap = ix->head_cap = 0;
  ix->graph_source = nullptr;
}
static void free_tc(ngtgpu_index *ix) {
  if (ix->d_tc_tiles) cudaFree(ix->d_tc_tiles);
  if (ix->d_tc_norms) cudaFree(ix->d_tc_norms);
  ix->d_tc_tiles = nullptr;
  ix->d_tc_norms = nullptr;
  ix->tc_rows_valid = false;
}
static void free_pivots(ngtgpu_index *ix) {
  if (ix->d_pivot_rows) cudaFree(ix->d_pivot_rows);
  if (ix->d_pivot_ids) cudaFree(ix->d_pivot_ids);
  ix->d_pivot_rows = nullptr;
  ix->d_pivot_ids = nullptr;
  ix->n_pivots = 0;
  ix->pivot_cap = 0;
}

extern "C" int ngtgpu_index_destroy(ngtgpu_index *ix) {
  if (!ix) return NGTGPU_OK;
  cudaSetDevice(ix->device);
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  if (ix->d_objects) cudaFree(ix->d_objects);
  if (ix->d_valid) cudaFree(ix->d_valid);
  free_graph(ix);
  free_pivots(ix);
  free_tc(ix);
  for (int i = 0; i < SCR_COUNT; i++)
    if (ix->d_scratch[i]) cudaFree(ix->d_scratch[i]);
  for (auto &L : ix->lanes) {
    if (L.stream) cudaStreamSynchronize(L.stream);
    for (int i = 0; i < SCR_COUNT; i++)
      if (L.d_scratch[i]) cudaFree(L.d_scratch[i]);
    if (L.stream) cudaStreamDestroy(L.stream);
  }
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
  return NGTGPU_OK;
}

extern "C" uint64_t ngtgpu_index_size(const ngtgpu_index *ix) { return ix ? ix->n : 0; }
extern "C" uint32_t ngtgpu_index_padded_dimension(const ngtgpu_index *ix) { return ix ? ix->padded_dim : 0; }
extern "C" uint64_t ngtgpu_index_launch_count(const ngtgpu_index *ix) { return ix ? ix->launches.load() : 0; }
extern "C" uint64_t ngtgpu_index_last_overflows(const ngtgpu_index *ix) { return ix ? ix->last_overflows.load() : 0; }

// ---- row preparation ---------------------------------------------------------------------------------
// One warp per row: cast `src` (float or uint8, `dim` wide) to the object type, zero pad to the padded
// dimension, optionally divide by the L2 norm (ObjectSpace.h:251-266: float accumulator, element / norm).
// flags[0] is set when a zero vector meets normalisation (the reference throws there).
template <typename SRC, typename DST>
__global__ void prepare_rows_kernel(const SRC *__restrict__ src, uint64_t n_rows, uint32_t dim, uint32_t padded,
                                    DST *__restrict__ dst, int normalize, int *__restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n_rows; r += nwarps) {
    const SRC *s = src + r * dim;
    DST *d = dst + r * padded;
    float inv = 1.0f;
    bool zero = false;
    if (normalize) {
      float sum = 0.f;
      for (uint32_t i = lane; i < dim; i += 32) {
        float v = (float)s[i];
        sum = fmaf(v, v, sum);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (sum == 0.0f) {
        zero = true;
        if (lane == 0) atomicExch(flags, 1);
      } else {
        inv = __fsqrt_rn(sum);
      }
    }
    for (uint32_t i = lane; i < padded; i += 32) {
      DST v = (DST)0;
      if (i < dim) {
        if (normalize && !zero) {
          v = (DST)__fdiv_rn((float)s[i], inv);
        } else {
          v = (DST)s[i];
        }
      }
      d[i] = v;
    }
  }
}

template <typename SRC, typename DST>
static int launch_prepare(ngtgpu_index *ix, const void *src, uint64_t n_rows, void *dst, int normalize, int *d_flags,
                          cudaStream_t stream) {
  if (n_rows == 0) return NGTGPU_OK;
  uint64_t blocks = (n_rows + 7) / 8;
  if (blocks > (uint64_t)ix->sm_count * 32) blocks = (uint64_t)ix->sm_count * 32;
  prepare_rows_kernel<SRC, DST><<<(unsigned)blocks, 256, 0, stream>>>((const SRC *)src, n_rows, ix->dim,
                                                                       ix->padded_dim, (DST *)dst, normalize, d_flags);
  ix->launches++;
  CUDA_TRY(cudaGetLastError());
  return NGTGPU_OK;
}

static int prepare_dispatch(ngtgpu_index *ix, const void *d_src, int src_type, uint64_t n_rows, void *d_dst,
                            int normalize, int *d_flags, cudaStream_t stream) {
  if (ix->object_type == NGTGPU_OBJECT_FLOAT) {
    if (src_type == NGTGPU_OBJECT_FLOAT)
      return launch_prepare<float, float>(ix, d_src, n_rows, d_dst, normalize, d_flags, stream);
    return launch_prepare<uint8_t, float>(ix, d_src, n_rows, d_dst, normalize, d_flags, stream);
  }
  // uint8 objects: float sources are narrowed the way the reference's allocateObject does
  // (ObjectRepository.h:222-258 static_cast to the object type); no normalisation for integer spaces.
  if (src_type == NGTGPU_OBJECT_FLOAT)
    return launch_prepare<float, uint8_t>(ix, d_src, n_rows, d_dst, 0, d_flags, stream);
  return launch_prepare<uint8_t, uint8_t>(ix, d_src, n_rows, d_dst, 0, d_flags, stream);
}

static int check_zero_flag(int *d_flags, cudaStream_t stream, const char *what) {
  int h = 0;
  CUDA_TRY(cudaMemcpyAsync(&h, d_flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (h) NGTGPU_FAIL(NGTGPU_ERR_ZERO_VECTOR, std::string(what) + ": normalization of a zero vector (ObjectSpace.h:256-260)");
  return NGTGPU_OK;
}

int ngtgpu_prepare_queries(ngtgpu_index *ix, const void *queries, int query_type, uint32_t nq, bool on_device,
                           uint8_t *d_out, cudaStream_t stream) {
  if (query_type != NGTGPU_OBJECT_FLOAT && query_type != NGTGPU_OBJECT_UINT8)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "query_type must be NGTGPU_OBJECT_FLOAT or NGTGPU_OBJECT_UINT8");
  const void *d_src = queries;
  size_t src_bytes = (size_t)nq * ix->dim * (query_type == NGTGPU_OBJECT_FLOAT ? 4 : 1);
  if (!on_device) {
    void *raw = nullptr;
    NGTGPU_TRY(ngtgpu_scratch(ix, SCR_RAW_QUERIES, src_bytes + 16, &raw));
    CUDA_TRY(cudaMemcpyAsync(raw, queries, src_bytes, cudaMemcpyHostToDevice, stream));
    d_src = raw;
  }
  int *d_flags = nullptr;
  int normalize = ix->normalizes ? 1 : 0;
  if (normalize) {
    // the flag lives at the tail of the raw-query scratch when we own it, else in its own slot
    void *f = nullptr;
    NGTGPU_TRY(ngtgpu_scratch(ix, SCR_SEED_DISTS, 256, &f));
    d_flags = (int *)f;
    CUDA_TRY(cudaMemsetAsync(d_flags, 0, sizeof(int), stream));
  }
  NGTGPU_TRY(prepare_dispatch(ix, d_src, query_type, nq, d_out, normalize, d_flags, stream));
  if (normalize) NGTGPU_TRY(check_zero_flag(d_flags, stream, "query"));
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_objects(ngtgpu_index *ix, const void *objects, uint64_t n, int normalize,
                                        int on_device) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!objects && n) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_objects: null objects");
  if (n >= 0xfffffffeull) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_objects: ObjectID is 32 bits (Common.h:46)");
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  if (ix->d_objects) CUDA_TRY(cudaFree(ix->d_objects));
  if (ix->d_valid) CUDA_TRY(cudaFree(ix->d_valid));
  ix->d_objects = nullptr;
  ix->d_valid = nullptr;
  free_graph(ix);
  free_pivots(ix);
  free_tc(ix);
  ix->n = 0;
  size_t bytes = (size_t)(n + 1) * ix->row_bytes;
  CUDA_TRY(cudaMalloc(&ix->d_objects, bytes));
  CUDA_TRY(cudaMemsetAsync(ix->d_objects, 0, ix->row_bytes, ix->stream));
  const void *d_src = objects;
  void *staging = nullptr;
  size_t src_row = (size_t)ix->dim * ix->elem_size;
  const size_t CHUNK_ROWS = (size_t)1 << 20;
  int *d_flags = nullptr;
  CUDA_TRY(cudaMalloc(&d_flags, sizeof(int)));
  CUDA_TRY(cudaMemsetAsync(d_flags, 0, sizeof(int), ix->stream));
  int norm = (normalize && ix->object_type == NGTGPU_OBJECT_FLOAT) ? 1 : 0;
  int rc = NGTGPU_OK;
  if (!on_device) {
    cudaError_t e = cudaMalloc(&staging, CHUNK_ROWS * src_row);
    if (e != cudaSuccess) {
      cudaFree(d_flags);
      NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("cudaMalloc staging: ") + cudaGetErrorString(e));
    }
  }
  for (uint64_t s = 0; s < n && rc == NGTGPU_OK; s += CHUNK_ROWS) {
    uint64_t m = n - s < CHUNK_ROWS ? n - s : CHUNK_ROWS;
    const uint8_t *src = (const uint8_t *)objects + s * src_row;
    if (!on_device) {
      cudaError_t e = cudaMemcpyAsync(staging, src, m * src_row, cudaMemcpyHostToDevice, ix->stream);
      if (e != cudaSuccess) {
        ngtgpu_set_error(std::string("cudaMemcpyAsync objects: ") + cudaGetErrorString(e));
        rc = NGTGPU_ERR_CUDA;
        break;
      }
      d_src = staging;
    } else {
      d_src = src;
    }
    rc = prepare_dispatch(ix, d_src, ix->object_type, m, ix->d_objects + (s + 1) * ix->row_bytes, norm, d_flags,
                          ix->stream);
    if (!on_device && rc == NGTGPU_OK) {
      // the staging buffer is reused by the next chunk
      cudaError_t e = cudaStreamSynchronize(ix->stream);
      if (e != cudaSuccess) {
        ngtgpu_set_error(std::string("set_objects: ") + cudaGetErrorString(e));
        rc = NGTGPU_ERR_CUDA;
      }
    }
  }
  if (rc == NGTGPU_OK) rc = check_zero_flag(d_flags, ix->stream, "ngtgpu_index_set_objects");
  if (staging) cudaFree(staging);
  cudaFree(d_flags);
  if (rc != NGTGPU_OK) {
    cudaFree(ix->d_objects);
    ix->d_objects = nullptr;
    return rc;
  }
  ix->n = n;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_removed(ngtgpu_index *ix, const uint32_t *ids, uint64_t count) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!ix->d_objects) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_set_removed: objects are not set");
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  if (ix->d_valid) CUDA_TRY(cudaFree(ix->d_valid));
  ix->d_valid = nullptr;
  free_tc(ix);   // the packed norms carry the empty-slot marks
  if (count == 0) return NGTGPU_OK;
  std::vector<uint8_t> valid(ix->n + 1, 1);
  valid[0] = 0;
  for (uint64_t i = 0; i < count; i++) {
    if (ids[i] == 0 || ids[i] > ix->n) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_removed: id out of range");
    valid[ids[i]] = 0;
  }
  CUDA_TRY(cudaMalloc(&ix->d_valid, ix->n + 1));
  CUDA_TRY(cudaMemcpy(ix->d_valid, valid.data(), ix->n + 1, cudaMemcpyHostToDevice));
  return NGTGPU_OK;
}

// fixed-stride copy of the first NGTGPU_HEAD_WIDTH (128) edges of every node (zero padded): the traversal kernel reads one node's
// edges with a single coalesced access instead of row_ptr -> col (two dependent ones)
__global__ void build_head_kernel(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, uint64_t n,
                                  uint32_t *__restrict__ head) {
  const uint64_t total = (n + 1) * NGTGPU_HEAD_WIDTH;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t id = i / NGTGPU_HEAD_WIDTH;
    const uint32_t e = (uint32_t)(i % NGTGPU_HEAD_WIDTH);
    const uint64_t b = row_ptr[id], deg = row_ptr[id + 1] - b;
    head[i] = e < deg ? col[b + e] : 0u;
  }
}

extern "C" int ngtgpu_index_set_graph(ngtgpu_index *ix, const uint64_t *row_ptr, const uint32_t *col, int on_device) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!ix->d_objects) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_set_graph: objects are not set");
  if (!row_ptr) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_graph: null row_ptr");
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  ix->graph_source = nullptr;
  uint64_t nnz = 0;
  cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (on_device) {
    CUDA_TRY(cudaMemcpy(&nnz, row_ptr + ix->n + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  } else {
    nnz = row_ptr[ix->n + 1];
    for (uint64_t i = 0; i <= ix->n; i++)
      if (row_ptr[i] > row_ptr[i + 1]) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_graph: row_ptr is not monotone");
    for (uint64_t j = 0; j < nnz; j++)
      if (col[j] == 0 || col[j] > ix->n) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_graph: edge to an id out of range");
  }
  if (nnz && !col) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_graph: null col");
  // the buffers of the previous graph are reused when they are large enough (the edge list grows by half when it is not)
  if (ix->row_ptr_cap < ix->n + 2) {
    if (ix->d_row_ptr) cudaFree(ix->d_row_ptr);
    ix->d_row_ptr = nullptr, ix->row_ptr_cap = 0;
    CUDA_TRY(cudaMalloc(&ix->d_row_ptr, (ix->n + 2) * sizeof(uint64_t)));
    ix->row_ptr_cap = ix->n + 2;
  }
  if (ix->col_cap < (nnz ? nnz : 1)) {
    if (ix->d_col) cudaFree(ix->d_col);
    ix->d_col = nullptr;
    const uint64_t want = ix->col_cap ? nnz + nnz / 2 : (nnz ? nnz : 1);
    ix->col_cap = 0;
    CUDA_TRY(cudaMalloc(&ix->d_col, want * sizeof(uint32_t)));
    ix->col_cap = want;
  }
  // On the index's stream, in order with the head-table kernel below: a synchronous cudaMemcpy between two device
  // buffers does NOT wait on the host and runs on the legacy default stream, which the (non-blocking) stream of the index
  // is not ordered with -- the kernel could read the previous graph's edges (seen as refineANNG runs on 100k objects
  // that differed from one call to the next).
  CUDA_TRY(cudaMemcpyAsync(ix->d_row_ptr, row_ptr, (ix->n + 2) * sizeof(uint64_t), kind, ix->stream));
  if (nnz) CUDA_TRY(cudaMemcpyAsync(ix->d_col, col, nnz * sizeof(uint32_t), kind, ix->stream));
  ix->nnz = nnz;
  if (ix->head_cap < (ix->n + 1) * NGTGPU_HEAD_WIDTH) {
    if (ix->d_head) cudaFree(ix->d_head);
    ix->d_head = nullptr, ix->head_cap = 0;
    CUDA_TRY(cudaMalloc(&ix->d_head, (ix->n + 1) * NGTGPU_HEAD_WIDTH * sizeof(uint32_t)));
    ix->head_cap = (ix->n + 1) * NGTGPU_HEAD_WIDTH;
  }
  build_head_kernel<<<ix->sm_count * 8, 256, 0, ix->stream>>>(ix->d_row_ptr, ix->d_col, ix->n, ix->d_head);
  ix->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_search_property(ngtgpu_index *ix, int64_t edge_size_for_search,
                                                int64_t dynamic_edge_size_base, int64_t dynamic_edge_size_rate) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ix->edge_size_for_search = edge_size_for_search;
  ix->dyn_base = dynamic_edge_size_base;
  ix->dyn_rate = dynamic_edge_size_rate;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_search_workspace(ngtgpu_index *ix, uint32_t hash_bits, uint32_t queue_cap) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  if (hash_bits < 8 || hash_bits > 17) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "hash_bits must be in [8, 17]");
  if (queue_cap < 64 || queue_cap > 8192) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "queue_cap must be in [64, 8192]");
  ix->hash_bits = hash_bits;
  ix->hash_bits_auto = false;
  ix->queue_cap = queue_cap;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_fast_shape(ngtgpu_index *ix, int warps_per_query, int ctas_per_sm) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  if (warps_per_query != 0 && warps_per_query != 1 && warps_per_query != 2 && warps_per_query != 4)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "warps per query must be 0 (by row width), 1, 2 or 4");
  if (ctas_per_sm < 0 || ctas_per_sm > 32) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "CTAs per SM must be in [0, 32]");
  ix->fast_warps = warps_per_query;
  ix->fast_ctas_per_sm = ctas_per_sm;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_tensor_core(ngtgpu_index *ix, int enabled) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ix->tc_enabled = enabled != 0;
  return NGTGPU_OK;
}
extern "C" uint64_t ngtgpu_index_tensor_core_batches(const ngtgpu_index *ix) { return ix ? ix->tc_batches : 0; }

extern "C" int ngtgpu_index_set_stage_bytes(ngtgpu_index *ix, uint32_t bytes) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  if (bytes < 2048 || bytes > 131072) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "stage bytes must be in [2048, 131072]");
  ix->stage_bytes = bytes;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_onchip_tiers(ngtgpu_index *ix, int tiers) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  if (tiers < 1 || tiers > 2) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "tiers must be 1 or 2");
  ix->onchip_tiers = tiers;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_seed_fusion(ngtgpu_index *ix, int enabled) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ix->fuse_seeds = enabled != 0;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_set_fast_kernel(ngtgpu_index *ix, int enabled) {
  if (!ix) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null index handle");
  ix->fast_kernel = enabled != 0;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_get_object(const ngtgpu_index *ix, uint32_t id, void *out) {
  if (!ix || !out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_get_object: null argument");
  if (!ix->d_objects || id == 0 || id > ix->n) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_get_object: no such object");
  CUDA_TRY(cudaSetDevice(ix->device));
  CUDA_TRY(cudaMemcpy(out, ix->d_objects + (size_t)id * ix->row_bytes, (size_t)ix->dim * ix->elem_size,
                      cudaMemcpyDeviceToHost));
  return NGTGPU_OK;
}

// The stored rows of objects first..first+count-1 in ONE strided copy (padded device rows -> packed host rows).
extern "C" int ngtgpu_index_get_objects(const ngtgpu_index *ix, uint32_t first, uint64_t count, void *out) {
  if (!ix || !out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_get_objects: null argument");
  if (!ix->d_objects || first == 0 || count == 0 || (uint64_t)first + count - 1 > ix->n)
    NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_get_objects: no such objects");
  CUDA_TRY(cudaSetDevice(ix->device));
  const size_t width = (size_t)ix->dim * ix->elem_size;
  CUDA_TRY(cudaMemcpy2D(out, width, ix->d_objects + (size_t)first * ix->row_bytes, ix->row_bytes, width, count,
                        cudaMemcpyDeviceToHost));
  return NGTGPU_OK;
}

// Copies the padded device rows of objects first..first+count-1 (device pointer out) -- used by shards.
extern "C" const void *ngtgpu_index_device_objects(const ngtgpu_index *ix) { return ix ? ix->d_objects : nullptr; }

// ---- seed table ----------------------------------------------------------------------------------------
__global__ void gather_pivots_kernel(const uint8_t *__restrict__ objects, uint32_t row_bytes,
                                     const uint32_t *__restrict__ ids, uint32_t n_pivots,
                                     uint8_t *__restrict__ out) {
  const uint32_t chunks = row_bytes / 16;
  uint64_t total = (uint64_t)n_pivots * chunks;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t p = (uint32_t)(i / chunks), c = (uint32_t)(i % chunks);
    reinterpret_cast<uint4 *>(out)[i] =
        __ldg(reinterpret_cast<const uint4 *>(objects + (size_t)ids[p] * row_bytes) + c);
  }
}

static inline uint64_t splitmix64(uint64_t &x) {
  uint64_t z = (x += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

static int install_pivots(ngtgpu_index *ix, const std::vector<uint32_t> &ids) {
  const uint32_t n_pivots = (uint32_t)ids.size();
  if (ix->pivot_cap < n_pivots) {
    free_pivots(ix);
    CUDA_TRY(cudaMalloc(&ix->d_pivot_ids, n_pivots * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&ix->d_pivot_rows, (size_t)n_pivots * ix->row_bytes));
    ix->pivot_cap = n_pivots;
  }
  CUDA_TRY(cudaMemcpy(ix->d_pivot_ids, ids.data(), n_pivots * sizeof(uint32_t), cudaMemcpyHostToDevice));
  gather_pivots_kernel<<<ix->sm_count * 4, 256, 0, ix->stream>>>(ix->d_objects, ix->row_bytes, ix->d_pivot_ids,
                                                                 n_pivots, ix->d_pivot_rows);
  ix->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  ix->n_pivots = n_pivots;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_index_build_seed_table(ngtgpu_index *ix, uint32_t n_pivots, uint64_t rng_seed) {
  return ngtgpu_index_build_seed_table_range(ix, n_pivots, rng_seed, 0);
}

// pivots sampled from ids 1..limit only (limit == 0: all objects): while a graph is being grown batch by batch the
// seeds must be nodes that are already in it
extern "C" int ngtgpu_index_build_seed_table_range(ngtgpu_index *ix, uint32_t n_pivots, uint64_t rng_seed, uint64_t limit) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!ix->d_objects || ix->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_build_seed_table: objects are not set");
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  ix->n_pivots = 0;   // (the buffers stay: install_pivots reuses them)
  if (n_pivots == 0) {
    free_pivots(ix);
    return NGTGPU_OK;
  }
  const uint64_t range = (limit == 0 || limit > ix->n) ? ix->n : limit;
  if (n_pivots > range) n_pivots = (uint32_t)range;
  // evenly strided sample with a random phase per stride: distinct ids, spread over the id range
  std::vector<uint32_t> ids(n_pivots);
  std::vector<uint8_t> valid;
  if (ix->d_valid) {
    valid.resize(ix->n + 1);
    CUDA_TRY(cudaMemcpy(valid.data(), ix->d_valid, ix->n + 1, cudaMemcpyDeviceToHost));
  }
  uint64_t st = rng_seed;
  uint32_t kept = 0;
  for (uint32_t p = 0; p < n_pivots; p++) {
    uint64_t lo = (uint64_t)p * range / n_pivots, hi = (uint64_t)(p + 1) * range / n_pivots;
    if (hi <= lo) continue;
    uint64_t id = 1 + lo + splitmix64(st) % (hi - lo);
    if (!valid.empty()) {
      uint64_t tries = hi - lo;
      while (tries-- && !valid[id]) id = 1 + lo + (id - lo) % (hi - lo);
      if (!valid[id]) continue;
    }
    ids[kept++] = (uint32_t)id;
  }
  if (kept == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_build_seed_table: no valid objects");
  ids.resize(kept);
  return install_pivots(ix, ids);
}

// An explicit pivot list. ids 1..seedSize is the reference's SeedTypeFixedNodes (lib/NGT/Index.h:1122-1127): with
// n_seeds == the table size every search starts from exactly those nodes.
extern "C" int ngtgpu_index_set_seed_table_ids(ngtgpu_index *ix, const uint32_t *pivot_ids, uint32_t n_pivots) {
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (!ix->d_objects || ix->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_index_set_seed_table_ids: objects are not set");
  if (!pivot_ids || n_pivots == 0) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_seed_table_ids: empty pivot list");
  for (uint32_t i = 0; i < n_pivots; i++)
    if (pivot_ids[i] == 0 || pivot_ids[i] > ix->n)
      NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_index_set_seed_table_ids: id " + std::to_string(pivot_ids[i]) + " out of range");
  CUDA_TRY(cudaStreamSynchronize(ix->stream));
  ix->n_pivots = 0;
  return install_pivots(ix, std::vector<uint32_t>(pivot_ids, pivot_ids + n_pivots));
}

// NeighborhoodGraph::getEdgeSize, lib/NGT/Graph.h:675-692. Returns < 0 for invalid parameters.
int64_t ngtgpu_effective_edge_size(const ngtgpu_index *ix, const ngtgpu_search_params *p) {
  int64_t esize = p->edge_size == -1 ? ix->edge_size_for_search : p->edge_size;
  const int64_t all = 0x7fffffff;
  if (esize == 0) return all;
  if (esize > 0) return esize > all ? all : esize;
  if (esize == -2) {
    float coef = (float)((double)p->epsilon + 1.0);  // Common.h:2041
    double add = pow(10.0, ((double)coef - 1.0) * (double)(float)ix->dyn_rate);
    if (add >= (double)all) return all;
    double v = (double)ix->dyn_base + add;
    if (v < 0.0) return -1;
    return (int64_t)v;
  }
  return -1;
}

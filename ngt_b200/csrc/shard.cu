// shard.cu -- the row-sharded index: one shard per GPU, per-shard top-k lists exchanged with ONE NCCL all-gather and
// merged on the device (the reference has no distribution, bin/ngt/README.md:30).
//
// Shard g owns a contiguous block of the objects as its own index (own graph, own seed table; no cross-shard edges;
// global id = local id + id_offset). Every shard answers the full query batch; the traversal kernel writes its results
// as 64-bit keys (ordered distance bits << 32 | global id) STRAIGHT INTO this rank's slot of the gather buffer (no pack
// pass), ncclAllGather runs in place over NVLink, and scan_merge_kernel consumes the gathered buffer directly, keeping the
// k smallest keys per query -- the order of ObjectDistance (lib/NGT/Common.h:1946-1952), i.e. what a single priority queue
// over the union keeps. So merged results equal a search of the union bit for bit.
//
// Two hosts are served by the same code:
//   * one process per GPU (torchrun, MPI): ngtgpu_comm_* + ngtgpu_shard_*_device; the launcher only carries the 128-byte
//     NCCL id from rank 0 to the others;
//   * one process driving several GPUs (what a program written against lib/NGT/Capi.h is): ngtgpu_sharded_*
//     (ncclCommInitAll, grouped collectives), also behind ngt_open_index when NGTGPU_DEVICES names several devices.
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already in the process when torch loaded one), so the
// single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "ngtgpu_internal.cuh"

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *env = getenv("NGTGPU_NCCL_SO");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = std::string("NCCL is not loadable (libnccl.so.2; set NGTGPU_NCCL_SO): ") + (dlerror() ? dlerror() : "");
      return;
    }
#define NCCL_SYM(field, name)                                                      \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));     \
  if (!api.field) api.error = std::string("NCCL symbol missing: ") + name;
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommInitAll, "ncclCommInitAll")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllGather, "ncclAllGather")
    NCCL_SYM(Broadcast, "ncclBroadcast")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
  });
  return &api;
}

#define NCCL_TRY(expr)                                                                          \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != ncclSuccess) {                                                                    \
      ngtgpu_set_error(std::string(#expr) + ": " + nccl_api()->GetErrorString(_r));             \
      return NGTGPU_ERR_CUDA;                                                                   \
    }                                                                                           \
  } while (0)

int need_nccl(NcclApi **out) {
  NcclApi *a = nccl_api();
  if (!a->error.empty()) NGTGPU_FAIL(NGTGPU_ERR_STATE, a->error);
  *out = a;
  return NGTGPU_OK;
}

// device buffer that only grows
struct Grow {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t need(size_t b) {
    if (b <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, b);
    if (e == cudaSuccess) bytes = b;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// one process per GPU
struct ngtgpu_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  Grow gather, counts, ids, dists;
  bool timing = false;
  std::vector<cudaEvent_t> events;   // groups of four: start, after the shard's search, after the all-gather, after the merge
};

extern "C" int ngtgpu_comm_get_unique_id(void *id_out) {
  if (!id_out) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_comm_get_unique_id: null buffer");
  NcclApi *nc;
  NGTGPU_TRY(need_nccl(&nc));
  static_assert(sizeof(ncclUniqueId) == NGTGPU_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCCL_TRY(nc->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return NGTGPU_OK;
}

extern "C" int ngtgpu_comm_create(ngtgpu_comm **out, const void *id128, int rank, int world, int device) {
  if (!out || !id128) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_comm_create: null argument");
  if (world < 1 || world > 32 || rank < 0 || rank >= world) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_comm_create: rank / world out of range (1..32 ranks)");
  NcclApi *nc;
  NGTGPU_TRY(need_nccl(&nc));
  CUDA_TRY(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ngtgpu_comm *c = new ngtgpu_comm;
  c->rank = rank, c->world = world, c->device = device;
  ncclResult_t r = nc->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("ncclCommInitRank: ") + nc->GetErrorString(r));
  }
  *out = c;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_comm_destroy(ngtgpu_comm *c) {
  if (!c) return NGTGPU_OK;
  cudaSetDevice(c->device);
  for (cudaEvent_t e : c->events) cudaEventDestroy(e);
  c->gather.release();
  c->counts.release();
  c->ids.release();
  c->dists.release();
  if (c->comm) nccl_api()->CommDestroy(c->comm);
  delete c;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_comm_set_timing(ngtgpu_comm *c, int enabled) {
  if (!c) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null communicator");
  c->timing = enabled != 0;
  return NGTGPU_OK;
}

// Sums the recorded calls: ms[0] this shard's search (query preparation + seed selection + traversal or scan),
// ms[1] all-gather (includes waiting for the slowest shard), ms[2] merge.
extern "C" int ngtgpu_comm_pop_timing(ngtgpu_comm *c, double *ms3, uint64_t *calls) {
  if (!c || !ms3 || !calls) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_comm_pop_timing: null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  ms3[0] = ms3[1] = ms3[2] = 0.0;
  *calls = 0;
  for (size_t i = 0; i + 3 < c->events.size(); i += 4) {
    CUDA_TRY(cudaEventSynchronize(c->events[i + 3]));
    for (int j = 0; j < 3; j++) {
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, c->events[i + j], c->events[i + j + 1]));
      ms3[j] += ms;
    }
    (*calls)++;
  }
  for (cudaEvent_t e : c->events) cudaEventDestroy(e);
  c->events.clear();
  return NGTGPU_OK;
}

namespace {

int mark(ngtgpu_comm *c, cudaStream_t stream) {
  if (!c->timing) return NGTGPU_OK;
  cudaEvent_t e;
  CUDA_TRY(cudaEventCreate(&e));
  CUDA_TRY(cudaEventRecord(e, stream));
  c->events.push_back(e);
  return NGTGPU_OK;
}

// this rank's keys are in slot `rank` of the gather buffer: all-gather in place, merge
int exchange_and_merge(ngtgpu_comm *c, uint32_t nq, uint32_t k, uint32_t *ids, float *dists, uint32_t *counts, cudaStream_t stream) {
  NcclApi *nc = nccl_api();
  uint64_t *g = static_cast<uint64_t *>(c->gather.p);
  const size_t per = (size_t)nq * k;
  if (c->world > 1) NCCL_TRY(nc->AllGather(g + (size_t)c->rank * per, g, per, ncclUint64, c->comm, stream));
  NGTGPU_TRY(mark(c, stream));
  NGTGPU_TRY(ngtgpu_merge_keys(g, (uint32_t)c->world, nq, k, ids, dists, counts, stream));
  NGTGPU_TRY(mark(c, stream));
  return NGTGPU_OK;
}

}  // namespace

extern "C" int ngtgpu_shard_search_device(ngtgpu_index *ix, ngtgpu_comm *c, const void *queries, int query_type, uint32_t nq,
                                          const ngtgpu_search_params *params, uint32_t n_seeds, uint32_t id_offset,
                                          uint32_t *ids, float *dists, uint32_t *counts, void *stream_) {
  if (!ix || !c || !params) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_search_device: null argument");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (ix->device != c->device) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_search_device: index and communicator are on different devices");
  cudaStream_t stream = (cudaStream_t)stream_;
  const uint32_t k = params->size;
  if (nq == 0) return NGTGPU_OK;
  if (!counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_search_device: null buffer");
  if (k == 0) {
    CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)nq * 4, stream));
    return NGTGPU_OK;
  }
  if (!ids || !dists) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_search_device: null result buffer");
  const size_t per = (size_t)nq * k;
  CUDA_TRY(c->gather.need(per * 8 * c->world));
  CUDA_TRY(c->counts.need((size_t)nq * 4));
  NGTGPU_TRY(mark(c, stream));
  NGTGPU_TRY(ngtgpu_search_keys_device(ix, queries, query_type, nq, params, n_seeds, id_offset,
                                       static_cast<uint64_t *>(c->gather.p) + (size_t)c->rank * per,
                                       static_cast<uint32_t *>(c->counts.p), stream));
  NGTGPU_TRY(mark(c, stream));
  return exchange_and_merge(c, nq, k, ids, dists, counts, stream);
}

extern "C" int ngtgpu_shard_linear_search_device(ngtgpu_index *ix, ngtgpu_comm *c, const void *queries, int query_type,
                                                 uint32_t nq, uint32_t size, float radius, uint32_t id_offset, uint32_t *ids,
                                                 float *dists, uint32_t *counts, void *stream_) {
  if (!ix || !c) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_linear_search_device: null argument");
  NGTGPU_TRY(ngtgpu_check_device(ix));
  if (ix->device != c->device) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_linear_search_device: index and communicator are on different devices");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (nq == 0) return NGTGPU_OK;
  if (!counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_linear_search_device: null buffer");
  if (size == 0) {
    CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)nq * 4, stream));
    return NGTGPU_OK;
  }
  if (!ids || !dists) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_shard_linear_search_device: null result buffer");
  const size_t per = (size_t)nq * size;
  CUDA_TRY(c->gather.need(per * 8 * c->world));
  CUDA_TRY(c->counts.need((size_t)nq * 4));
  CUDA_TRY(c->ids.need(per * 4));
  CUDA_TRY(c->dists.need(per * 4));
  NGTGPU_TRY(mark(c, stream));
  uint32_t *l_ids = static_cast<uint32_t *>(c->ids.p), *l_cnt = static_cast<uint32_t *>(c->counts.p);
  float *l_d = static_cast<float *>(c->dists.p);
  NGTGPU_TRY(ngtgpu_linear_search_device(ix, queries, query_type, nq, size, radius, l_ids, l_d, l_cnt, stream));
  NGTGPU_TRY(ngtgpu_pack_keys(l_ids, l_d, l_cnt, nq, size, id_offset, static_cast<uint64_t *>(c->gather.p) + (size_t)c->rank * per, stream));
  NGTGPU_TRY(mark(c, stream));
  return exchange_and_merge(c, nq, size, ids, dists, counts, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// one process, several GPUs
struct ngtgpu_sharded {
  std::vector<int> devices;
  std::vector<ngtgpu_index *> shards;
  std::vector<uint64_t> offset, count;   // global id = local id + offset[g]
  std::vector<ncclComm_t> comms;
  std::vector<Grow> gather, counts, queries, l_ids, l_dists;
  Grow out_ids, out_dists, out_counts;   // merged results on devices[0]
  int object_type = 0, distance_type = 0;
  uint32_t dim = 0;
  uint64_t n = 0;
  double last_ms[5] = {0, 0, 0, 0, 0};   // upload + broadcast, search, all-gather, merge, download (devices[0]'s stream)
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::mutex mtx;                        // searches on one handle are serialised (the reference allows concurrent callers)
};

extern "C" int ngtgpu_sharded_destroy(ngtgpu_sharded *s) {
  if (!s) return NGTGPU_OK;
  for (size_t g = 0; g < s->devices.size(); g++) {
    cudaSetDevice(s->devices[g]);
    if (g < s->gather.size()) {
      s->gather[g].release();
      s->counts[g].release();
      s->queries[g].release();
      s->l_ids[g].release();
      s->l_dists[g].release();
    }
    if (g < s->comms.size() && s->comms[g]) nccl_api()->CommDestroy(s->comms[g]);
    if (g < s->shards.size() && s->shards[g]) ngtgpu_index_destroy(s->shards[g]);
    if (g == 0) {
      s->out_ids.release();
      s->out_dists.release();
      s->out_counts.release();
      for (auto &e : s->ev)
        if (e) cudaEventDestroy(e);
    }
  }
  delete s;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_sharded_create(ngtgpu_sharded **out, const int *devices, int n_devices, int object_type, int distance_type,
                                     uint32_t dimension) {
  if (!out || !devices) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_create: null argument");
  if (n_devices < 1 || n_devices > 32) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_create: 1..32 devices");
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_create: a device is listed twice");
  NcclApi *nc = nullptr;
  if (n_devices > 1) NGTGPU_TRY(need_nccl(&nc));
  ngtgpu_sharded *s = new ngtgpu_sharded;
  s->object_type = object_type, s->distance_type = distance_type, s->dim = dimension;
  s->devices.assign(devices, devices + n_devices);
  s->shards.assign(n_devices, nullptr);
  s->offset.assign(n_devices, 0);
  s->count.assign(n_devices, 0);
  s->gather.resize(n_devices), s->counts.resize(n_devices), s->queries.resize(n_devices);
  s->l_ids.resize(n_devices), s->l_dists.resize(n_devices);
  for (int g = 0; g < n_devices; g++) {
    int rc = ngtgpu_index_create(&s->shards[g], devices[g], object_type, distance_type, dimension);
    if (rc != NGTGPU_OK) {
      ngtgpu_sharded_destroy(s);
      return rc;
    }
  }
  if (n_devices > 1) {
    s->comms.assign(n_devices, nullptr);
    ncclResult_t r = nc->CommInitAll(s->comms.data(), n_devices, s->devices.data());
    if (r != ncclSuccess) {
      ngtgpu_sharded_destroy(s);
      NGTGPU_FAIL(NGTGPU_ERR_CUDA, std::string("ncclCommInitAll: ") + nc->GetErrorString(r));
    }
  }
  cudaSetDevice(devices[0]);
  for (auto &e : s->ev) cudaEventCreate(&e);
  *out = s;
  return NGTGPU_OK;
}

extern "C" int ngtgpu_sharded_shard_count(const ngtgpu_sharded *s) { return s ? (int)s->devices.size() : 0; }

extern "C" int ngtgpu_sharded_shard(ngtgpu_sharded *s, int shard, ngtgpu_index **index, uint64_t *id_offset, uint64_t *count) {
  if (!s || shard < 0 || shard >= (int)s->devices.size()) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_shard: no such shard");
  if (index) *index = s->shards[shard];
  if (id_offset) *id_offset = s->offset[shard];
  if (count) *count = s->count[shard];
  return NGTGPU_OK;
}

// rows [g*n/G, (g+1)*n/G) go to shard g
extern "C" int ngtgpu_sharded_set_objects(ngtgpu_sharded *s, const void *objects, uint64_t n, int normalize) {
  if (!s || !objects) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_set_objects: null argument");
  const uint64_t G = s->devices.size();
  if (n < G) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_set_objects: fewer objects than shards");
  if (n >= 0xffffffffull) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_set_objects: ids do not fit 32 bits");
  const size_t rec = (size_t)s->dim * (s->object_type == NGTGPU_OBJECT_UINT8 ? 1 : 4);
  for (uint64_t g = 0; g < G; g++) {
    const uint64_t lo = g * n / G, hi = (g + 1) * n / G;
    s->offset[g] = lo;
    s->count[g] = hi - lo;
    NGTGPU_TRY(ngtgpu_index_set_objects(s->shards[g], static_cast<const uint8_t *>(objects) + lo * rec, hi - lo, normalize, 0));
  }
  s->n = n;
  return NGTGPU_OK;
}

// every shard builds its own ONNG (ngtgpu_index_build_onng) and seed table, all devices at once (a host thread each)
extern "C" int ngtgpu_sharded_build_onng(ngtgpu_sharded *s, uint32_t knn, uint32_t outgoing, uint32_t incoming,
                                         int shortcut_reduction, int64_t edge_size_for_search, uint32_t n_pivots) {
  if (!s) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "null sharded handle");
  if (s->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "ngtgpu_sharded_build_onng: objects are not set");
  const size_t G = s->devices.size();
  std::vector<int> rc(G, NGTGPU_OK);
  std::vector<std::string> msg(G);
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++)
    th.emplace_back([&, g] {
      ngtgpu_index *ix = s->shards[g];
      int r = ngtgpu_index_build_onng(ix, knn, outgoing, incoming, shortcut_reduction, 0, nullptr, nullptr);
      if (r == NGTGPU_OK) r = ngtgpu_index_set_search_property(ix, edge_size_for_search, 30, 20);
      if (r == NGTGPU_OK) r = ngtgpu_index_build_seed_table(ix, (uint32_t)std::min<uint64_t>(n_pivots, s->count[g]), 1);
      rc[g] = r;
      if (r != NGTGPU_OK) msg[g] = ngtgpu_last_error();
    });
  for (auto &t : th) t.join();
  for (size_t g = 0; g < G; g++)
    if (rc[g] != NGTGPU_OK) NGTGPU_FAIL(rc[g], "shard " + std::to_string(g) + ": " + msg[g]);
  return NGTGPU_OK;
}

namespace {

// upload to devices[0], broadcast over NVLink, run `per_shard` on every device (it must leave the shard's keys in slot g
// of gather[g]), all-gather in place, merge on devices[0], download
template <typename F>
int sharded_run(ngtgpu_sharded *s, const void *queries, int query_type, uint32_t nq, uint32_t k, uint32_t *ids, float *dists,
                uint32_t *counts, F per_shard) {
  std::lock_guard<std::mutex> lock(s->mtx);
  const int G = (int)s->devices.size();
  NcclApi *nc = nccl_api();
  const size_t qbytes = (size_t)nq * s->dim * (query_type == NGTGPU_OBJECT_UINT8 ? 1 : 4);
  const size_t per = (size_t)nq * k;
  for (int g = 0; g < G; g++) {
    CUDA_TRY(cudaSetDevice(s->devices[g]));
    CUDA_TRY(s->queries[g].need(qbytes));
    CUDA_TRY(s->gather[g].need(per * 8 * G));
    CUDA_TRY(s->counts[g].need((size_t)nq * 4));
  }
  CUDA_TRY(cudaSetDevice(s->devices[0]));
  CUDA_TRY(s->out_ids.need(per * 4));
  CUDA_TRY(s->out_dists.need(per * 4));
  CUDA_TRY(s->out_counts.need((size_t)nq * 4));
  cudaStream_t s0 = s->shards[0]->stream;
  CUDA_TRY(cudaEventRecord(s->ev[0], s0));
  CUDA_TRY(cudaMemcpyAsync(s->queries[0].p, queries, qbytes, cudaMemcpyHostToDevice, s0));
  if (G > 1) {
    NCCL_TRY(nc->GroupStart());
    for (int g = 0; g < G; g++)
      NCCL_TRY(nc->Broadcast(s->queries[0].p, s->queries[g].p, qbytes, ncclUint8, 0, s->comms[g], s->shards[g]->stream));
    NCCL_TRY(nc->GroupEnd());
  }
  CUDA_TRY(cudaSetDevice(s->devices[0]));
  CUDA_TRY(cudaEventRecord(s->ev[1], s0));
  for (int g = 0; g < G; g++) NGTGPU_TRY(per_shard(g));
  CUDA_TRY(cudaSetDevice(s->devices[0]));
  CUDA_TRY(cudaEventRecord(s->ev[2], s0));
  if (G > 1) {
    NCCL_TRY(nc->GroupStart());
    for (int g = 0; g < G; g++) {
      uint64_t *gb = static_cast<uint64_t *>(s->gather[g].p);
      NCCL_TRY(nc->AllGather(gb + (size_t)g * per, gb, per, ncclUint64, s->comms[g], s->shards[g]->stream));
    }
    NCCL_TRY(nc->GroupEnd());
  }
  CUDA_TRY(cudaSetDevice(s->devices[0]));
  CUDA_TRY(cudaEventRecord(s->ev[3], s0));
  uint32_t *o_ids = static_cast<uint32_t *>(s->out_ids.p), *o_cnt = static_cast<uint32_t *>(s->out_counts.p);
  float *o_d = static_cast<float *>(s->out_dists.p);
  NGTGPU_TRY(ngtgpu_merge_keys(static_cast<uint64_t *>(s->gather[0].p), (uint32_t)G, nq, k, o_ids, o_d, o_cnt, s0));
  CUDA_TRY(cudaEventRecord(s->ev[4], s0));
  CUDA_TRY(cudaMemcpyAsync(ids, o_ids, per * 4, cudaMemcpyDeviceToHost, s0));
  CUDA_TRY(cudaMemcpyAsync(dists, o_d, per * 4, cudaMemcpyDeviceToHost, s0));
  CUDA_TRY(cudaMemcpyAsync(counts, o_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, s0));
  CUDA_TRY(cudaEventRecord(s->ev[5], s0));
  CUDA_TRY(cudaStreamSynchronize(s0));
  for (int g = 1; g < G; g++) {   // the other shards' all-gathers are done too before their buffers are reused
    CUDA_TRY(cudaSetDevice(s->devices[g]));
    CUDA_TRY(cudaStreamSynchronize(s->shards[g]->stream));
  }
  for (int i = 0; i < 5; i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]);
    s->last_ms[i] = ms;
  }
  return NGTGPU_OK;
}

}  // namespace

extern "C" int ngtgpu_sharded_search(ngtgpu_sharded *s, const void *queries, int query_type, uint32_t nq,
                                     const ngtgpu_search_params *params, uint32_t n_seeds, uint32_t *ids, float *dists,
                                     uint32_t *counts) {
  if (!s || !params) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_search: null argument");
  if (s->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "search: the index holds no objects");
  if (nq == 0) return NGTGPU_OK;
  if (!queries || !counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: null buffer");
  const uint32_t k = params->size;
  if (k == 0) {
    memset(counts, 0, (size_t)nq * 4);
    return NGTGPU_OK;
  }
  if (!ids || !dists) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "search: null result buffer");
  const size_t per = (size_t)nq * k;
  return sharded_run(s, queries, query_type, nq, k, ids, dists, counts, [&](int g) {
    ngtgpu_index *ix = s->shards[g];
    return ngtgpu_search_keys_device(ix, s->queries[g].p, query_type, nq, params, n_seeds, (uint32_t)s->offset[g],
                                     static_cast<uint64_t *>(s->gather[g].p) + (size_t)g * per,
                                     static_cast<uint32_t *>(s->counts[g].p), ix->stream);
  });
}

extern "C" int ngtgpu_sharded_linear_search(ngtgpu_sharded *s, const void *queries, int query_type, uint32_t nq, uint32_t size,
                                            float radius, uint32_t *ids, float *dists, uint32_t *counts) {
  if (!s) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_linear_search: null argument");
  if (s->n == 0) NGTGPU_FAIL(NGTGPU_ERR_STATE, "linear search: the index holds no objects");
  if (nq == 0) return NGTGPU_OK;
  if (!queries || !counts) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "linear search: null buffer");
  if (size == 0) {
    memset(counts, 0, (size_t)nq * 4);
    return NGTGPU_OK;
  }
  if (!ids || !dists) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "linear search: null result buffer");
  const size_t per = (size_t)nq * size;
  return sharded_run(s, queries, query_type, nq, size, ids, dists, counts, [&](int g) {
    ngtgpu_index *ix = s->shards[g];
    NGTGPU_TRY(ngtgpu_check_device(ix));
    CUDA_TRY(s->l_ids[g].need(per * 4));
    CUDA_TRY(s->l_dists[g].need(per * 4));
    uint32_t *l_ids = static_cast<uint32_t *>(s->l_ids[g].p), *l_cnt = static_cast<uint32_t *>(s->counts[g].p);
    float *l_d = static_cast<float *>(s->l_dists[g].p);
    NGTGPU_TRY(ngtgpu_linear_search_device(ix, s->queries[g].p, query_type, nq, size, radius, l_ids, l_d, l_cnt, ix->stream));
    return ngtgpu_pack_keys(l_ids, l_d, l_cnt, nq, size, (uint32_t)s->offset[g],
                            static_cast<uint64_t *>(s->gather[g].p) + (size_t)g * per, ix->stream);
  });
}

// milliseconds of the last sharded call on devices[0]'s stream: upload + broadcast, search, all-gather (waits for the
// slowest shard), merge, download
extern "C" int ngtgpu_sharded_last_timing(const ngtgpu_sharded *s, double *ms5) {
  if (!s || !ms5) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_sharded_last_timing: null argument");
  for (int i = 0; i < 5; i++) ms5[i] = s->last_ms[i];
  return NGTGPU_OK;
}

// ref_shim.cpp -- thin C-ABI driver over the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).
//
// Built by oracle/Makefile against the reference's own sources compiled where they lie under
// /root/reference (outputs only in oracle/_ref/, git-ignored). It lets tests/golden/make_golden.py,
// tests/ (oracle pinning) and bench.py's cpu_baseline / `--impl reference` legs drive
// NGT::Index::search / linearSearch / createIndex / GraphOptimizer exactly as a user of the
// reference would. Nothing here is product code and nothing in
// ngt_b200/ links or loads it.
#include <NGT/Index.h>
#include <NGT/GraphOptimizer.h>
#include <NGT/GraphReconstructor.h>
#include <omp.h>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;
#define REF_TRY try {
#define REF_CATCH(ret)                                      \
  }                                                         \
  catch (std::exception & e) {                              \
    g_err = e.what();                                       \
    return ret;                                             \
  }

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// ngt create (Command.cpp:172-213) through the library API: createGraphAndTree + append + createIndex.
// objtype: 'f' float, 'c' uint8. disttype: ObjectSpace::DistanceType value. indextype: 't' graph+tree, 'g' graph.
int ref_build_index(const char *path, const float *data, size_t n, int dim, char objtype, int disttype,
                    int edge_creation, int edge_search, char indextype, int threads) {
  REF_TRY
  NGT::Property p;
  p.dimension = dim;
  p.objectType = objtype == 'c' ? NGT::ObjectSpace::ObjectType::Uint8 : NGT::ObjectSpace::ObjectType::Float;
  p.distanceType = (NGT::Index::Property::DistanceType)disttype;
  p.edgeSizeForCreation = edge_creation;
  p.edgeSizeForSearch = edge_search;
  if (indextype == 'g') {
    NGT::Index::createGraph(path, p, "", 0, true);
  } else {
    NGT::Index::createGraphAndTree(path, p, true);
  }
  NGT::Index idx(path);
  idx.disableLog();
  idx.append(data, n);
  idx.createIndex(threads);
  idx.save();
  return 0;
  REF_CATCH(-1)
}

// ngt reconstruct-graph -o out -i in (GraphOptimizer.h:230-300) without search-parameter tuning.
int ref_build_onng(const char *anng_path, const char *onng_path, int outgoing, int incoming, int shortcut_reduction) {
  REF_TRY
  NGT::GraphOptimizer go(true);
  go.set(outgoing, incoming, 100, 20);
  go.setProcessingModes(shortcut_reduction != 0, false, false, false);
  go.execute(anng_path, onng_path);
  return 0;
  REF_CATCH(-1)
}

void *ref_open(const char *path, int readonly) {
  REF_TRY
  NGT::Index *idx = new NGT::Index(path, readonly != 0);
  idx->disableLog();
  return idx;
  REF_CATCH(nullptr)
}

void ref_close(void *h) { delete static_cast<NGT::Index *>(h); }

// info[0]=repository size (n+1), [1]=dimension, [2]=padded dimension, [3]=object type, [4]=distance type,
// [5]=edgeSizeForSearch, [6]=dynamicEdgeSizeBase, [7]=dynamicEdgeSizeRate, [8]=seedSize, [9]=byte size of object
int ref_info(void *h, int64_t *info) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  NGT::Property p;
  idx.getProperty(p);
  info[0] = idx.getObjectRepositorySize();
  info[1] = p.dimension;
  info[2] = idx.getObjectSpace().getPaddedDimension();
  info[3] = p.objectType;
  info[4] = p.distanceType;
  info[5] = p.edgeSizeForSearch;
  info[6] = p.dynamicEdgeSizeBase;
  info[7] = p.dynamicEdgeSizeRate;
  info[8] = p.seedSize;
  info[9] = idx.getObjectSpace().getByteSizeOfObject();
  return 0;
  REF_CATCH(-1)
}

// Copies object `id`'s stored bytes (getByteSizeOfObject of them; normalised if the space normalises).
int ref_get_object(void *h, uint32_t id, void *out) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  NGT::ObjectSpace &os = idx.getObjectSpace();
  std::memcpy(out, os.getObject(id), os.getByteSizeOfObject());
  return 0;
  REF_CATCH(-1)
}

// Exports the writable graph repository as CSR (index must be opened with readonly=0).
// Pass col==NULL to get nnz only. row_ptr has (n+2) entries over ids 0..n.
int64_t ref_export_graph(void *h, uint64_t *row_ptr, uint32_t *col, float *dist) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  NGT::GraphIndex &g = static_cast<NGT::GraphIndex &>(idx.getIndex());
  size_t rs = g.repository.size();
  uint64_t nnz = 0;
  for (size_t id = 0; id < rs; id++) {
    if (row_ptr) row_ptr[id] = nnz;
    if (id == 0 || g.repository.isEmpty(id)) continue;
    NGT::GraphNode &node = *g.getNode(id);
    for (size_t i = 0; i < node.size(); i++) {
      if (col) { col[nnz] = node[i].id; }
      if (dist) { dist[nnz] = node[i].distance; }
      nnz++;
    }
  }
  if (row_ptr) row_ptr[rs] = nnz;
  return (int64_t)nnz;
  REF_CATCH(-1)
}

// Batched driver around Index::search (seeds==NULL: GraphAndTreeIndex::search with tree seeds,
// Index.h:1570-1577) or GraphIndex::search(sc, seeds) (Index.h:1140-1179) with explicit seeds.
// queries are float rows of `dim`; for uint8 indexes they are cast like Capi.cpp:393 does.
// stats (nullable) = per query {distanceComputationCount, visitCount}.
int ref_search(void *h, const float *queries, size_t nq, int dim, size_t k, float eps, float radius, int edge_size,
               const uint32_t *seeds, size_t nseeds, uint32_t *ids, float *dists, uint32_t *counts,
               uint64_t *stats, int threads, double *seconds) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  if (threads <= 0) threads = omp_get_max_threads();
  int failed = 0;
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
  for (long q = 0; q < (long)nq; q++) {
    try {
      std::vector<float> v(queries + (size_t)q * dim, queries + (size_t)(q + 1) * dim);
      NGT::Object *o = idx.allocateObject(v);
      NGT::SearchContainer sc(*o);
      NGT::ObjectDistances r;
      sc.setResults(&r);
      sc.setSize(k);
      sc.setRadius(radius < 0.0f ? FLT_MAX : radius);
      sc.setEpsilon(eps);
      sc.setEdgeSize(edge_size);
      sc.distanceComputationCount = 0;
      sc.visitCount = 0;
      if (seeds) {
        NGT::ObjectDistances sd;
        for (size_t i = 0; i < nseeds; i++) sd.push_back(NGT::ObjectDistance(seeds[(size_t)q * nseeds + i], 0.0));
        idx.search(sc, sd);
      } else {
        idx.search(sc);
      }
      idx.deleteObject(o);
      counts[q] = (uint32_t)r.size();
      for (size_t i = 0; i < r.size() && i < k; i++) {
        ids[(size_t)q * k + i] = r[i].id;
        dists[(size_t)q * k + i] = r[i].distance;
      }
      if (stats) {
        stats[2 * q] = sc.distanceComputationCount;
        stats[2 * q + 1] = sc.visitCount;
      }
    } catch (std::exception &e) {
#pragma omp critical
      { g_err = e.what(); failed = 1; }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return failed ? -1 : 0;
  REF_CATCH(-1)
}

// Batched driver around Index::linearSearch (Index.h:729-734 -> ObjectSpaceRepository.h:466-502).
int ref_linear_search(void *h, const float *queries, size_t nq, int dim, size_t k, float radius, uint32_t *ids,
                      float *dists, uint32_t *counts, int threads, double *seconds) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  if (threads <= 0) threads = omp_get_max_threads();
  int failed = 0;
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (long q = 0; q < (long)nq; q++) {
    try {
      std::vector<float> v(queries + (size_t)q * dim, queries + (size_t)(q + 1) * dim);
      NGT::Object *o = idx.allocateObject(v);
      NGT::SearchContainer sc(*o);
      NGT::ObjectDistances r;
      sc.setResults(&r);
      sc.setSize(k);
      sc.setRadius(radius < 0.0f ? FLT_MAX : radius);
      idx.linearSearch(sc);
      idx.deleteObject(o);
      counts[q] = (uint32_t)r.size();
      for (size_t i = 0; i < r.size() && i < k; i++) {
        ids[(size_t)q * k + i] = r[i].id;
        dists[(size_t)q * k + i] = r[i].distance;
      }
    } catch (std::exception &e) {
#pragma omp critical
      { g_err = e.what(); failed = 1; }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return failed ? -1 : 0;
  REF_CATCH(-1)
}

// Seeds the DVP-tree hands to the graph search for each query (Index.h:1524-1567). getSeedsFromTree is a
// protected member, so it is reached through a derived-class using-declaration; no reference code is altered.
namespace {
struct TreePeek : public NGT::GraphAndTreeIndex {
  using NGT::GraphAndTreeIndex::getSeedsFromTree;
};
}  // namespace
// seeds: [nq, max_seeds] (0-filled), nseeds: [nq]
int ref_tree_seeds(void *h, const float *queries, size_t nq, int dim, size_t k, uint32_t *seeds, size_t max_seeds,
                   uint32_t *nseeds) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  TreePeek &t = static_cast<TreePeek &>(static_cast<NGT::GraphAndTreeIndex &>(idx.getIndex()));
  for (size_t q = 0; q < nq; q++) {
    std::vector<float> v(queries + q * dim, queries + (q + 1) * dim);
    NGT::Object *o = idx.allocateObject(v);
    NGT::SearchContainer sc(*o);
    sc.setSize(k);
    NGT::ObjectDistances sd;
    t.getSeedsFromTree(sc, sd);
    idx.deleteObject(o);
    nseeds[q] = (uint32_t)sd.size();
    for (size_t i = 0; i < sd.size() && i < max_seeds; i++) seeds[q * max_seeds + i] = sd[i].id;
  }
  return 0;
  REF_CATCH(-1)
}


// ---- construction path, made deterministic with the reference's OWN switches only:
// a graph-only index (Index::createGraph) whose SeedType property is FixedNodes, so every search of the build loop
// starts from ids 1..seedSize (Index.h:1122-1127) instead of rand() / the growing DVP-tree. createIndex(threads > 1)
// is the batched loop (Index.cpp:721-792: searchMultipleQueryForCreation on the frozen graph, then
// insertMultipleSearchResults sorted by batchIdx, then insertNode -> insertANNGNode/addEdge, Graph.h:611-626,845-886);
// its result does not depend on thread scheduling.
// The first n_first rows are appended and indexed, then the remaining rows are appended and indexed into the finished
// graph (incremental insertion through the same loop). n_first == n: one pass.
int ref_build_anng_fixed_seeds(const char *path, const float *data, size_t n, size_t n_first, int dim, char objtype,
                               int disttype, int edge_creation, int edge_search, int seed_size, int batch_size,
                               int threads) {
  REF_TRY
  NGT::Property p;
  p.dimension = dim;
  p.objectType = objtype == 'c' ? NGT::ObjectSpace::ObjectType::Uint8 : NGT::ObjectSpace::ObjectType::Float;
  p.distanceType = (NGT::Index::Property::DistanceType)disttype;
  p.edgeSizeForCreation = edge_creation;
  p.edgeSizeForSearch = edge_search;
  p.seedType = NGT::NeighborhoodGraph::SeedTypeFixedNodes;
  p.seedSize = seed_size;
  p.batchSizeForCreation = batch_size;
  NGT::Index::createGraph(path, p, "", 0, true);
  NGT::Index idx(path);
  idx.disableLog();
  if (n_first > n) n_first = n;
  idx.append(data, n_first);
  idx.createIndex(threads);
  if (n_first < n) {
    idx.append(data + n_first * (size_t)dim, n - n_first);
    idx.createIndex(threads);
  }
  idx.save();
  return 0;
  REF_CATCH(-1)
}

// GraphReconstructor::refineANNG (GraphReconstructor.h:814-924) on an open writable index; searches take the
// index's own seeds (FixedNodes for indexes made by ref_build_anng_fixed_seeds), so the result is deterministic.
int ref_refine_anng(void *h, float epsilon, float accuracy, int no_of_edges, int explore_edge_size, size_t batch_size) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  NGT::GraphReconstructor::refineANNG(idx, epsilon, accuracy, no_of_edges, explore_edge_size, batch_size);
  return 0;
  REF_CATCH(-1)
}

// insertMultipleSearchResults's in-batch step in isolation cannot be reached (file-static, Index.cpp:670), but its
// two halves can: the comparator the step calls (ObjectSpace::getComparator()(a, b)) on two stored objects ...
double ref_object_distance(void *h, uint32_t a, uint32_t b) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  NGT::ObjectSpace &os = idx.getObjectSpace();
  return os.getComparator()(*os.getRepository().get(a), *os.getRepository().get(b));
  REF_CATCH(-1.0)
}
// ... and insertNode (GraphIndex::insertNode -> insertANNGNode, Graph.h:611-626) with a caller-given result list.
int ref_insert_node(void *h, uint32_t id, const uint32_t *ids, const float *dists, size_t count) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  NGT::GraphIndex &g = static_cast<NGT::GraphIndex &>(idx.getIndex());
  NGT::ObjectDistances r;
  for (size_t i = 0; i < count; i++) r.push_back(NGT::ObjectDistance(ids[i], dists[i]));
  g.insertNode(id, r);
  return 0;
  REF_CATCH(-1)
}

// GraphOptimizer::execute with the accuracy-table step on (GraphOptimizer.h:230-372); the timed tuning steps stay off.
int ref_build_onng_with_accuracy_table(const char *anng_path, const char *onng_path, int outgoing, int incoming,
                                       int shortcut_reduction, int n_queries, int n_results) {
  REF_TRY
  NGT::GraphOptimizer go(true);
  go.set(outgoing, incoming, n_queries, n_results);
  go.setProcessingModes(shortcut_reduction != 0, false, false, true);
  go.execute(anng_path, onng_path);
  return 0;
  REF_CATCH(-1)
}

// NGT::Index::remove(id) (Index.h:463 -> GraphIndex::remove, Index.h:803-815 -> removeEdgesReliably, Graph.cpp:641-864)
int ref_remove(void *h, uint32_t id) {
  REF_TRY
  NGT::Index &idx = *static_cast<NGT::Index *>(h);
  idx.remove(id);
  return 0;
  REF_CATCH(-1)
}

// Index::AccuracyTable (Index.h:293-360): set(string) + getEpsilon(accuracy), standalone (what GraphIndex::search
// applies when expectedAccuracy > 0, Index.h:1156-1158). Returns NaN and sets the error text when the table throws.
float ref_epsilon_from_accuracy_table(const char *table, double accuracy) {
  REF_TRY
  NGT::Index::AccuracyTable t{std::string(table)};
  return t.getEpsilon(accuracy);
  REF_CATCH(std::nanf(""))
}

int ref_max_threads() { return omp_get_max_threads(); }

}  // extern "C"

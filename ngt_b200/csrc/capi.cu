// capi.cu -- NGT's C API (lib/NGT/Capi.h:60-212) over the B200 engine, for the hot path and its life cycle.
//
// Same names, argument meaning, sentinels and error convention as the reference (lib/NGT/Capi.cpp): every
// function catches, writes "Capi : <func>() : Error: <what>" into *error (a std::string) when one is given, else to
// stderr (Capi.cpp:25-38), and returns false / NULL / 0. Handles are opaque pointers owned by the caller and freed
// with the matching ngt_destroy_* / ngt_close_index. A program written against libngt's C API (python/ngt/base.py
// binds exactly these) links against this library unchanged for open / search / linear search / insert / build / save;
// single-query calls are batches of one, and two additive batch entry points feed the device properly.
// Host code only: all distance work goes through the ngtgpu_* C ABI (include/ngtgpu.h). What is outside the hot
// path (optimizer, refine_anng, in-memory-only tree functions) fails loudly with a message instead of pretending.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>
#include <unistd.h>

#include "ngtgpu_internal.cuh"

extern "C" {
int ngtgpu_io_obj_info(const char *path, uint32_t record_bytes, uint64_t *slots, uint64_t *present);
int ngtgpu_io_read_obj(const char *path, uint32_t record_bytes, void *rows, uint8_t *present);
int ngtgpu_io_write_obj(const char *path, uint32_t record_bytes, const void *rows, uint64_t n, const uint8_t *present);
int ngtgpu_io_grp_info(const char *path, uint64_t *slots, uint64_t *nnz);
int ngtgpu_io_read_grp(const char *path, uint64_t *row_ptr, uint32_t *col, float *dist, uint8_t *present);
int ngtgpu_io_write_grp(const char *path, uint64_t n, const uint64_t *row_ptr, const uint32_t *col, const float *dist,
                        const uint8_t *present);
int ngtgpu_index_knn_graph(ngtgpu_index *ix, uint32_t k, uint32_t first_id, uint32_t count, uint32_t *d_ids,
                           float *d_dists, uint32_t *d_counts, void *stream);
}

typedef unsigned int ObjectID;
typedef void *NGTIndex;
typedef void *NGTProperty;
typedef void *NGTObjectSpace;
typedef void *NGTObjectDistances;
typedef void *NGTError;
typedef void *NGTOptimizer;
typedef struct {
  ObjectID id;
  float distance;
} NGTObjectDistance;
typedef struct {
  float *query;
  size_t size;
  float epsilon;
  float accuracy;
  float radius;
  size_t edge_size;
} NGTQuery;

typedef struct {   // Capi.h:49-58
  size_t no_of_queries;
  size_t no_of_results;
  size_t no_of_threads;
  float target_accuracy;
  size_t target_no_of_objects;
  size_t no_of_sample_objects;
  size_t max_of_no_of_edges;
  bool log;
} NGTAnngEdgeOptimizationParameter;

// Index::AccuracyTable (lib/NGT/Index.h:293-360): `table` is "epsilon:accuracy,..." as the optimizer's tuning wrote it
// into `prf` (GraphOptimizer.h:355-365); getEpsilon interpolates linearly between the two entries around the asked
// accuracy (the last two when it is above the table) and clamps at -0.9. Host arithmetic only.
extern "C" int ngtgpu_epsilon_from_accuracy_table(const char *table_string, double accuracy, float *epsilon) {
  if (!table_string || !epsilon) NGTGPU_FAIL(NGTGPU_ERR_INVALID, "ngtgpu_epsilon_from_accuracy_table: null argument");
  std::vector<std::pair<float, double>> table;
  const std::string str(table_string);
  std::vector<std::string> tokens;
  {
    std::stringstream ss(str);
    std::string tok;
    while (std::getline(ss, tok, ','))
      if (!tok.empty()) tokens.push_back(tok);
  }
  if (tokens.size() >= 2) {   // fewer than two tokens: the table stays empty (Index.h:303-305)
    for (auto &t : tokens) {
      const size_t c = t.find(':');
      if (c == std::string::npos || t.find(':', c + 1) != std::string::npos)
        NGTGPU_FAIL(NGTGPU_ERR_INVALID, "AccuracyTable: Invalid accuracy table string " + t + ":" + str);
      table.push_back(std::make_pair((float)strtod(t.substr(0, c).c_str(), nullptr), strtod(t.substr(c + 1).c_str(), nullptr)));
    }
  }
  if (table.size() <= 2)
    NGTGPU_FAIL(NGTGPU_ERR_STATE, "AccuracyTable: The accuracy table is not set yet. The table size=" + std::to_string(table.size()));
  if (accuracy > 1.0) accuracy = 1.0;
  size_t i = 0;
  for (; i < table.size(); ++i)
    if (table[i].second >= accuracy) break;
  if (i == table.size()) {
    i -= 2;
  } else if (i != 0) {
    i--;
  }
  const std::pair<float, double> lower = table[i], upper = table[i + 1];
  float e = lower.first + (upper.first - lower.first) * (accuracy - lower.second) / (upper.second - lower.second);
  if (e < -0.9) e = -0.9;
  *epsilon = e;
  return NGTGPU_OK;
}

namespace {

struct CapiProperty {   // the part of NGT::Property the C API exposes (Index.h:60-154, Graph.h:385-454)
  int32_t dimension = 0;
  int16_t edge_size_for_creation = 10;
  int16_t edge_size_for_search = 40;
  int32_t object_type = NGTGPU_OBJECT_FLOAT;
  int32_t distance_type = NGTGPU_DISTANCE_L2;
};

struct CapiIndex {
  std::string path;
  std::map<std::string, std::string> prf;
  CapiProperty prop;
  std::vector<uint8_t> objects;    // ids 1..n, unpadded rows as `obj` stores them
  std::vector<uint8_t> present;    // n+1
  std::vector<uint64_t> row_ptr;   // n+2
  std::vector<uint32_t> col;
  std::vector<float> dist;
  ngtgpu_index *gpu = nullptr;
  ngtgpu_sharded *sharded = nullptr;   // opened over several GPUs (NGTGPU_DEVICES / ngt_open_index_sharded): read-only
  size_t pending = 0;              // appended since the last build
  size_t raw_from = 0;             // first id whose row is not normalised yet (Normalized* types)
  uint64_t num_dist = 0;           // distance computations of the single-query searches so far (SearchContainer::distanceComputationCount)
  size_t record_bytes() const { return (size_t)prop.dimension * (prop.object_type == NGTGPU_OBJECT_UINT8 ? 1 : 4); }
  size_t n() const { return present.empty() ? 0 : present.size() - 1; }
};

void check(int rc) {
  if (rc != NGTGPU_OK) throw std::runtime_error(ngtgpu_last_error());
}

void operate_error_string(const std::stringstream &ss, NGTError error) {   // Capi.cpp:25-38
  if (error != NULL) {
    try {
      *static_cast<std::string *>(error) = ss.str();
    } catch (...) {
      std::cerr << ss.str() << " > Failed to track error details" << std::endl;
    }
  } else {
    std::cerr << ss.str() << std::endl;
  }
}
#define CAPI_CATCH(ret)                                                   \
  catch (std::exception & err) {                                          \
    std::stringstream ss;                                                 \
    ss << "Capi : " << __FUNCTION__ << "() : Error: " << err.what();      \
    operate_error_string(ss, error);                                      \
    return ret;                                                           \
  }

const char *object_type_name(int t) { return t == NGTGPU_OBJECT_UINT8 ? "Integer-1" : "Float-4"; }
int object_type_of(const std::string &s) {
  if (s == "Integer-1") return NGTGPU_OBJECT_UINT8;
  if (s == "Float-4") return NGTGPU_OBJECT_FLOAT;
  throw std::runtime_error("Invalid Object Type in the property. " + s);
}
const std::map<std::string, int> &distance_names() {
  static const std::map<std::string, int> m = {
      {"L1", 0}, {"L2", NGTGPU_DISTANCE_L2}, {"Hamming", NGTGPU_DISTANCE_HAMMING}, {"Angle", NGTGPU_DISTANCE_ANGLE},
      {"Cosine", NGTGPU_DISTANCE_COSINE}, {"NormalizedAngle", NGTGPU_DISTANCE_NORMALIZED_ANGLE},
      {"NormalizedCosine", NGTGPU_DISTANCE_NORMALIZED_COSINE}, {"Jaccard", 7}, {"NormalizedL2", NGTGPU_DISTANCE_NORMALIZED_L2}};
  return m;
}
std::string distance_type_name(int t) {
  for (auto &kv : distance_names())
    if (kv.second == t) return kv.first;
  return "None";
}
bool normalizes(int dt) {
  return dt == NGTGPU_DISTANCE_NORMALIZED_ANGLE || dt == NGTGPU_DISTANCE_NORMALIZED_COSINE || dt == NGTGPU_DISTANCE_NORMALIZED_L2;
}

std::map<std::string, std::string> default_prf() {   // what `ngt create` writes (Index.h:60-103, Graph.h:385-420)
  return {{"AccuracyTable", ""}, {"BatchSizeForCreation", "200"}, {"BuildTimeLimit", "0"}, {"DatabaseType", "Memory"},
          {"Dimension", "0"}, {"DistanceType", "L2"}, {"DynamicEdgeSizeBase", "30"}, {"DynamicEdgeSizeRate", "20"},
          {"EdgeSizeForCreation", "10"}, {"EdgeSizeForSearch", "40"}, {"EdgeSizeLimitForCreation", "5"},
          {"EpsilonForCreation", "0.1"}, {"GraphType", "ANNG"}, {"IncomingEdge", "80"},
          {"IncrimentalEdgeSizeLimitForTruncation", "0"}, {"IndexType", "Graph"}, {"ObjectAlignment", "False"},
          {"ObjectType", "Float-4"}, {"OutgoingEdge", "10"}, {"PathAdjustmentInterval", "0"}, {"PrefetchOffset", "0"},
          {"PrefetchSize", "0"}, {"SeedSize", "10"}, {"SeedType", "None"}, {"ThreadPoolSize", "24"},
          {"TruncationThreadPoolSize", "8"}};
}

void read_prf(CapiIndex &ix) {
  std::ifstream f(ix.path + "/prf");
  if (!f.is_open()) throw std::runtime_error("PropertySet::load: Cannot load the property file " + ix.path + "/prf.");
  std::string line;
  while (std::getline(f, line)) {
    size_t t = line.find('\t');
    if (t == std::string::npos) continue;
    ix.prf[line.substr(0, t)] = line.substr(t + 1);
  }
  ix.prop.dimension = std::stoi(ix.prf.at("Dimension"));
  ix.prop.object_type = object_type_of(ix.prf.at("ObjectType"));
  auto d = distance_names().find(ix.prf.at("DistanceType"));
  if (d == distance_names().end()) throw std::runtime_error("Invalid Distance Type in the property. " + ix.prf.at("DistanceType"));
  ix.prop.distance_type = d->second;
  if (ix.prf.count("EdgeSizeForCreation")) ix.prop.edge_size_for_creation = (int16_t)std::stoi(ix.prf["EdgeSizeForCreation"]);
  if (ix.prf.count("EdgeSizeForSearch")) ix.prop.edge_size_for_search = (int16_t)std::stoi(ix.prf["EdgeSizeForSearch"]);
}

void write_prf(CapiIndex &ix, const std::string &path) {
  ix.prf["Dimension"] = std::to_string(ix.prop.dimension);
  ix.prf["ObjectType"] = object_type_name(ix.prop.object_type);
  ix.prf["DistanceType"] = distance_type_name(ix.prop.distance_type);
  ix.prf["EdgeSizeForCreation"] = std::to_string(ix.prop.edge_size_for_creation);
  ix.prf["EdgeSizeForSearch"] = std::to_string(ix.prop.edge_size_for_search);
  ix.prf["IndexType"] = "Graph";   // no `tre` is written; the reference opens Graph indexes without one (Index.cpp:93-111)
  std::ofstream f(path + "/prf");
  if (!f.is_open()) throw std::runtime_error("PropertySet::save: Cannot save. " + path + "/prf");
  for (auto &kv : ix.prf) f << kv.first << "\t" << kv.second << "\n";
}

long prf_long(CapiIndex &ix, const char *key, long dflt) {
  auto it = ix.prf.find(key);
  if (it == ix.prf.end() || it->second.empty()) return dflt;
  return std::stol(it->second);
}

// expectedAccuracy -> epsilon through the `AccuracyTable` line of `prf` (ngtgpu_epsilon_from_accuracy_table below):
// GraphIndex::search replaces the query's epsilon by it when expectedAccuracy > 0 (Index.h:1156-1158).
float epsilon_from_expected_accuracy(CapiIndex &ix, double accuracy) {
  auto it = ix.prf.find("AccuracyTable");
  float e = 0.f;
  check(ngtgpu_epsilon_from_accuracy_table(it == ix.prf.end() ? "" : it->second.c_str(), accuracy, &e));
  return e;
}

// The device an index handle of the C API lives on: NGTGPU_DEVICE when set, else the calling thread's current CUDA device
// (a one-process-per-GPU host selects its GPU the usual way, with cudaSetDevice).
int default_device() {
  const char *env = getenv("NGTGPU_DEVICE");
  if (env && *env) return atoi(env);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return dev;
}

// Size of the device seed table that stands in for the DVP-tree (each query starts from its nearest SeedSize pivots).
// 256 measured best for batch throughput on 1M x 128 (one warp per query walks the table);
// NGTGPU_PIVOTS overrides.
size_t seed_table_pivots() {
  const char *env = getenv("NGTGPU_PIVOTS");
  long v = env ? atol(env) : 256;
  return (size_t)std::max<long>(1, v);
}

void upload(CapiIndex &ix) {
  const size_t n = ix.n();
  if (!ix.gpu) check(ngtgpu_index_create(&ix.gpu, default_device(), ix.prop.object_type, ix.prop.distance_type, (uint32_t)ix.prop.dimension));
  if (n == 0) return;
  check(ngtgpu_index_set_objects(ix.gpu, ix.objects.data(), n, 0, 0));
  std::vector<uint32_t> removed;
  for (size_t id = 1; id <= n; id++)
    if (!ix.present[id]) removed.push_back((uint32_t)id);
  if (!removed.empty()) check(ngtgpu_index_set_removed(ix.gpu, removed.data(), removed.size()));
  if (ix.row_ptr.size() == n + 2) {
    check(ngtgpu_index_set_graph(ix.gpu, ix.row_ptr.data(), ix.col.data(), 0));
    check(ngtgpu_index_set_search_property(ix.gpu, ix.prop.edge_size_for_search, prf_long(ix, "DynamicEdgeSizeBase", 30),
                                           prf_long(ix, "DynamicEdgeSizeRate", 20)));
  }
  uint32_t pivots = (uint32_t)std::min<size_t>(seed_table_pivots(), n - removed.size());
  if (pivots) check(ngtgpu_index_build_seed_table(ix.gpu, pivots, 1));
}

// NGTGPU_DEVICES="0,1,2,3": the devices an index is sharded over when it is opened (rows split evenly, a graph per shard)
std::vector<int> devices_from_env() {
  std::vector<int> d;
  const char *env = getenv("NGTGPU_DEVICES");
  if (!env) return d;
  std::stringstream ss(env);
  std::string tok;
  while (std::getline(ss, tok, ','))
    if (!tok.empty()) d.push_back(std::stoi(tok));
  return d;
}

// Rows sharded over `devices`: every shard gets its own graph from the reference's recipe on an exact neighbour table
// (ngtgpu_index_build_onng: -E EdgeSizeForCreation, then -o OutgoingEdge -i IncomingEdge + shortcut reduction when the
// index is an ONNG, the symmetric closure of the table -- what insertANNGNode converges to -- when it is an ANNG);
// the index's own `grp` spans shards and is kept only for ngt_get_edges.
void shard_index(CapiIndex &ix, const std::vector<int> &devices) {
  const size_t n = ix.n();
  for (size_t id = 1; id <= n; id++)
    if (!ix.present[id]) throw std::runtime_error("a sharded index cannot hold removed objects (id " + std::to_string(id) + ")");
  check(ngtgpu_sharded_create(&ix.sharded, devices.data(), (int)devices.size(), ix.prop.object_type, ix.prop.distance_type,
                              (uint32_t)ix.prop.dimension));
  check(ngtgpu_sharded_set_objects(ix.sharded, ix.objects.data(), n, 0));
  const uint32_t e = (uint32_t)std::max<int>(1, ix.prop.edge_size_for_creation);
  const bool onng = ix.prf.count("GraphType") && ix.prf["GraphType"] == "ONNG";
  const uint32_t out = onng ? (uint32_t)std::max<long>(0, prf_long(ix, "OutgoingEdge", 10)) : e;
  const uint32_t in = onng ? (uint32_t)std::max<long>(0, prf_long(ix, "IncomingEdge", 80)) : e;
  check(ngtgpu_sharded_build_onng(ix.sharded, e, out, in, onng ? 1 : 0, ix.prop.edge_size_for_search, 1024));
}

// NeighborhoodGraph::removeEdgesReliably (lib/NGT/Graph.cpp:641-864) on the host copy of the graph: the removed node's
// back edges go (each neighbour must hold one: the graph is an ANNG), then the neighbours are chained -- neighbour i is
// linked both ways to the nearest of the neighbours after it, which then takes place i + 1 -- so that they stay connected;
// the distances of that chain come from the device (the engine's exact distance). The graph covers ids 1..n_graph
// (= n when nothing is queued): an id appended since the last build has no node yet and is simply dropped.
void remove_edges_reliably(CapiIndex &ix, ObjectID id) {
  if (ix.row_ptr.size() < 2 || ix.row_ptr.size() > ix.n() + 2) return;
  const size_t n_graph = ix.row_ptr.size() - 2;
  if (id > n_graph) return;
  typedef std::pair<float, uint32_t> Edge;   // ordered like ObjectDistance: (distance, id)
  std::vector<Edge> node;
  for (uint64_t e = ix.row_ptr[id]; e < ix.row_ptr[id + 1]; e++)
    if (ix.col[e] != id) node.push_back(Edge(ix.dist[e], ix.col[e]));
  std::map<uint32_t, std::vector<Edge>> edited;   // the lists this removal touches
  auto list_of = [&](uint32_t nid) -> std::vector<Edge> & {
    auto it = edited.find(nid);
    if (it != edited.end()) return it->second;
    std::vector<Edge> &l = edited[nid];
    for (uint64_t e = ix.row_ptr[nid]; e < ix.row_ptr[nid + 1]; e++) l.push_back(Edge(ix.dist[e], ix.col[e]));
    return l;
  };
  for (const Edge &e : node) {
    std::vector<Edge> &n = list_of(e.second);
    auto pos = std::lower_bound(n.begin(), n.end(), Edge(e.first, id));
    // (no back edge -- e.g. an ONNG, whose lists are not symmetric: the reference is built with NGT_FORCED_REMOVE,
    // defines.h.in:36, reports it on stderr and goes on, Graph.cpp:727-733; so does this)
    if (pos != n.end() && pos->second == id) n.erase(pos);
  }
  const uint32_t m = (uint32_t)node.size();
  if (m > 1) {
    std::vector<uint32_t> order(m);
    for (uint32_t i = 0; i < m; i++) order[i] = node[i].second;
    std::vector<float> D((size_t)m * m);
    check(ngtgpu_index_pairwise_distances(ix.gpu, order.data(), m, D.data()));
    std::vector<uint32_t> slot(m);   // slot[i]: which row of D the object now at place i is
    for (uint32_t i = 0; i < m; i++) slot[i] = i;
    for (uint32_t i = 0; i + 1 < m; i++) {
      int minj = -1;
      float mind = FLT_MAX;
      for (uint32_t j = i + 1; j < m; j++) {
        const float d = D[(size_t)slot[i] * m + slot[j]];
        if (d < mind) {
          minj = (int)j;
          mind = d;
        }
      }
      if (minj < 0) throw std::runtime_error("removeEdgesReliably : Relink error ID=" + std::to_string(id));
      bool inserted[2];
      const uint32_t pair[2][2] = {{order[i], order[minj]}, {order[minj], order[i]}};
      for (int s2 = 0; s2 < 2; s2++) {
        std::vector<Edge> &n = list_of(pair[s2][0]);
        const Edge obj(mind, pair[s2][1]);
        auto pos = std::lower_bound(n.begin(), n.end(), obj);
        inserted[s2] = pos == n.end() || pos->second != obj.second;
        if (inserted[s2]) n.insert(pos, obj);
      }
      // (inserted[0] != inserted[1]: "Lost conectivity! Isn't this ANNG?" -- a warning under NGT_FORCED_REMOVE, Graph.cpp:805-813)
      if (i + 1 != (uint32_t)minj) {
        std::swap(order[i + 1], order[minj]);
        std::swap(slot[i + 1], slot[minj]);
      }
    }
  }
  edited[id].clear();
  // the CSR again, with the touched lists replaced
  std::vector<uint64_t> rp(n_graph + 2, 0);
  std::vector<uint32_t> col;
  std::vector<float> dist;
  col.reserve(ix.col.size() + 2 * m);
  dist.reserve(ix.col.size() + 2 * m);
  for (size_t s2 = 1; s2 <= n_graph; s2++) {
    rp[s2] = col.size();
    auto it = edited.find((uint32_t)s2);
    if (it != edited.end()) {
      for (const Edge &e : it->second) {
        col.push_back(e.second);
        dist.push_back(e.first);
      }
    } else {
      // (an asymmetric graph -- an ONNG -- can hold edges to the removed node in lists the node itself does not name;
      // the reference leaves those dangling, here they go too, so a search never walks into an empty slot)
      for (uint64_t e = ix.row_ptr[s2]; e < ix.row_ptr[s2 + 1]; e++)
        if (ix.col[e] != id) {
          col.push_back(ix.col[e]);
          dist.push_back(ix.dist[e]);
        }
    }
  }
  rp[n_graph + 1] = col.size();
  ix.row_ptr.swap(rp);
  ix.col.swap(col);
  ix.dist.swap(dist);
}

void require_built(CapiIndex &ix);

// ---- Optimizer::generateAccuracyTable (lib/NGT/Optimizer.h:1494-1573), the batched self-search consumer behind the
// `AccuracyTable` of an optimised index (GraphOptimizer::execute, GraphOptimizer.h:355-368): queries are objects taken
// at even strides from the repository (extractQueries, :1139-1190); a pseudo ground truth comes from raising epsilon
// until the answers stop changing (generatePseudoGroundTruth, :1418-1492) and one search with all edges at that epsilon;
// then accuracy(epsilon) -- relevant = id in the ground truth or distance within its farthest, :496-507 -- is sampled
// with the reference's adaptive step. Every search of the reference's loops is ONE batch call here. The reference ends
// the ground-truth loop when a sweep takes 40x the time of the first; here the distance computations of the sweep are
// the clock (same meaning, no timer in the result).
struct BatchAnswer {
  std::vector<uint32_t> ids, counts;
  std::vector<float> dists;
  uint64_t n_dist = 0;
};
void batch_answer(CapiIndex &ix, const std::vector<float> &queries, uint32_t nq, uint32_t size, float epsilon, int64_t edge_size,
                  BatchAnswer &out) {
  out.ids.assign((size_t)nq * size, 0);
  out.dists.assign((size_t)nq * size, 0.f);
  out.counts.assign(nq, 0);
  std::vector<uint32_t> stats((size_t)nq * 3, 0);
  ngtgpu_search_params p = {size, epsilon, -1.0f, edge_size};
  uint32_t seeds = (uint32_t)prf_long(ix, "SeedSize", 10);
  check(ngtgpu_search(ix.gpu, queries.data(), NGTGPU_OBJECT_FLOAT, nq, &p, nullptr, seeds ? seeds : 10, out.ids.data(), out.dists.data(),
                      out.counts.data(), stats.data()));
  out.n_dist = 0;
  for (uint32_t q = 0; q < nq; q++) out.n_dist += stats[(size_t)q * 3];
}

std::vector<std::pair<float, double>> generate_accuracy_table(CapiIndex &ix, size_t n_results, size_t n_queries) {
  if (ix.prop.edge_size_for_search != 0 && ix.prop.edge_size_for_search != -2) {
    std::stringstream msg;
    msg << "Optimizer::generateAccuracyTable: edgeSizeForSearch is invalid to call generateAccuracyTable, because accuracy 1.0 cannot "
           "be achieved with the setting. edgeSizeForSearch=" << ix.prop.edge_size_for_search << ".";
    throw std::runtime_error(msg.str());
  }
  require_built(ix);
  const size_t osize = ix.n() + 1, dim = (size_t)ix.prop.dimension;
  if (n_queries == 0 || osize / n_queries == 0) throw std::runtime_error("Optimizer::extractQueries: too few objects for the queries asked for");
  // extractQueries: the objects at even strides, the next present one where a slot is empty
  std::vector<float> queries;
  const size_t interval = osize / n_queries;
  size_t count = 0;
  for (size_t id1 = 1; id1 < osize && count < n_queries; id1 += interval, count++) {
    size_t oft = 0;
    while (!ix.present[id1 + oft]) {
      oft++;
      if (id1 + oft >= osize) {
        std::stringstream msg;
        msg << "Too many empty entries to extract. Object repository size=" << osize << " " << id1 << ":" << oft;
        throw std::runtime_error(msg.str());
      }
    }
    const uint8_t *row = &ix.objects[(id1 + oft - 1) * ix.record_bytes()];
    for (size_t j = 0; j < dim; j++)
      queries.push_back(ix.prop.object_type == NGTGPU_OBJECT_UINT8 ? (float)row[j] : reinterpret_cast<const float *>(row)[j]);
  }
  const uint32_t nq = (uint32_t)count, size = (uint32_t)n_results;
  BatchAnswer ans;
  // generatePseudoGroundTruth
  float max_epsilon = 0.0f;
  {
    int identity_count = 0;
    std::vector<float> last(nq, 0.f);
    uint64_t work0 = 0;
    double step = 0.02;
    for (float e = 0.0f; e < 10.0f; e += step) {
      batch_answer(ix, queries, nq, size, e, -1, ans);
      bool identity = true;
      for (uint32_t q = 0; q < nq; q++) {
        const float d = ans.counts[q] ? ans.dists[(size_t)q * size + ans.counts[q] - 1] : 0.f;
        if (d != last[q]) identity = false;
        last[q] = d;
      }
      if (e == 0.0f) work0 = ans.n_dist;
      if (ans.n_dist > work0 * 40) {
        max_epsilon = e;
        break;
      }
      if (identity) {
        identity_count++;
        step *= 1.2;
        if (identity_count > 5) {
          max_epsilon = e;
          break;
        }
      } else {
        identity_count = 0;
      }
    }
  }
  BatchAnswer gt;
  batch_answer(ix, queries, nq, size, max_epsilon, 0, gt);   // all edges: the best accuracy the graph gives
  auto accuracy_at = [&](float epsilon) {
    batch_answer(ix, queries, nq, size, epsilon, -1, ans);
    double sum = 0.0;
    for (uint32_t q = 0; q < nq; q++) {
      const uint32_t gn = gt.counts[q];
      if (gn == 0) continue;
      const uint32_t *gi = &gt.ids[(size_t)q * size];
      const float farthest = gt.dists[(size_t)q * size + gn - 1];
      uint32_t relevant = 0;
      for (uint32_t r = 0; r < ans.counts[q]; r++) {
        const uint32_t id = ans.ids[(size_t)q * size + r];
        const float d = ans.dists[(size_t)q * size + r];
        if (std::find(gi, gi + gn, id) != gi + gn) relevant++;
        else if (farthest > 0.0f && d <= farthest) relevant++;
      }
      sum += (double)relevant / (double)gn;
    }
    return sum / (double)nq;
  };
  std::map<float, double> map;
  {
    float interval2 = 0.05f, prev = 0.0f, epsilon = -0.6f;
    double accuracy;
    do {
      auto pair = map.find(epsilon);
      if (pair == map.end()) {
        accuracy = accuracy_at(epsilon);
        map.insert(std::make_pair(epsilon, accuracy));
      } else {
        accuracy = pair->second;
      }
      if (prev != 0.0f) {
        if (accuracy - prev < 0.02) {
          interval2 *= 2.0f;
        } else if (accuracy - prev > 0.05 && interval2 > 0.0001f) {
          epsilon -= interval2;
          interval2 /= 2.0f;
          accuracy = prev;
        }
      }
      prev = (float)accuracy;
      epsilon += interval2;
      if (accuracy > 0.98 && epsilon > max_epsilon) break;
    } while (accuracy < 1.0);
  }
  std::vector<std::pair<float, double>> table;
  std::pair<float, double> prev(0.0f, -1.0);
  for (auto &kv : map) {
    if (fabs(kv.first - prev.first) <= FLT_EPSILON) continue;
    if (kv.second - prev.second < DBL_EPSILON) continue;
    table.push_back(kv);
    if (kv.second >= 1.0) break;
    prev = kv;
  }
  return table;
}

std::string accuracy_table_string(const std::vector<std::pair<float, double>> &t) {   // Index::AccuracyTable::getString, Index.h:349-358
  std::stringstream str;
  for (size_t i = 0; i < t.size(); i++) {
    str << t[i].first << ":" << t[i].second;
    if (i + 1 != t.size()) str << ",";
  }
  return str.str();
}

void require_single(CapiIndex &ix, const char *what) {
  if (ix.sharded) throw std::runtime_error(std::string(what) + ": the index is sharded over several GPUs and read-only");
}

CapiIndex *open_index(const char *path, const std::vector<int> &devices = devices_from_env()) {
  std::unique_ptr<CapiIndex> ix(new CapiIndex);
  ix->path = path;
  read_prf(*ix);
  uint64_t slots = 0, pres = 0, gslots = 0, nnz = 0;
  const std::string obj = ix->path + "/obj", grp = ix->path + "/grp";
  check(ngtgpu_io_obj_info(obj.c_str(), (uint32_t)ix->record_bytes(), &slots, &pres));
  const size_t n = slots ? slots - 1 : 0;
  ix->objects.assign(n * ix->record_bytes(), 0);
  ix->present.assign(n + 1, 0);
  check(ngtgpu_io_read_obj(obj.c_str(), (uint32_t)ix->record_bytes(), ix->objects.data(), ix->present.data()));
  check(ngtgpu_io_grp_info(grp.c_str(), &gslots, &nnz));
  std::vector<uint64_t> rp(gslots + 1, 0);
  ix->col.assign(nnz ? nnz : 1, 0);
  ix->dist.assign(nnz ? nnz : 1, 0.f);
  check(ngtgpu_io_read_grp(grp.c_str(), rp.data(), ix->col.data(), ix->dist.data(), nullptr));
  ix->col.resize(nnz);
  ix->dist.resize(nnz);
  ix->row_ptr.assign(n + 2, nnz);
  for (size_t i = 0; i < rp.size() && i < n + 2; i++) ix->row_ptr[i] = rp[i];
  if (devices.size() > 1 && n >= devices.size()) {
    try {
      shard_index(*ix, devices);
    } catch (...) {
      if (ix->sharded) ngtgpu_sharded_destroy(ix->sharded);
      throw;
    }
  } else {
    upload(*ix);
  }
  return ix.release();
}

void require_built(CapiIndex &ix) {
  if (ix.sharded) return;
  if (ix.pending) throw std::runtime_error("objects were appended: call ngt_create_index() before searching");
  if (!ix.gpu || ix.n() == 0) throw std::runtime_error("the index holds no objects");
}

void fill_results(NGTObjectDistances results, const uint32_t *ids, const float *dists, uint32_t count) {
  auto &r = *static_cast<std::vector<NGTObjectDistance> *>(results);
  r.clear();   // the container is cleared and overwritten per call (Graph.cpp:631-635, ObjectSpace.h:49-57)
  for (uint32_t i = 0; i < count; i++) r.push_back(NGTObjectDistance{ids[i], dists[i]});
}

void search_one(CapiIndex &ix, const float *q, int32_t dim, size_t size, float epsilon, float radius, int64_t edge_size,
                NGTObjectDistances results) {
  if (dim != ix.prop.dimension) throw std::runtime_error("ObjectSpace::allocateObject: the specified dimension is invalid");
  require_built(ix);
  if (!ix.sharded && ix.row_ptr.size() != ix.n() + 2) throw std::runtime_error("the index has no graph: call ngt_create_index()");
  ngtgpu_search_params p = {(uint32_t)size, epsilon, radius >= FLT_MAX ? -1.0f : radius, edge_size};
  std::vector<uint32_t> ids(size ? size : 1), cnt(1, 0);
  std::vector<float> ds(size ? size : 1);
  uint32_t seeds = (uint32_t)prf_long(ix, "SeedSize", 10);
  if (seeds == 0) seeds = 10;
  if (ix.sharded) check(ngtgpu_sharded_search(ix.sharded, q, NGTGPU_OBJECT_FLOAT, 1, &p, seeds, ids.data(), ds.data(), cnt.data()));
  else {
    uint32_t stats[3] = {0, 0, 0};   // {distance computations, adjacency entries, expansions}: Graph.cpp:592,604
    check(ngtgpu_search(ix.gpu, q, NGTGPU_OBJECT_FLOAT, 1, &p, nullptr, seeds, ids.data(), ds.data(), cnt.data(), stats));
    __atomic_fetch_add(&ix.num_dist, (uint64_t)stats[0], __ATOMIC_RELAXED);   // searches may run concurrently on one index
  }
  fill_results(results, ids.data(), ds.data(), cnt[0]);
}

void linear_one(CapiIndex &ix, const float *q, int32_t dim, size_t size, float radius, NGTObjectDistances results) {
  if (dim != ix.prop.dimension) throw std::runtime_error("ObjectSpace::allocateObject: the specified dimension is invalid");
  if ((!ix.gpu && !ix.sharded) || ix.n() == 0) throw std::runtime_error("the index holds no objects");
  std::vector<uint32_t> ids(size ? size : 1), cnt(1, 0);
  std::vector<float> ds(size ? size : 1);
  if (ix.sharded)
    check(ngtgpu_sharded_linear_search(ix.sharded, q, NGTGPU_OBJECT_FLOAT, 1, (uint32_t)size, radius >= FLT_MAX ? -1.0f : radius,
                                       ids.data(), ds.data(), cnt.data()));
  else
    check(ngtgpu_linear_search(ix.gpu, q, NGTGPU_OBJECT_FLOAT, 1, (uint32_t)size, radius >= FLT_MAX ? -1.0f : radius, ids.data(),
                               ds.data(), cnt.data()));
  fill_results(results, ids.data(), ds.data(), cnt[0]);
}

ObjectID append_rows(CapiIndex &ix, const float *rows, size_t count, uint32_t dim) {
  require_single(ix, "insert");
  if ((int32_t)dim != ix.prop.dimension) throw std::runtime_error("ObjectSpace::allocateObject: the specified dimension is invalid");
  const size_t rb = ix.record_bytes();
  const size_t first = ix.n() + 1;
  if (ix.present.empty()) ix.present.push_back(0);
  const size_t old = ix.objects.size();
  ix.objects.resize(old + count * rb);
  for (size_t r = 0; r < count; r++) {
    if (ix.prop.object_type == NGTGPU_OBJECT_UINT8) {
      for (uint32_t j = 0; j < dim; j++) ix.objects[old + r * rb + j] = (uint8_t)rows[r * dim + j];   // ObjectRepository.h:222-258
    } else {
      if (normalizes(ix.prop.distance_type)) {
        bool zero = true;
        for (uint32_t j = 0; j < dim; j++) zero = zero && rows[r * dim + j] == 0.0f;
        if (zero) throw std::runtime_error("ObjectSpace::normalize: Error! the object is an invalid zero vector for the cosine similarity or normalized distances.");
      }
      memcpy(&ix.objects[old + r * rb], rows + r * dim, rb);
    }
    ix.present.push_back(1);
  }
  if (normalizes(ix.prop.distance_type) && ix.raw_from == 0) ix.raw_from = first;
  ix.pending += count;
  return (ObjectID)first;
}

// NGT::Index::createIndex on the device: exact kNN of every object (ngtgpu_index_knn_graph), then the ANNG is the
// symmetric closure of those lists (out-edges + reverse edges, sorted by (distance,id), duplicates dropped) -- what
// insertANNGNode converges to (lib/NGT/Graph.h:611-626).
void build_graph(CapiIndex &ix) {
  const size_t n = ix.n();
  if (n == 0 || ix.sharded) return;   // a sharded index is built when it is opened
  // createIndex only indexes objects that are not in the graph yet (Index.cpp:645-648): with nothing queued and a
  // graph in place (loaded ONNG, refined or optimised graph) it is a no-op.
  if (ix.pending == 0 && ix.row_ptr.size() == n + 2 && !ix.col.empty()) return;
  if (!ix.gpu) check(ngtgpu_index_create(&ix.gpu, default_device(), ix.prop.object_type, ix.prop.distance_type, (uint32_t)ix.prop.dimension));
  const size_t rb = ix.record_bytes();
  if (ix.raw_from) {
    // ObjectSpace::normalize (ObjectSpace.h:251-266) on the device for the newly appended rows
    ngtgpu_index *tmp = nullptr;
    check(ngtgpu_index_create(&tmp, default_device(), ix.prop.object_type, ix.prop.distance_type, (uint32_t)ix.prop.dimension));
    int rc = ngtgpu_index_set_objects(tmp, &ix.objects[(ix.raw_from - 1) * rb], n - ix.raw_from + 1, 1, 0);
    if (rc == NGTGPU_OK) rc = ngtgpu_index_get_objects(tmp, 1, n - ix.raw_from + 1, &ix.objects[(ix.raw_from - 1) * rb]);
    ngtgpu_index_destroy(tmp);
    check(rc);
    ix.raw_from = 0;
  }
  check(ngtgpu_index_set_objects(ix.gpu, ix.objects.data(), n, 0, 0));
  std::vector<uint32_t> removed;
  for (size_t id = 1; id <= n; id++)
    if (!ix.present[id]) removed.push_back((uint32_t)id);
  if (!removed.empty()) check(ngtgpu_index_set_removed(ix.gpu, removed.data(), removed.size()));
  size_t live = n - removed.size();
  // Objects appended to an index that already has its graph are INSERTED into it the way the reference's
  // construction loop does (batches searched on the frozen graph, in-batch distances, reverse edges; Index.cpp:721-792),
  // not by rebuilding everything.
  if (ix.pending > 0 && ix.pending < n && ix.row_ptr.size() == n - ix.pending + 2 && !ix.col.empty()) {
    const size_t n_old = n - ix.pending;
    const uint32_t e = (uint32_t)std::max<int>(1, ix.prop.edge_size_for_creation);
    const uint64_t cap = ix.col.size() + 2ull * ix.pending * e + 16;
    uint64_t *d_rp = nullptr, nnz = 0;
    uint32_t *d_col = nullptr;
    float *d_dist = nullptr;
    if (cudaMalloc(&d_rp, (n + 2) * 8) != cudaSuccess || cudaMalloc(&d_col, cap * 4) != cudaSuccess ||
        cudaMalloc(&d_dist, cap * 4) != cudaSuccess) {
      cudaFree(d_rp);
      cudaFree(d_col);
      cudaFree(d_dist);
      throw std::runtime_error("cudaMalloc failed while inserting into the graph");
    }
    std::vector<uint64_t> rp(n + 2, ix.col.size());
    for (size_t i = 0; i < n_old + 2; i++) rp[i] = ix.row_ptr[i];
    cudaMemcpy(d_rp, rp.data(), (n + 2) * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_col, ix.col.data(), ix.col.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_dist, ix.dist.data(), ix.dist.size() * 4, cudaMemcpyHostToDevice);
    int rc = NGTGPU_OK;
    const size_t batch = (size_t)std::max<long>(1, prf_long(ix, "BatchSizeForCreation", 200));
    const float eps = ix.prf.count("EpsilonForCreation") ? std::stof(ix.prf["EpsilonForCreation"]) : 0.1f;
    check(ngtgpu_index_set_search_property(ix.gpu, ix.prop.edge_size_for_search, prf_long(ix, "DynamicEdgeSizeBase", 30),
                                           prf_long(ix, "DynamicEdgeSizeRate", 20)));
    for (size_t s = n_old + 1; s <= n && rc == NGTGPU_OK; s += batch) {
      const uint32_t m = (uint32_t)std::min<size_t>(batch, n - s + 1);
      rc = ngtgpu_index_insert_batch(ix.gpu, (uint32_t)s, m, e, eps, -1, 10, 1024, 1, cap, d_rp, d_col, d_dist, &nnz);
    }
    if (rc == NGTGPU_OK) {
      ix.row_ptr.assign(n + 2, 0);
      ix.col.assign(nnz, 0);
      ix.dist.assign(nnz, 0.f);
      cudaMemcpy(ix.row_ptr.data(), d_rp, (n + 2) * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(ix.col.data(), d_col, nnz * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(ix.dist.data(), d_dist, nnz * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_rp);
    cudaFree(d_col);
    cudaFree(d_dist);
    check(rc);
    ix.pending = 0;
    check(ngtgpu_index_build_seed_table(ix.gpu, (uint32_t)std::min<size_t>(seed_table_pivots(), live), 1));
    return;
  }
  uint32_t k = (uint32_t)std::max<long>(1, std::min<long>(ix.prop.edge_size_for_creation, (long)live - 1));
  uint32_t *d_ids = nullptr, *d_cnt = nullptr;
  float *d_d = nullptr;
  if (cudaMalloc(&d_ids, n * k * 4) != cudaSuccess || cudaMalloc(&d_d, n * k * 4) != cudaSuccess ||
      cudaMalloc(&d_cnt, n * 4) != cudaSuccess)
    throw std::runtime_error("cudaMalloc failed while building the graph");
  int rc = NGTGPU_OK;
  for (size_t s = 0; s < n && rc == NGTGPU_OK; s += 131072) {
    uint32_t m = (uint32_t)std::min<size_t>(131072, n - s);
    rc = ngtgpu_index_knn_graph(ix.gpu, k, (uint32_t)s + 1, m, d_ids + s * k, d_d + s * k, d_cnt + s, nullptr);
  }
  // the ANNG: out-edges + reverse edges, sorted by (distance, id), repeated ids dropped -- on the device
  uint64_t *d_rp = nullptr, nnz = 0;
  uint32_t *d_col = nullptr;
  float *d_dist = nullptr;
  uint8_t *d_valid = nullptr;
  const uint64_t cap = (uint64_t)n * k * 2;
  if (rc == NGTGPU_OK &&
      (cudaMalloc(&d_rp, (n + 2) * 8) != cudaSuccess || cudaMalloc(&d_col, cap * 4) != cudaSuccess ||
       cudaMalloc(&d_dist, cap * 4) != cudaSuccess || cudaMalloc(&d_valid, n + 1) != cudaSuccess))
    rc = NGTGPU_ERR_CUDA;
  if (rc == NGTGPU_OK) {
    cudaDeviceSynchronize();
    cudaMemcpy(d_valid, ix.present.data(), n + 1, cudaMemcpyHostToDevice);
    rc = ngtgpu_graph_from_knn_table(n, d_ids, d_d, d_cnt, k, d_valid, 1, cap, d_rp, d_col, d_dist, &nnz, nullptr);
  }
  if (rc == NGTGPU_OK) {
    ix.row_ptr.assign(n + 2, 0);
    ix.col.assign(nnz, 0);
    ix.dist.assign(nnz, 0.f);
    cudaMemcpy(ix.row_ptr.data(), d_rp, (n + 2) * 8, cudaMemcpyDeviceToHost);
    if (nnz) {
      cudaMemcpy(ix.col.data(), d_col, nnz * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(ix.dist.data(), d_dist, nnz * 4, cudaMemcpyDeviceToHost);
    }
  }
  cudaFree(d_ids);
  cudaFree(d_d);
  cudaFree(d_cnt);
  cudaFree(d_rp);
  cudaFree(d_col);
  cudaFree(d_dist);
  cudaFree(d_valid);
  check(rc);
  ix.pending = 0;
  ix.prf["GraphType"] = "ANNG";
  check(ngtgpu_index_set_graph(ix.gpu, ix.row_ptr.data(), ix.col.data(), 0));
  check(ngtgpu_index_set_search_property(ix.gpu, ix.prop.edge_size_for_search, prf_long(ix, "DynamicEdgeSizeBase", 30),
                                         prf_long(ix, "DynamicEdgeSizeRate", 20)));
  check(ngtgpu_index_build_seed_table(ix.gpu, (uint32_t)std::min<size_t>(seed_table_pivots(), live), 1));
}

}  // namespace

extern "C" {

// ---- error objects, Capi.cpp:784-812 --------------------------------------------------------------------
NGTError ngt_create_error_object() {
  try {
    return static_cast<NGTError>(new std::string());
  } catch (std::exception &err) {
    std::cerr << "Capi : " << __FUNCTION__ << "() : Error: " << err.what();
    return NULL;
  }
}
const char *ngt_get_error_string(const NGTError error) { return static_cast<std::string *>(error)->c_str(); }
void ngt_clear_error_string(NGTError error) { *static_cast<std::string *>(error) = ""; }
void ngt_destroy_error_object(NGTError error) { delete static_cast<std::string *>(error); }

// ---- properties, Capi.cpp:113-325 -----------------------------------------------------------------------
NGTProperty ngt_create_property(NGTError error) {
  try {
    return static_cast<NGTProperty>(new CapiProperty());
  }
  CAPI_CATCH(NULL)
}
void ngt_destroy_property(NGTProperty prop) {
  if (prop) delete static_cast<CapiProperty *>(prop);
}
#define PROP_CHECK(ret)                                                                                      \
  if (prop == NULL) {                                                                                        \
    std::stringstream ss;                                                                                    \
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: prop = " << prop;                              \
    operate_error_string(ss, error);                                                                         \
    return ret;                                                                                              \
  }
bool ngt_get_property(const NGTIndex index, NGTProperty prop, NGTError error) {
  if (index == NULL || prop == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " prop = " << prop;
    operate_error_string(ss, error);
    return false;
  }
  *static_cast<CapiProperty *>(prop) = static_cast<CapiIndex *>(index)->prop;
  return true;
}
int32_t ngt_get_property_dimension(NGTProperty prop, NGTError error) { PROP_CHECK(-1) return static_cast<CapiProperty *>(prop)->dimension; }
bool ngt_set_property_dimension(NGTProperty prop, int32_t value, NGTError error) { PROP_CHECK(false) static_cast<CapiProperty *>(prop)->dimension = value; return true; }
bool ngt_set_property_edge_size_for_creation(NGTProperty prop, int16_t value, NGTError error) { PROP_CHECK(false) static_cast<CapiProperty *>(prop)->edge_size_for_creation = value; return true; }
bool ngt_set_property_edge_size_for_search(NGTProperty prop, int16_t value, NGTError error) { PROP_CHECK(false) static_cast<CapiProperty *>(prop)->edge_size_for_search = value; return true; }
int16_t ngt_get_property_edge_size_for_creation(NGTProperty prop, NGTError error) { PROP_CHECK(-1) return static_cast<CapiProperty *>(prop)->edge_size_for_creation; }
int16_t ngt_get_property_edge_size_for_search(NGTProperty prop, NGTError error) { PROP_CHECK(-1) return static_cast<CapiProperty *>(prop)->edge_size_for_search; }
int32_t ngt_get_property_object_type(NGTProperty prop, NGTError error) { PROP_CHECK(-1) return static_cast<CapiProperty *>(prop)->object_type; }
int32_t ngt_get_property_distance_type(NGTProperty prop, NGTError error) { PROP_CHECK(-1) return static_cast<CapiProperty *>(prop)->distance_type; }
bool ngt_is_property_object_type_float(int32_t t) { return t == NGTGPU_OBJECT_FLOAT; }
bool ngt_is_property_object_type_integer(int32_t t) { return t == NGTGPU_OBJECT_UINT8; }
bool ngt_set_property_object_type_float(NGTProperty prop, NGTError error) { PROP_CHECK(false) static_cast<CapiProperty *>(prop)->object_type = NGTGPU_OBJECT_FLOAT; return true; }
bool ngt_set_property_object_type_integer(NGTProperty prop, NGTError error) { PROP_CHECK(false) static_cast<CapiProperty *>(prop)->object_type = NGTGPU_OBJECT_UINT8; return true; }
#define SET_DISTANCE(fn, value)                                                                              \
  bool fn(NGTProperty prop, NGTError error) { PROP_CHECK(false) static_cast<CapiProperty *>(prop)->distance_type = value; return true; }
SET_DISTANCE(ngt_set_property_distance_type_l1, 0)
SET_DISTANCE(ngt_set_property_distance_type_l2, NGTGPU_DISTANCE_L2)
SET_DISTANCE(ngt_set_property_distance_type_angle, NGTGPU_DISTANCE_ANGLE)
SET_DISTANCE(ngt_set_property_distance_type_hamming, NGTGPU_DISTANCE_HAMMING)
SET_DISTANCE(ngt_set_property_distance_type_jaccard, 7)
SET_DISTANCE(ngt_set_property_distance_type_cosine, NGTGPU_DISTANCE_COSINE)
SET_DISTANCE(ngt_set_property_distance_type_normalized_angle, NGTGPU_DISTANCE_NORMALIZED_ANGLE)
SET_DISTANCE(ngt_set_property_distance_type_normalized_cosine, NGTGPU_DISTANCE_NORMALIZED_COSINE)
// additive (include/ngt_capi_ext.h): Capi.h:86-104 has no setter for DistanceTypeNormalizedL2, ngtpy.create takes it
SET_DISTANCE(ngt_set_property_distance_type_normalized_l2, NGTGPU_DISTANCE_NORMALIZED_L2)

// ---- index life cycle, Capi.cpp:40-111, 694-711 ------------------------------------------------------------
NGTIndex ngt_open_index(const char *index_path, NGTError error) {
  try {
    return static_cast<NGTIndex>(open_index(index_path));
  }
  CAPI_CATCH(NULL)
}
NGTIndex ngt_open_index_as_read_only(const char *index_path, NGTError error) { return ngt_open_index(index_path, error); }
// additive: the index opened with its rows sharded over `devices` (what NGTGPU_DEVICES does for ngt_open_index)
NGTIndex ngt_open_index_sharded(const char *index_path, const int *devices, int n_devices, NGTError error) {
  try {
    if (!devices || n_devices < 1) throw std::runtime_error("no devices given");
    return static_cast<NGTIndex>(open_index(index_path, std::vector<int>(devices, devices + n_devices)));
  }
  CAPI_CATCH(NULL)
}

static NGTIndex create_empty(const char *database, NGTProperty prop, NGTError error, const char *fn) {
  try {
    if (prop == NULL) throw std::runtime_error("property is NULL");
    std::unique_ptr<CapiIndex> ix(new CapiIndex);
    ix->prop = *static_cast<CapiProperty *>(prop);
    ix->prf = default_prf();
    ix->present.push_back(0);
    ix->row_ptr.assign(2, 0);
    if (database) {
      ix->path = database;
      mkdir(database, 0755);
      write_prf(*ix, ix->path);
      check(ngtgpu_io_write_obj((ix->path + "/obj").c_str(), (uint32_t)ix->record_bytes(), nullptr, 0, nullptr));
      check(ngtgpu_io_write_grp((ix->path + "/grp").c_str(), 0, ix->row_ptr.data(), nullptr, nullptr, nullptr));
    }
    check(ngtgpu_index_create(&ix->gpu, default_device(), ix->prop.object_type, ix->prop.distance_type, (uint32_t)ix->prop.dimension));
    return static_cast<NGTIndex>(ix.release());
  } catch (std::exception &err) {
    std::stringstream ss;
    ss << "Capi : " << fn << "() : Error: " << err.what();
    operate_error_string(ss, error);
    return NULL;
  }
}
NGTIndex ngt_create_graph_and_tree(const char *database, NGTProperty prop, NGTError error) {
  return create_empty(database, prop, error, __FUNCTION__);
}
NGTIndex ngt_create_graph_and_tree_in_memory(NGTProperty prop, NGTError error) {
  return create_empty(nullptr, prop, error, __FUNCTION__);
}

bool ngt_save_index(const NGTIndex index, const char *database, NGTError error) {
  if (index == NULL || database == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " database = " << database;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    if (ix.pending) throw std::runtime_error("objects were appended: call ngt_create_index() before ngt_save_index()");
    mkdir(database, 0755);
    write_prf(ix, database);
    const std::string d(database);
    check(ngtgpu_io_write_obj((d + "/obj").c_str(), (uint32_t)ix.record_bytes(), ix.objects.data(), ix.n(), ix.present.data()));
    check(ngtgpu_io_write_grp((d + "/grp").c_str(), ix.n(), ix.row_ptr.data(), ix.col.data(), ix.dist.data(), ix.present.data()));
    // `prf` now says IndexType Graph: a DVP-tree file left by the reference in this directory would describe another
    // object set, so it goes (the reference opens Graph indexes without one, Index.cpp:93-111)
    ::unlink((d + "/tre").c_str());
  }
  CAPI_CATCH(false)
  return true;
}

void ngt_close_index(NGTIndex index) {
  if (index == NULL) return;
  CapiIndex *ix = static_cast<CapiIndex *>(index);
  if (ix->gpu) ngtgpu_index_destroy(ix->gpu);
  if (ix->sharded) ngtgpu_sharded_destroy(ix->sharded);
  delete ix;
}

// ---- results, Capi.cpp:541-578 ---------------------------------------------------------------------------
NGTObjectDistances ngt_create_empty_results(NGTError error) {
  try {
    return static_cast<NGTObjectDistances>(new std::vector<NGTObjectDistance>());
  }
  CAPI_CATCH(NULL)
}
void ngt_destroy_results(NGTObjectDistances results) {
  if (results) delete static_cast<std::vector<NGTObjectDistance> *>(results);
}
uint32_t ngt_get_result_size(NGTObjectDistances results, NGTError error) {
  if (results == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: results = " << results;
    operate_error_string(ss, error);
    return 0;
  }
  return (uint32_t) static_cast<std::vector<NGTObjectDistance> *>(results)->size();
}
int32_t ngt_get_size(NGTObjectDistances results, NGTError error) {   // deprecated twin
  if (results == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: results = " << results;
    operate_error_string(ss, error);
    return -1;
  }
  return (int32_t) static_cast<std::vector<NGTObjectDistance> *>(results)->size();
}
NGTObjectDistance ngt_get_result(const NGTObjectDistances results, const uint32_t i, NGTError error) {
  try {
    return static_cast<std::vector<NGTObjectDistance> *>(results)->at(i);
  } catch (std::exception &err) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : Error: " << err.what();
    operate_error_string(ss, error);
    NGTObjectDistance e = {0, 0};
    return e;
  }
}

// ---- search, Capi.cpp:327-539 ------------------------------------------------------------------------------
#define SEARCH_ARGCHECK(qptr)                                                                               \
  if (index == NULL || qptr == NULL || results == NULL || query_dim <= 0) {                                 \
    std::stringstream ss;                                                                                   \
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " query = " << (const void *)qptr \
       << " results = " << results << " query_dim = " << query_dim;                                         \
    operate_error_string(ss, error);                                                                        \
    return false;                                                                                           \
  }

bool ngt_search_index(NGTIndex index, double *query, int32_t query_dim, size_t size, float epsilon, float radius,
                      NGTObjectDistances results, NGTError error) {
  SEARCH_ARGCHECK(query)
  try {
    if (radius < 0.0) radius = FLT_MAX;   // Capi.cpp:384-386
    std::vector<float> q(query, query + query_dim);
    search_one(*static_cast<CapiIndex *>(index), q.data(), query_dim, size, epsilon, radius, -1, results);
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_search_index_as_float(NGTIndex index, float *query, int32_t query_dim, size_t size, float epsilon, float radius,
                               NGTObjectDistances results, NGTError error) {
  SEARCH_ARGCHECK(query)
  try {
    if (radius < 0.0) radius = FLT_MAX;
    search_one(*static_cast<CapiIndex *>(index), query, query_dim, size, epsilon, radius, -1, results);
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_search_index_with_query(NGTIndex index, NGTQuery query, NGTObjectDistances results, NGTError error) {
  if (index == NULL || query.query == NULL || results == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " query = " << (const void *)query.query
       << " results = " << results;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    float radius = query.radius < 0.0 ? FLT_MAX : query.radius;
    int64_t es = query.edge_size == (size_t)INT_MIN ? -1 : (int64_t)(int)query.edge_size;
    // Capi.cpp:346-375 passes query.accuracy as the expected accuracy; Index.h:1156-1158 turns it into epsilon
    const float eps = query.accuracy > 0.0f ? epsilon_from_expected_accuracy(ix, query.accuracy) : query.epsilon;
    search_one(ix, query.query, ix.prop.dimension, query.size, eps, radius, es, results);
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_linear_search_index(NGTIndex index, double *query, int32_t query_dim, size_t size, NGTObjectDistances results,
                             NGTError error) {
  SEARCH_ARGCHECK(query)
  try {
    std::vector<float> q(query, query + query_dim);
    linear_one(*static_cast<CapiIndex *>(index), q.data(), query_dim, size, FLT_MAX, results);
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_linear_search_index_as_float(NGTIndex index, float *query, int32_t query_dim, size_t size, NGTObjectDistances results,
                                      NGTError error) {
  SEARCH_ARGCHECK(query)
  try {
    linear_one(*static_cast<CapiIndex *>(index), query, query_dim, size, FLT_MAX, results);
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_linear_search_index_with_query(NGTIndex index, NGTQuery query, NGTObjectDistances results, NGTError error) {
  if (index == NULL || query.query == NULL || results == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " query = " << (const void *)query.query
       << " results = " << results;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    linear_one(ix, query.query, ix.prop.dimension, query.size, query.radius < 0.0 ? FLT_MAX : query.radius, results);
  }
  CAPI_CATCH(false)
  return true;
}

// ---- additive batch entry points: nq queries at once, flat outputs -------------
bool ngt_batch_search_index_as_float(NGTIndex index, const float *queries, uint32_t nq, int32_t query_dim, size_t size,
                                     float epsilon, float radius, int64_t edge_size, uint32_t *ids, float *dists,
                                     uint32_t *counts, NGTError error) {
  if (index == NULL || queries == NULL || ids == NULL || dists == NULL || counts == NULL || query_dim <= 0) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " queries = " << (const void *)queries;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    if (query_dim != ix.prop.dimension) throw std::runtime_error("ObjectSpace::allocateObject: the specified dimension is invalid");
    require_built(ix);
    ngtgpu_search_params p = {(uint32_t)size, epsilon, radius, edge_size};
    uint32_t seeds = (uint32_t)prf_long(ix, "SeedSize", 10);
    if (ix.sharded) check(ngtgpu_sharded_search(ix.sharded, queries, NGTGPU_OBJECT_FLOAT, nq, &p, seeds ? seeds : 10, ids, dists, counts));
    else check(ngtgpu_search(ix.gpu, queries, NGTGPU_OBJECT_FLOAT, nq, &p, nullptr, seeds ? seeds : 10, ids, dists, counts, nullptr));
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_batch_linear_search_index_as_float(NGTIndex index, const float *queries, uint32_t nq, int32_t query_dim, size_t size,
                                            float radius, uint32_t *ids, float *dists, uint32_t *counts, NGTError error) {
  if (index == NULL || queries == NULL || ids == NULL || dists == NULL || counts == NULL || query_dim <= 0) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " queries = " << (const void *)queries;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    if (query_dim != ix.prop.dimension) throw std::runtime_error("ObjectSpace::allocateObject: the specified dimension is invalid");
    if ((!ix.gpu && !ix.sharded) || ix.n() == 0) throw std::runtime_error("the index holds no objects");
    if (ix.sharded) check(ngtgpu_sharded_linear_search(ix.sharded, queries, NGTGPU_OBJECT_FLOAT, nq, (uint32_t)size, radius, ids, dists, counts));
    else check(ngtgpu_linear_search(ix.gpu, queries, NGTGPU_OBJECT_FLOAT, nq, (uint32_t)size, radius, ids, dists, counts));
  }
  CAPI_CATCH(false)
  return true;
}

// uint8 twins: the queries are already bytes, as the objects of an Integer-1 index are stored
bool ngt_batch_search_index_as_uint8(NGTIndex index, const uint8_t *queries, uint32_t nq, int32_t query_dim, size_t size,
                                     float epsilon, float radius, int64_t edge_size, uint32_t *ids, float *dists,
                                     uint32_t *counts, NGTError error) {
  if (index == NULL || queries == NULL || ids == NULL || dists == NULL || counts == NULL || query_dim <= 0) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " queries = " << (const void *)queries;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    if (query_dim != ix.prop.dimension) throw std::runtime_error("ObjectSpace::allocateObject: the specified dimension is invalid");
    if (ix.prop.object_type != NGTGPU_OBJECT_UINT8) throw std::runtime_error("the object type of the index is not integer (uint8)");
    require_built(ix);
    ngtgpu_search_params p = {(uint32_t)size, epsilon, radius, edge_size};
    uint32_t seeds = (uint32_t)prf_long(ix, "SeedSize", 10);
    if (ix.sharded) check(ngtgpu_sharded_search(ix.sharded, queries, NGTGPU_OBJECT_UINT8, nq, &p, seeds ? seeds : 10, ids, dists, counts));
    else check(ngtgpu_search(ix.gpu, queries, NGTGPU_OBJECT_UINT8, nq, &p, nullptr, seeds ? seeds : 10, ids, dists, counts, nullptr));
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_batch_linear_search_index_as_uint8(NGTIndex index, const uint8_t *queries, uint32_t nq, int32_t query_dim, size_t size,
                                            float radius, uint32_t *ids, float *dists, uint32_t *counts, NGTError error) {
  if (index == NULL || queries == NULL || ids == NULL || dists == NULL || counts == NULL || query_dim <= 0) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " queries = " << (const void *)queries;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    if (query_dim != ix.prop.dimension) throw std::runtime_error("ObjectSpace::allocateObject: the specified dimension is invalid");
    if (ix.prop.object_type != NGTGPU_OBJECT_UINT8) throw std::runtime_error("the object type of the index is not integer (uint8)");
    if ((!ix.gpu && !ix.sharded) || ix.n() == 0) throw std::runtime_error("the index holds no objects");
    if (ix.sharded) check(ngtgpu_sharded_linear_search(ix.sharded, queries, NGTGPU_OBJECT_UINT8, nq, (uint32_t)size, radius, ids, dists, counts));
    else check(ngtgpu_linear_search(ix.gpu, queries, NGTGPU_OBJECT_UINT8, nq, (uint32_t)size, radius, ids, dists, counts));
  }
  CAPI_CATCH(false)
  return true;
}

// ---- insertion and construction, Capi.cpp:580-711 ---------------------------------------------------------
ObjectID ngt_append_index_as_float(NGTIndex index, float *obj, uint32_t obj_dim, NGTError error) {
  if (index == NULL || obj == NULL || obj_dim == 0) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " obj = " << (void *)obj << " obj_dim = " << obj_dim;
    operate_error_string(ss, error);
    return 0;
  }
  try {
    return append_rows(*static_cast<CapiIndex *>(index), obj, 1, obj_dim);
  }
  CAPI_CATCH(0)
}
ObjectID ngt_insert_index_as_float(NGTIndex index, float *obj, uint32_t obj_dim, NGTError error) {
  return ngt_append_index_as_float(index, obj, obj_dim, error);
}
ObjectID ngt_append_index(NGTIndex index, double *obj, uint32_t obj_dim, NGTError error) {
  if (index == NULL || obj == NULL || obj_dim == 0) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " obj = " << (void *)obj << " obj_dim = " << obj_dim;
    operate_error_string(ss, error);
    return 0;
  }
  try {
    std::vector<float> v(obj, obj + obj_dim);
    return append_rows(*static_cast<CapiIndex *>(index), v.data(), 1, obj_dim);
  }
  CAPI_CATCH(0)
}
ObjectID ngt_insert_index(NGTIndex index, double *obj, uint32_t obj_dim, NGTError error) {
  return ngt_append_index(index, obj, obj_dim, error);
}
bool ngt_batch_append_index(NGTIndex index, float *obj, uint32_t data_count, NGTError error) {
  if (index == NULL || obj == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " obj = " << (void *)obj;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    append_rows(ix, obj, data_count, (uint32_t)ix.prop.dimension);
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_batch_insert_index(NGTIndex index, float *obj, uint32_t data_count, uint32_t *ids, NGTError error) {
  if (index == NULL || obj == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " obj = " << (void *)obj;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    ObjectID first = append_rows(ix, obj, data_count, (uint32_t)ix.prop.dimension);
    if (ids)
      for (uint32_t i = 0; i < data_count; i++) ids[i] = first + i;
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_create_index(NGTIndex index, uint32_t pool_size, NGTError error) {
  if (index == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: idnex = " << index;
    operate_error_string(ss, error);
    return false;
  }
  try {
    (void)pool_size;   // the thread pool of the reference (Index.cpp:737-741) is replaced by device batching
    build_graph(*static_cast<CapiIndex *>(index));
  }
  CAPI_CATCH(false)
  return true;
}
bool ngt_remove_index(NGTIndex index, ObjectID id, NGTError error) {
  if (index == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    require_single(ix, "remove");
    if (id == 0 || id > ix.n() || !ix.present[id]) throw std::runtime_error("remove: the specified object does not exist. ID=" + std::to_string(id));
    remove_edges_reliably(ix, id);
    ix.present[id] = 0;
    upload(ix);
  }
  CAPI_CATCH(false)
  return true;
}

// ---- objects and edges, Capi.cpp:713-782, 1006-1040 ----------------------------------------------------------
NGTObjectSpace ngt_get_object_space(NGTIndex index, NGTError error) {
  if (index == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: idnex = " << index;
    operate_error_string(ss, error);
    return NULL;
  }
  return index;   // the object space of this engine lives in the index handle
}
float *ngt_get_object_as_float(NGTObjectSpace object_space, ObjectID id, NGTError error) {
  try {
    if (object_space == NULL) throw std::runtime_error("object space is NULL");
    CapiIndex &ix = *static_cast<CapiIndex *>(object_space);
    if (ix.prop.object_type != NGTGPU_OBJECT_FLOAT) throw std::runtime_error("the object type is not float");
    if (id == 0 || id > ix.n() || !ix.present[id]) throw std::runtime_error("ObjectSpace::getObject: the object does not exist. ID=" + std::to_string(id));
    return reinterpret_cast<float *>(&ix.objects[(size_t)(id - 1) * ix.record_bytes()]);   // points into the index, as Capi.cpp:750-765
  }
  CAPI_CATCH(NULL)
}
uint8_t *ngt_get_object_as_integer(NGTObjectSpace object_space, ObjectID id, NGTError error) {
  try {
    if (object_space == NULL) throw std::runtime_error("object space is NULL");
    CapiIndex &ix = *static_cast<CapiIndex *>(object_space);
    if (ix.prop.object_type != NGTGPU_OBJECT_UINT8) throw std::runtime_error("the object type is not integer");
    if (id == 0 || id > ix.n() || !ix.present[id]) throw std::runtime_error("ObjectSpace::getObject: the object does not exist. ID=" + std::to_string(id));
    return &ix.objects[(size_t)(id - 1) * ix.record_bytes()];
  }
  CAPI_CATCH(NULL)
}
bool ngt_get_edges(NGTIndex index, ObjectID id, NGTObjectDistances edges, NGTError error) {
  if (index == NULL || edges == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index << " edges = " << edges;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    if (id == 0 || id > ix.n() || ix.row_ptr.size() != ix.n() + 2) throw std::runtime_error("the specified node does not exist. ID=" + std::to_string(id));
    auto &r = *static_cast<std::vector<NGTObjectDistance> *>(edges);
    r.clear();
    for (uint64_t e = ix.row_ptr[id]; e < ix.row_ptr[id + 1]; e++) r.push_back(NGTObjectDistance{ix.col[e], ix.dist[e]});
  }
  CAPI_CATCH(false)
  return true;
}
uint32_t ngt_get_object_repository_size(NGTIndex index, NGTError error) {
  if (index == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index;
    operate_error_string(ss, error);
    return 0;
  }
  return (uint32_t) static_cast<CapiIndex *>(index)->present.size();
}

// ---- refineANNG and the ONNG recipe (GraphReconstructor / GraphOptimizer) on the device ------------------------
}  // extern "C"

namespace {

struct DeviceGraph {   // a CSR with distances in device buffers of `cap` entries
  size_t n = 0;
  uint64_t *rp = nullptr;
  uint32_t *col = nullptr;
  float *dist = nullptr;
  uint64_t cap = 0, nnz = 0;
  DeviceGraph(size_t n_, uint64_t capacity) : n(n_), cap(std::max<uint64_t>(capacity, 1)) {
    if (cudaMalloc(&rp, (n + 2) * 8) != cudaSuccess || cudaMalloc(&col, cap * 4) != cudaSuccess ||
        cudaMalloc(&dist, cap * 4) != cudaSuccess) {
      release();
      throw std::runtime_error("cudaMalloc failed (graph buffers)");
    }
    cudaMemset(rp, 0, (n + 2) * 8);
    cudaStreamSynchronize(cudaStreamLegacy);   // (the fill runs on the legacy stream; the users' streams are not ordered with it)
  }
  DeviceGraph(const DeviceGraph &) = delete;
  ~DeviceGraph() { release(); }
  void release() {
    cudaFree(rp);
    cudaFree(col);
    cudaFree(dist);
    rp = nullptr, col = nullptr, dist = nullptr;
  }
  void from_host(const CapiIndex &ix) {
    nnz = ix.col.size();
    cudaMemcpy(rp, ix.row_ptr.data(), (n + 2) * 8, cudaMemcpyHostToDevice);
    if (nnz) {
      cudaMemcpy(col, ix.col.data(), nnz * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dist, ix.dist.data(), nnz * 4, cudaMemcpyHostToDevice);
    }
  }
  void to_host(CapiIndex &ix) const {
    ix.row_ptr.assign(n + 2, 0);
    ix.col.assign(nnz, 0);
    ix.dist.assign(nnz, 0.f);
    cudaMemcpy(ix.row_ptr.data(), rp, (n + 2) * 8, cudaMemcpyDeviceToHost);
    if (nnz) {
      cudaMemcpy(ix.col.data(), col, nnz * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(ix.dist.data(), dist, nnz * 4, cudaMemcpyDeviceToHost);
    }
  }
};

struct CapiOptimizer {   // the settings of NGT::GraphOptimizer the C API reaches (GraphOptimizer.h:60-78, 600-650)
  bool log_disabled = false;
  int outgoing = 10, incoming = 120, queries = 100, results = 20;
  size_t min_edges = 0;
  bool shortcut_reduction = true;
  bool search_parameter = true, prefetch_parameter = true, accuracy_table = true;
};

// GraphOptimizer::execute (GraphOptimizer.h:230-300): copy the index, reconstructGraph, path adjustment, save the graph
// and the property. The search-parameter tuning that follows in the reference (optimizeSearchParameters, :302-350:
// repeated timed searches on the host) is outside the hot path and is not run: `prf` keeps its search parameters.
void optimizer_execute(const CapiOptimizer &o, const std::string &in, const std::string &out) {
  if (access(out.c_str(), 0) == 0) throw std::runtime_error("Optimizer::execute: The specified index exists. " + out);
  {   // GraphOptimizer.h:244-248 shells out to `cp -r`; same effect without a shell
    std::error_code ec;
    std::filesystem::copy(in, out, std::filesystem::copy_options::recursive, ec);
    if (ec) throw std::runtime_error("Optimizer::execute: Cannot create the specified index. " + out);
  }
  CapiIndex ix;
  ix.path = out;
  read_prf(ix);
  uint64_t gslots = 0, nnz = 0;
  const std::string grp = out + "/grp";
  check(ngtgpu_io_grp_info(grp.c_str(), &gslots, &nnz));
  if (gslots < 2) throw std::runtime_error("Optimizer::execute: the index holds no graph");
  const size_t n = gslots - 1;
  std::vector<uint64_t> rp(gslots + 1, 0);
  ix.present.assign(n + 1, 0);
  ix.col.assign(nnz ? nnz : 1, 0);
  ix.dist.assign(nnz ? nnz : 1, 0.f);
  check(ngtgpu_io_read_grp(grp.c_str(), rp.data(), ix.col.data(), ix.dist.data(), ix.present.data()));
  ix.col.resize(nnz);
  ix.dist.resize(nnz);
  ix.row_ptr.assign(n + 2, nnz);
  for (size_t i = 0; i < rp.size() && i < n + 2; i++) ix.row_ptr[i] = rp[i];
  std::unique_ptr<DeviceGraph> g(new DeviceGraph(n, nnz));
  g->from_host(ix);
  if (o.outgoing > 0 || o.incoming > 0) {
    if (ix.prf["GraphType"] != "ANNG") {   // convertToANNG (GraphReconstructor.h:389-423): add the reverse of every edge
      std::unique_ptr<DeviceGraph> h(new DeviceGraph(n, 2 * g->nnz));
      check(ngtgpu_graph_reconstruct(n, g->rp, g->col, g->dist, 0xffffffffu, 0xffffffffu, h->cap, h->rp, h->col, h->dist, &h->nnz,
                                     nullptr));
      g.swap(h);
    }
    std::unique_ptr<DeviceGraph> h(new DeviceGraph(n, 2 * g->nnz));
    check(ngtgpu_graph_reconstruct(n, g->rp, g->col, g->dist, (uint32_t)std::max(o.outgoing, 0), (uint32_t)std::max(o.incoming, 0),
                                   h->cap, h->rp, h->col, h->dist, &h->nnz, nullptr));
    g.swap(h);
    ix.prf["GraphType"] = "ONNG";
  }
  if (o.shortcut_reduction) {
    uint8_t *keep = nullptr;
    if (cudaMalloc(&keep, std::max<uint64_t>(g->nnz, 1)) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (edge mask)");
    int rc = ngtgpu_graph_adjust_paths(n, g->rp, g->col, g->dist, (uint32_t)o.min_edges, keep, nullptr, nullptr);
    std::unique_ptr<DeviceGraph> h;
    if (rc == NGTGPU_OK) {
      h.reset(new DeviceGraph(n, g->nnz));
      rc = ngtgpu_graph_select_edges(n, g->rp, g->col, g->dist, keep, h->rp, h->col, h->dist, &h->nnz, nullptr);
    }
    cudaFree(keep);
    check(rc);
    g.swap(h);
  }
  g->to_host(ix);
  check(ngtgpu_io_write_grp(grp.c_str(), n, ix.row_ptr.data(), ix.col.data(), ix.dist.data(), ix.present.data()));
  // (IndexType and the rest of the property stay as they were: only GraphType changes, GraphOptimizer.h:272-274)
  auto save_prf = [&]() {
    std::ofstream f(out + "/prf");
    if (!f.is_open()) throw std::runtime_error("PropertySet::save: Cannot save. " + out + "/prf");
    for (auto &kv : ix.prf) f << kv.first << "\t" << kv.second << "\n";
  };
  save_prf();
  // The accuracy table (GraphOptimizer.h:355-368): generated on the device when the index's edge-size mode allows an
  // accuracy of 1.0 (0 or -2, Optimizer.h:1497-1502). The reference reaches -2 through its timed coefficient tuning, which
  // is not run here; an index still on a fixed edge cap keeps its (empty) table instead of failing the whole call.
  if (o.accuracy_table && (ix.prop.edge_size_for_search == 0 || ix.prop.edge_size_for_search == -2)) {
    std::unique_ptr<CapiIndex, void (*)(CapiIndex *)> opened(open_index(out.c_str(), std::vector<int>()), [](CapiIndex *p) {
      if (p->gpu) ngtgpu_index_destroy(p->gpu);
      delete p;
    });
    try {
      ix.prf["AccuracyTable"] = accuracy_table_string(generate_accuracy_table(*opened, (size_t)o.results, (size_t)o.queries));
    } catch (std::exception &err) {
      throw std::runtime_error(std::string("Optimizer::execute: Cannot generate the accuracy table. ") + err.what());
    }
    save_prf();
  }
}

}  // namespace

extern "C" {

// GraphReconstructor::refineANNG behind Capi.cpp:976-1004. expectedAccuracy > 0 is turned into epsilon through the
// index's accuracy table (Index.h:1156-1158); an index without one fails with the reference's message.
bool ngt_refine_anng(NGTIndex index, float epsilon, float expectedAccuracy, int noOfEdges, int edgeSize, size_t batchSize,
                     NGTError error) {
  if (index == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index;
    operate_error_string(ss, error);
    return false;
  }
  try {
    CapiIndex &ix = *static_cast<CapiIndex *>(index);
    require_single(ix, "refineANNG");
    require_built(ix);
    if (expectedAccuracy > 0.0f) epsilon = epsilon_from_expected_accuracy(ix, expectedAccuracy);   // GraphReconstructor.h:843-847
    if (ix.row_ptr.size() != ix.n() + 2) throw std::runtime_error("refineANNG: the index holds no graph");
    const size_t n = ix.n();
    const int ec = ix.prop.edge_size_for_creation;
    const uint32_t k = (uint32_t)(noOfEdges < 0 ? -noOfEdges : std::max(noOfEdges, ec));   // GraphReconstructor.h:825
    if (k == 0) throw std::runtime_error("refineANNG: no edges to search for");
    if (batchSize == 0) batchSize = 10000;
    DeviceGraph g(n, ix.col.size() + 2ull * n * k);
    g.from_host(ix);
    const int64_t es = edgeSize == INT_MIN ? -1 : (int64_t)edgeSize;
    check(ngtgpu_index_refine_anng(ix.gpu, epsilon, noOfEdges, es, k, batchSize, 10, g.cap, g.rp, g.col, g.dist, &g.nnz));
    g.to_host(ix);
    check(ngtgpu_index_set_search_property(ix.gpu, ix.prop.edge_size_for_search, prf_long(ix, "DynamicEdgeSizeBase", 30),
                                           prf_long(ix, "DynamicEdgeSizeRate", 20)));
  }
  CAPI_CATCH(false)
  return true;
}

NGTOptimizer ngt_create_optimizer(bool logDisabled, NGTError error) {
  try {
    CapiOptimizer *o = new CapiOptimizer;
    o->log_disabled = logDisabled;
    return static_cast<NGTOptimizer>(o);
  }
  CAPI_CATCH(NULL)
}

#define OPT_CHECK(ret)                                                                        \
  if (optimizer == NULL) {                                                                    \
    std::stringstream ss;                                                                     \
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: optimizer = " << optimizer;     \
    operate_error_string(ss, error);                                                          \
    return ret;                                                                               \
  }

// Search-coefficient tuning (GraphOptimizer::adjustSearchCoefficients: timed host searches) is outside the hot path.
bool ngt_optimizer_adjust_search_coefficients(NGTOptimizer optimizer, const char *, NGTError error) {
  OPT_CHECK(false)
  std::stringstream ss;
  ss << "Capi : " << __FUNCTION__ << "() : Error: not provided by the B200 engine (search-parameter tuning is outside the hot path)";
  operate_error_string(ss, error);
  return false;
}

bool ngt_optimizer_execute(NGTOptimizer optimizer, const char *inIndex, const char *outIndex, NGTError error) {
  OPT_CHECK(false)
  try {
    if (inIndex == NULL || outIndex == NULL) throw std::runtime_error("Optimizer::execute: null index path");
    optimizer_execute(*static_cast<CapiOptimizer *>(optimizer), inIndex, outIndex);
  }
  CAPI_CATCH(false)
  return true;
}

// Capi.cpp:907-975: negative values leave a setting as it is (GraphOptimizer::set, GraphOptimizer.h:600-628)
bool ngt_optimizer_set(NGTOptimizer optimizer, int outgoing, int incoming, int nofqs, float, float, float, float, double,
                       double, NGTError error) {
  OPT_CHECK(false)
  CapiOptimizer &o = *static_cast<CapiOptimizer *>(optimizer);
  if (outgoing >= 0) o.outgoing = outgoing;
  if (incoming >= 0) o.incoming = incoming;
  if (nofqs > 0) o.queries = nofqs;
  return true;
}
bool ngt_optimizer_set_minimum(NGTOptimizer optimizer, int outgoing, int incoming, int nofqs, int nofrs, NGTError error) {
  OPT_CHECK(false)
  CapiOptimizer &o = *static_cast<CapiOptimizer *>(optimizer);
  if (outgoing >= 0) o.outgoing = outgoing;
  if (incoming >= 0) o.incoming = incoming;
  if (nofqs > 0) o.queries = nofqs;
  if (nofrs > 0) o.results = nofrs;
  return true;
}
bool ngt_optimizer_set_extension(NGTOptimizer optimizer, float, float, float, float, double, double, NGTError error) {
  OPT_CHECK(false)
  return true;   // accuracy ranges of the search-parameter tuning: kept for the signature, the tuning is not run here
}
bool ngt_optimizer_set_processing_modes(NGTOptimizer optimizer, bool searchParameter, bool prefetchParameter,
                                        bool accuracyTable, NGTError error) {
  OPT_CHECK(false)
  CapiOptimizer &o = *static_cast<CapiOptimizer *>(optimizer);
  o.search_parameter = searchParameter;
  o.prefetch_parameter = prefetchParameter;
  o.accuracy_table = accuracyTable;
  return true;
}
void ngt_destroy_optimizer(NGTOptimizer optimizer) { delete static_cast<CapiOptimizer *>(optimizer); }

// ---- additive, for the ngtpy module (csrc/ngtpy.cpp; declared in include/ngt_capi_ext.h) -----------------------
// GraphOptimizer::shortcutReduction (GraphOptimizer.h:73, set through setProcessingModes :640-650): Capi.h's
// ngt_optimizer_set_processing_modes carries the three tuning flags only, ngtpy's set_processing_modes has this one too.
bool ngt_optimizer_set_shortcut_reduction(NGTOptimizer optimizer, bool shortcutReduction, NGTError error) {
  OPT_CHECK(false)
  static_cast<CapiOptimizer *>(optimizer)->shortcut_reduction = shortcutReduction;
  return true;
}
// Sum of SearchContainer::distanceComputationCount (Graph.cpp:592) over the single-query searches this handle served:
// what ngtpy.Index.get_num_of_distance_computations accumulates (python/src/ngtpy.cpp:181,349).
uint64_t ngt_get_number_of_distance_computations(NGTIndex index, NGTError error) {
  if (index == NULL) {
    std::stringstream ss;
    ss << "Capi : " << __FUNCTION__ << "() : parametor error: index = " << index;
    operate_error_string(ss, error);
    return 0;
  }
  return __atomic_load_n(&static_cast<CapiIndex *>(index)->num_dist, __ATOMIC_RELAXED);
}

// Capi.cpp:1043-1058: the defaults of GraphOptimizer::ANNGEdgeOptimizationParameter (GraphOptimizer.h:28-36)
NGTAnngEdgeOptimizationParameter ngt_get_anng_edge_optimization_parameter() {
  NGTAnngEdgeOptimizationParameter p;
  p.no_of_queries = 200;
  p.no_of_results = 50;
  p.no_of_threads = 16;
  p.target_accuracy = 0.9f;
  p.target_no_of_objects = 0;
  p.no_of_sample_objects = 100000;
  p.max_of_no_of_edges = 100;
  p.log = false;
  return p;
}

// Capi.cpp:1060-1088 -> GraphOptimizer::optimizeNumberOfEdgesForANNG (GraphOptimizer.h:387-533): a tuner that rebuilds
// the index on growing samples and extrapolates the edge count from timed host searches. Like the other search-parameter
// tuners it is outside the hot path; the symbol exists so programs written against Capi.h link, and it fails loudly.
bool ngt_optimize_number_of_edges(const char *indexPath, NGTAnngEdgeOptimizationParameter parameter, NGTError error) {
  (void)indexPath;
  (void)parameter;
  std::stringstream ss;
  ss << "Capi : " << __FUNCTION__ << "() : Error: not provided by the B200 engine (edge-number tuning is outside the hot path)";
  operate_error_string(ss, error);
  return false;
}

}  // extern "C"

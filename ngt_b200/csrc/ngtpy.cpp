// ngtpy.cpp -- the `ngtpy` Python module (python/src/ngtpy.cpp:500-639 of the reference) over libngtgpu.so.
//
// The reference's module subclasses NGT::Index in C++; this one holds an NGTIndex handle and drives the engine through
// the C ABI only (the `ngt_*` functions of lib/NGT/Capi.h that csrc/capi.cu serves, plus include/ngt_capi_ext.h): same
// module name, classes, method names, keyword arguments, defaults, id numbering and return shapes, so a script written
// for ngtpy runs unchanged (`import ngtpy` with ngt_b200/ on sys.path, or `from ngt_b200 import ngtpy`).
// Additive: Index.batch_search / Index.batch_linear_search -- a GPU wants batches.
// Not provided (outside the hot path, fail loudly): QuantizedIndex (NGTQG), export_index / import_index (text dump),
// the timed tuners of Optimizer. No distance is computed on the host anywhere in this file.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cfloat>
#include <climits>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define NGT_CAPI_EXT_NO_TYPEDEFS
namespace py = pybind11;

extern "C" {
// lib/NGT/Capi.h:28-47
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
// the functions of lib/NGT/Capi.h:60-212 this module calls
NGTIndex ngt_open_index(const char *, NGTError);
NGTIndex ngt_open_index_as_read_only(const char *, NGTError);
NGTIndex ngt_create_graph_and_tree(const char *, NGTProperty, NGTError);
NGTProperty ngt_create_property(NGTError);
void ngt_destroy_property(NGTProperty);
bool ngt_get_property(const NGTIndex, NGTProperty, NGTError);
int32_t ngt_get_property_dimension(NGTProperty, NGTError);
int32_t ngt_get_property_object_type(NGTProperty, NGTError);
bool ngt_is_property_object_type_integer(int32_t);
bool ngt_set_property_dimension(NGTProperty, int32_t, NGTError);
bool ngt_set_property_edge_size_for_creation(NGTProperty, int16_t, NGTError);
bool ngt_set_property_edge_size_for_search(NGTProperty, int16_t, NGTError);
bool ngt_set_property_object_type_float(NGTProperty, NGTError);
bool ngt_set_property_object_type_integer(NGTProperty, NGTError);
bool ngt_set_property_distance_type_l1(NGTProperty, NGTError);
bool ngt_set_property_distance_type_l2(NGTProperty, NGTError);
bool ngt_set_property_distance_type_angle(NGTProperty, NGTError);
bool ngt_set_property_distance_type_hamming(NGTProperty, NGTError);
bool ngt_set_property_distance_type_jaccard(NGTProperty, NGTError);
bool ngt_set_property_distance_type_cosine(NGTProperty, NGTError);
bool ngt_set_property_distance_type_normalized_angle(NGTProperty, NGTError);
bool ngt_set_property_distance_type_normalized_cosine(NGTProperty, NGTError);
bool ngt_save_index(const NGTIndex, const char *, NGTError);
void ngt_close_index(NGTIndex);
NGTObjectDistances ngt_create_empty_results(NGTError);
void ngt_destroy_results(NGTObjectDistances);
uint32_t ngt_get_result_size(NGTObjectDistances, NGTError);
NGTObjectDistance ngt_get_result(const NGTObjectDistances, const uint32_t, NGTError);
bool ngt_search_index_with_query(NGTIndex, NGTQuery, NGTObjectDistances, NGTError);
bool ngt_linear_search_index_with_query(NGTIndex, NGTQuery, NGTObjectDistances, NGTError);
ObjectID ngt_insert_index(NGTIndex, double *, uint32_t, NGTError);
bool ngt_batch_append_index(NGTIndex, float *, uint32_t, NGTError);
bool ngt_create_index(NGTIndex, uint32_t, NGTError);
bool ngt_remove_index(NGTIndex, ObjectID, NGTError);
NGTObjectSpace ngt_get_object_space(NGTIndex, NGTError);
float *ngt_get_object_as_float(NGTObjectSpace, ObjectID, NGTError);
uint8_t *ngt_get_object_as_integer(NGTObjectSpace, ObjectID, NGTError);
NGTError ngt_create_error_object();
const char *ngt_get_error_string(const NGTError);
void ngt_destroy_error_object(NGTError);
bool ngt_refine_anng(NGTIndex, float, float, int, int, size_t, NGTError);
NGTOptimizer ngt_create_optimizer(bool, NGTError);
bool ngt_optimizer_adjust_search_coefficients(NGTOptimizer, const char *, NGTError);
bool ngt_optimizer_execute(NGTOptimizer, const char *, const char *, NGTError);
bool ngt_optimizer_set(NGTOptimizer, int, int, int, float, float, float, float, double, double, NGTError);
bool ngt_optimizer_set_processing_modes(NGTOptimizer, bool, bool, bool, NGTError);
void ngt_destroy_optimizer(NGTOptimizer);
NGTAnngEdgeOptimizationParameter ngt_get_anng_edge_optimization_parameter();
bool ngt_optimize_number_of_edges(const char *, NGTAnngEdgeOptimizationParameter, NGTError);
int ngtgpu_epsilon_from_accuracy_table(const char *, double, float *);   // include/ngtgpu.h
const char *ngtgpu_last_error();
}
#include "../../include/ngt_capi_ext.h"

namespace {

// One NGTError per call; a failed call becomes a RuntimeError carrying the library's message, as an NGT::Exception
// thrown through the reference's module does (pybind11 translates std::exception).
struct Err {
  NGTError e;
  Err() : e(ngt_create_error_object()) {}
  ~Err() { ngt_destroy_error_object(e); }
  std::string text() const {
    std::string s = ngt_get_error_string(e);
    const std::string mark = " : Error: ";   // "Capi : <func>() : Error: <what>" -> <what>
    size_t p = s.find(mark);
    return p == std::string::npos ? s : s.substr(p + mark.size());
  }
  [[noreturn]] void raise() const { throw std::runtime_error(text()); }
};

struct Results {
  NGTObjectDistances r;
  explicit Results(Err &err) : r(ngt_create_empty_results(err.e)) {
    if (!r) err.raise();
  }
  ~Results() { ngt_destroy_results(r); }
};

class Index {
public:
  Index(const std::string &path, bool readOnly, bool zeroBasedNumbering, bool treeDisabled, bool logDisabled)
      : path_(path), zeroNumbering(zeroBasedNumbering) {
    (void)treeDisabled;   // seeds come from the device seed table either way
    (void)logDisabled;
    Err err;
    index_ = readOnly ? ngt_open_index_as_read_only(path.c_str(), err.e) : ngt_open_index(path.c_str(), err.e);
    if (!index_) err.raise();
    NGTProperty prop = ngt_create_property(err.e);
    if (!prop) err.raise();
    bool ok = ngt_get_property(index_, prop, err.e);
    if (ok) {
      dimension_ = ngt_get_property_dimension(prop, err.e);
      integer_ = ngt_is_property_object_type_integer(ngt_get_property_object_type(prop, err.e));
    }
    ngt_destroy_property(prop);
    if (!ok) {
      ngt_close_index(index_);
      index_ = nullptr;
      err.raise();
    }
    // python/src/ngtpy.cpp:43-48
    defaultNumOfSearchObjects = 20;
    defaultEpsilon = 0.1f;
    defaultRadius = FLT_MAX;
    defaultEdgeSize = -1;
    defaultExpectedAccuracy = -1.0f;
  }
  ~Index() { close(); }
  Index(const Index &) = delete;
  Index &operator=(const Index &) = delete;

  static void create(const std::string &path, size_t dimension, int edgeSizeForCreation, int edgeSizeForSearch,
                     const std::string &distanceType, const std::string &objectType) {
    Err err;
    NGTProperty prop = ngt_create_property(err.e);
    if (!prop) err.raise();
    struct Guard {
      NGTProperty p;
      ~Guard() { ngt_destroy_property(p); }
    } guard{prop};
    bool ok = ngt_set_property_dimension(prop, (int32_t)dimension, err.e) &&
              ngt_set_property_edge_size_for_creation(prop, (int16_t)edgeSizeForCreation, err.e) &&
              ngt_set_property_edge_size_for_search(prop, (int16_t)edgeSizeForSearch, err.e);
    if (!ok) err.raise();
    // python/src/ngtpy.cpp:59-99
    if (objectType == "Float" || objectType == "float") ok = ngt_set_property_object_type_float(prop, err.e);
    else if (objectType == "Byte" || objectType == "byte") ok = ngt_set_property_object_type_integer(prop, err.e);
    else throw std::runtime_error("ngtpy::create: invalid object type. " + objectType);
    if (!ok) err.raise();
    if (distanceType == "L1") ok = ngt_set_property_distance_type_l1(prop, err.e);
    else if (distanceType == "L2") ok = ngt_set_property_distance_type_l2(prop, err.e);
    else if (distanceType == "Hamming") ok = ngt_set_property_distance_type_hamming(prop, err.e);
    else if (distanceType == "Jaccard") ok = ngt_set_property_distance_type_jaccard(prop, err.e);
    else if (distanceType == "Angle") ok = ngt_set_property_distance_type_angle(prop, err.e);
    else if (distanceType == "Normalized Angle") ok = ngt_set_property_distance_type_normalized_angle(prop, err.e);
    else if (distanceType == "Cosine") ok = ngt_set_property_distance_type_cosine(prop, err.e);
    else if (distanceType == "Normalized Cosine") ok = ngt_set_property_distance_type_normalized_cosine(prop, err.e);
    else if (distanceType == "Normalized L2") ok = ngt_set_property_distance_type_normalized_l2(prop, err.e);
    else throw std::runtime_error("ngtpy::create: invalid distance type. " + distanceType);
    if (!ok) err.raise();
    NGTIndex ix = ngt_create_graph_and_tree(path.c_str(), prop, err.e);   // NGT::Index::createGraphAndTree
    if (!ix) err.raise();
    ngt_close_index(ix);
  }

  void batchInsert(py::array_t<double, py::array::c_style | py::array::forcecast> objects, size_t numThreads, bool debug) {
    py::buffer_info info = objects.request();
    if (info.ndim != 2) throw std::runtime_error("ngtpy::insert: Error! a two-dimensional array is expected.");
    if (debug) std::cerr << info.ndim << ":" << info.shape[0] << ":" << info.shape[1] << std::endl;
    if ((py::ssize_t)dimension_ != info.shape[1]) {
      std::stringstream msg;   // python/src/ngtpy.cpp:117-121
      msg << "ngtpy::insert: Error! dimensions are inconsitency. " << dimension_ << ":" << info.shape[1];
      throw std::runtime_error(msg.str());
    }
    const double *src = static_cast<const double *>(info.ptr);
    std::vector<float> rows(src, src + (size_t)info.shape[0] * (size_t)info.shape[1]);
    const uint32_t count = (uint32_t)info.shape[0];
    Err err;
    bool ok;
    {
      py::gil_scoped_release nogil;
      ok = ngt_batch_append_index(handle(), rows.data(), count, err.e) && ngt_create_index(handle(), (uint32_t)numThreads, err.e);
    }
    if (!ok) err.raise();
    distBase_ = counter();
  }

  int insert(py::array_t<double, py::array::c_style | py::array::forcecast> object, bool debug) {
    py::buffer_info info = object.request();
    double *ptr = static_cast<double *>(info.ptr);
    if (debug) {
      for (py::ssize_t i = 0; i < info.size; i++) std::cerr << ptr[i] << " ";
      std::cerr << std::endl;
    }
    Err err;
    ObjectID id = ngt_insert_index(handle(), ptr, (uint32_t)info.size, err.e);
    if (id == 0) err.raise();
    distBase_ = counter();
    return zeroNumbering ? (int)id - 1 : (int)id;
  }

  void buildIndex(size_t numThreads, size_t targetSizeOfGraph) {
    (void)targetSizeOfGraph;
    Err err;
    bool ok;
    {
      py::gil_scoped_release nogil;
      ok = ngt_create_index(handle(), (uint32_t)numThreads, err.e);
    }
    if (!ok) err.raise();
  }

  py::object search(py::object query, size_t size, float epsilon, int edgeSize, float expectedAccuracy, bool withDistance) {
    py::array_t<float, py::array::c_style | py::array::forcecast> qobject(query);
    py::buffer_info qinfo = qobject.request();
    if ((size_t)qinfo.size != (size_t)dimension_) return dimensionError(withDistance);
    NGTQuery q;
    q.query = static_cast<float *>(qinfo.ptr);
    q.size = size == 0 ? defaultNumOfSearchObjects : size;
    q.radius = defaultRadius >= FLT_MAX ? -1.0f : defaultRadius;
    q.accuracy = expectedAccuracy > 0.0f ? expectedAccuracy : -1.0f;
    q.epsilon = epsilon <= -1.0f ? defaultEpsilon : epsilon;
    q.edge_size = (size_t)(int64_t)(edgeSize < -2 ? defaultEdgeSize : edgeSize);
    Err err;
    Results res(err);
    bool ok;
    {
      py::gil_scoped_release nogil;
      ok = ngt_search_index_with_query(handle(), q, res.r, err.e);
    }
    if (!ok) err.raise();
    return results(res, err, withDistance);
  }

  py::object linearSearch(py::object query, size_t size, bool withDistance) {
    py::array_t<float, py::array::c_style | py::array::forcecast> qobject(query);
    py::buffer_info qinfo = qobject.request();
    if ((size_t)qinfo.size != (size_t)dimension_) return dimensionError(withDistance);
    NGTQuery q;
    q.query = static_cast<float *>(qinfo.ptr);
    q.size = size == 0 ? defaultNumOfSearchObjects : size;
    q.radius = defaultRadius >= FLT_MAX ? -1.0f : defaultRadius;
    q.accuracy = -1.0f;
    q.epsilon = 0.0f;
    q.edge_size = (size_t)(int64_t)-1;
    Err err;
    Results res(err);
    bool ok;
    {
      py::gil_scoped_release nogil;
      ok = ngt_linear_search_index_with_query(handle(), q, res.r, err.e);
    }
    if (!ok) err.raise();
    return results(res, err, withDistance);
  }

  // additive: queries [nq, dim] -> (ids [nq, size] int64 in the index's numbering, -1 where a query has fewer results,
  // distances [nq, size] float32). Float queries for both object types; byte arrays go to the uint8 entry point.
  py::tuple batchSearch(py::array queries, size_t size, float epsilon, int edgeSize, float expectedAccuracy) {
    const size_t k = size == 0 ? defaultNumOfSearchObjects : size;
    float eps = epsilon <= -1.0f ? defaultEpsilon : epsilon;
    if (expectedAccuracy > 0.0f) eps = epsilonFromAccuracy(expectedAccuracy);
    const int64_t es = edgeSize < -2 ? defaultEdgeSize : edgeSize;
    const float radius = defaultRadius >= FLT_MAX ? -1.0f : defaultRadius;
    return batch(queries, k, [&](const void *q, bool bytes, uint32_t nq, uint32_t *ids, float *ds, uint32_t *cnt, NGTError e) {
      return bytes ? ngt_batch_search_index_as_uint8(handle(), (const uint8_t *)q, nq, dimension_, k, eps, radius, es, ids, ds, cnt, e)
                   : ngt_batch_search_index_as_float(handle(), (const float *)q, nq, dimension_, k, eps, radius, es, ids, ds, cnt, e);
    });
  }
  py::tuple batchLinearSearch(py::array queries, size_t size) {
    const size_t k = size == 0 ? defaultNumOfSearchObjects : size;
    const float radius = defaultRadius >= FLT_MAX ? -1.0f : defaultRadius;
    return batch(queries, k, [&](const void *q, bool bytes, uint32_t nq, uint32_t *ids, float *ds, uint32_t *cnt, NGTError e) {
      return bytes ? ngt_batch_linear_search_index_as_uint8(handle(), (const uint8_t *)q, nq, dimension_, k, radius, ids, ds, cnt, e)
                   : ngt_batch_linear_search_index_as_float(handle(), (const float *)q, nq, dimension_, k, radius, ids, ds, cnt, e);
    });
  }

  void remove(size_t id) {
    Err err;
    if (!ngt_remove_index(handle(), (ObjectID)(zeroNumbering ? id + 1 : id), err.e)) err.raise();
  }

  void refineANNG(float epsilon, float accuracy, int numOfEdges, int numOfExploredEdges, size_t batchSize) {
    Err err;
    bool ok;
    {
      py::gil_scoped_release nogil;
      ok = ngt_refine_anng(handle(), epsilon, accuracy, numOfEdges, numOfExploredEdges, batchSize, err.e);
    }
    if (!ok) err.raise();
  }

  std::vector<float> getObject(size_t id) {
    Err err;
    NGTObjectSpace space = ngt_get_object_space(handle(), err.e);
    if (!space) err.raise();
    const ObjectID oid = (ObjectID)(zeroNumbering ? id + 1 : id);
    std::vector<float> object;
    object.reserve(dimension_);
    if (integer_) {
      uint8_t *row = ngt_get_object_as_integer(space, oid, err.e);
      if (!row) err.raise();
      object.assign(row, row + dimension_);
    } else {
      float *row = ngt_get_object_as_float(space, oid, err.e);
      if (!row) err.raise();
      object.assign(row, row + dimension_);
    }
    return object;
  }

  void set(size_t numOfSearchObjects, float radius, float epsilon, int edgeSize, float expectedAccuracy) {
    // python/src/ngtpy.cpp:334-347
    defaultNumOfSearchObjects = numOfSearchObjects > 0 ? numOfSearchObjects : defaultNumOfSearchObjects;
    defaultEpsilon = epsilon > -1.0f ? epsilon : defaultEpsilon;
    defaultRadius = radius >= 0.0f ? radius : defaultRadius;
    defaultEdgeSize = edgeSize >= -2 ? edgeSize : defaultEdgeSize;
    defaultExpectedAccuracy = expectedAccuracy > 0.0f ? expectedAccuracy : defaultExpectedAccuracy;
  }

  size_t getNumOfDistanceComputations() { return (size_t)(counter() - distBase_); }

  void save() {
    Err err;
    if (!ngt_save_index(handle(), path_.c_str(), err.e)) err.raise();
  }
  void close() {
    if (index_) ngt_close_index(index_);
    index_ = nullptr;
  }
  void exportIndex(const std::string &) { throw std::runtime_error("export_index: not provided by the B200 engine (text dump, outside the hot path)"); }
  void importIndex(const std::string &) { throw std::runtime_error("import_index: not provided by the B200 engine (text dump, outside the hot path)"); }

private:
  NGTIndex handle() const {
    if (!index_) throw std::runtime_error("ngtpy::Index: the index is closed");
    return index_;
  }
  uint64_t counter() {
    Err err;
    return ngt_get_number_of_distance_computations(handle(), err.e);
  }
  py::object dimensionError(bool withDistance) {
    // the reference prints the allocateObject exception and returns nothing (python/src/ngtpy.cpp:159-168)
    std::cerr << "ObjectSpace::allocateObject: the specified dimension is invalid" << std::endl;
    if (!withDistance) return py::array_t<int>();
    return py::list();
  }
  py::object results(Results &res, Err &err, bool withDistance) {
    const uint32_t n = ngt_get_result_size(res.r, err.e);
    const int off = zeroNumbering ? 1 : 0;
    if (!withDistance) {
      py::array_t<int> ids(n);
      int *p = static_cast<int *>(ids.request().ptr);
      for (uint32_t i = 0; i < n; i++) p[i] = (int)ngt_get_result(res.r, i, err.e).id - off;
      return std::move(ids);
    }
    py::list out;
    for (uint32_t i = 0; i < n; i++) {
      NGTObjectDistance od = ngt_get_result(res.r, i, err.e);
      out.append(py::make_tuple((int)od.id - off, od.distance));
    }
    return std::move(out);
  }
  float epsilonFromAccuracy(float accuracy) {
    // Index::getEpsilonFromExpectedAccuracy (Index.h:1111): the table GraphOptimizer left in `prf` (GraphOptimizer.h:355-365)
    std::ifstream f(path_ + "/prf");
    std::string line, table;
    while (std::getline(f, line)) {
      if (line.compare(0, 14, "AccuracyTable\t") == 0) table = line.substr(14);
    }
    float eps = 0.0f;
    if (ngtgpu_epsilon_from_accuracy_table(table.c_str(), accuracy, &eps) != 0) throw std::runtime_error(ngtgpu_last_error());
    return eps;
  }
  template <class F> py::tuple batch(py::array queries, size_t k, F call) {
    const bool bytes = integer_ && py::isinstance<py::array_t<uint8_t>>(queries);
    py::array q = bytes ? py::array(py::array_t<uint8_t, py::array::c_style | py::array::forcecast>(queries))
                        : py::array(py::array_t<float, py::array::c_style | py::array::forcecast>(queries));
    if (q.ndim() != 2 || q.shape(1) != (py::ssize_t)dimension_) {
      std::stringstream msg;
      msg << "ngtpy::batch_search: queries must be [nq, " << dimension_ << "]";
      throw std::runtime_error(msg.str());
    }
    const uint32_t nq = (uint32_t)q.shape(0);
    std::vector<uint32_t> ids((size_t)nq * k + 1), cnt(nq + 1);
    py::array_t<float> dists({(py::ssize_t)nq, (py::ssize_t)k});
    py::array_t<int64_t> out({(py::ssize_t)nq, (py::ssize_t)k});
    float *dp = static_cast<float *>(dists.request().ptr);
    int64_t *op = static_cast<int64_t *>(out.request().ptr);
    if (nq && k) {
      const void *qp = q.data();
      Err err;
      bool ok;
      {
        py::gil_scoped_release nogil;
        ok = call(qp, bytes, nq, ids.data(), dp, cnt.data(), err.e);
      }
      if (!ok) err.raise();
      const int off = zeroNumbering ? 1 : 0;
      for (uint32_t i = 0; i < nq; i++)
        for (size_t j = 0; j < k; j++) op[(size_t)i * k + j] = j < cnt[i] ? (int64_t)ids[(size_t)i * k + j] - off : -1;
    }
    return py::make_tuple(out, dists);
  }

  std::string path_;
  NGTIndex index_ = nullptr;
  int32_t dimension_ = 0;
  bool integer_ = false;
  uint64_t distBase_ = 0;
  bool zeroNumbering;
  size_t defaultNumOfSearchObjects;
  float defaultEpsilon;
  float defaultRadius;
  int64_t defaultEdgeSize;
  float defaultExpectedAccuracy;
};

class Optimizer {
public:
  Optimizer(int numOfOutgoings, int numOfIncomings, int numOfQueries, int numOfObjects, float lowAccuracyFrom, float lowAccuracyTo,
            float highAccuracyFrom, float highAccuracyTo, double gtEpsilon, double margin, bool logDisabled) {
    Err err;
    opt_ = ngt_create_optimizer(logDisabled, err.e);
    if (!opt_) err.raise();
    set(numOfOutgoings, numOfIncomings, numOfQueries, numOfObjects, lowAccuracyFrom, lowAccuracyTo, highAccuracyFrom, highAccuracyTo,
        gtEpsilon, margin);
  }
  ~Optimizer() {
    if (opt_) ngt_destroy_optimizer(opt_);
  }
  Optimizer(const Optimizer &) = delete;
  Optimizer &operator=(const Optimizer &) = delete;

  void set(int outgoing, int incoming, int nofqs, int nofrs, float baseAccuracyFrom, float baseAccuracyTo, float rateAccuracyFrom,
           float rateAccuracyTo, double gte, double m) {
    (void)nofrs;   // GraphOptimizer::set, GraphOptimizer.h:600-628: negative values leave a setting as it is
    Err err;
    if (!ngt_optimizer_set(opt_, outgoing, incoming, nofqs, baseAccuracyFrom, baseAccuracyTo, rateAccuracyFrom, rateAccuracyTo, gte, m,
                           err.e))
      err.raise();
  }
  void setProcessingModes(bool shortcut, bool searchParameter, bool prefetchParameter, bool accuracyTable) {
    Err err;
    if (!ngt_optimizer_set_shortcut_reduction(opt_, shortcut, err.e) ||
        !ngt_optimizer_set_processing_modes(opt_, searchParameter, prefetchParameter, accuracyTable, err.e))
      err.raise();
  }
  void execute(const std::string &inPath, const std::string &outPath) {
    Err err;
    bool ok;
    {
      py::gil_scoped_release nogil;
      ok = ngt_optimizer_execute(opt_, inPath.c_str(), outPath.c_str(), err.e);
    }
    if (!ok) err.raise();
  }
  void adjustSearchCoefficients(const std::string &path) {
    Err err;
    if (!ngt_optimizer_adjust_search_coefficients(opt_, path.c_str(), err.e)) err.raise();
  }
  void optimizeSearchParameters(const std::string &path) { adjustSearchCoefficients(path); }   // GraphOptimizer.h:302-350: the same timed tuning
  int optimizeNumberOfEdgesForANNG(const std::string &path, int numOfQueries, int numOfResults, int numOfThreads, float targetAccuracy,
                                   int targetNoOfObjects, int numOfSampleObjects, int maxNoOfEdges) {
    NGTAnngEdgeOptimizationParameter p = ngt_get_anng_edge_optimization_parameter();   // python/src/ngtpy.cpp:365-385
    if (numOfQueries > 0) p.no_of_queries = numOfQueries;
    if (numOfResults > 0) p.no_of_results = numOfResults;
    if (numOfThreads >= 0) p.no_of_threads = numOfThreads;
    if (targetAccuracy > 0.0f) p.target_accuracy = targetAccuracy;
    if (targetNoOfObjects >= 0) p.target_no_of_objects = targetNoOfObjects;
    if (numOfSampleObjects >= 0) p.no_of_sample_objects = numOfSampleObjects;
    if (maxNoOfEdges >= 0) p.max_of_no_of_edges = maxNoOfEdges;
    Err err;
    if (!ngt_optimize_number_of_edges(path.c_str(), p, err.e)) err.raise();
    return 0;
  }

private:
  NGTOptimizer opt_ = nullptr;
};

class QuantizedIndex {   // NGTQG (lib/NGT/NGTQ/QuantizedGraph.h): a different algorithm, outside the hot path
public:
  QuantizedIndex(const std::string &, size_t, bool, bool, bool) {
    throw std::runtime_error("ngtpy.QuantizedIndex: not provided by the B200 engine (quantized graphs are outside the hot path)");
  }
};

}  // namespace

PYBIND11_MODULE(ngtpy, m) {
  m.doc() = "ngt python (B200 engine behind NGT's C API)";
  m.attr("__version__") = "1.13.8";   // the reference version whose ngtpy surface this module reproduces (VERSION)

  m.def("create", &Index::create, py::arg("path"), py::arg("dimension"), py::arg("edge_size_for_creation") = 10,
        py::arg("edge_size_for_search") = 40, py::arg("distance_type") = "L2", py::arg("object_type") = "Float");

  py::class_<Index>(m, "Index")
      .def(py::init<const std::string &, bool, bool, bool, bool>(), py::arg("path"), py::arg("read_only") = false,
           py::arg("zero_based_numbering") = true, py::arg("tree_disabled") = false, py::arg("log_disabled") = false)
      .def("search", &Index::search, py::arg("query"), py::arg("size") = 0, py::arg("epsilon") = -FLT_MAX,
           py::arg("edge_size") = INT_MIN, py::arg("expected_accuracy") = -FLT_MAX, py::arg("with_distance") = true)
      .def("linear_search", &Index::linearSearch, py::arg("query"), py::arg("size") = 0, py::arg("with_distance") = true)
      .def("batch_search", &Index::batchSearch, py::arg("queries"), py::arg("size") = 0, py::arg("epsilon") = -FLT_MAX,
           py::arg("edge_size") = INT_MIN, py::arg("expected_accuracy") = -FLT_MAX)
      .def("batch_linear_search", &Index::batchLinearSearch, py::arg("queries"), py::arg("size") = 0)
      .def("get_num_of_distance_computations", &Index::getNumOfDistanceComputations)
      .def("save", &Index::save)
      .def("close", &Index::close)
      .def("remove", &Index::remove, py::arg("object_id"))
      .def("build_index", &Index::buildIndex, py::arg("num_threads") = 8, py::arg("target_size_of_graph") = 0)
      .def("get_object", &Index::getObject, py::arg("object_id"))
      .def("batch_insert", &Index::batchInsert, py::arg("objects"), py::arg("num_threads") = 8, py::arg("debug") = false)
      .def("insert", &Index::insert, py::arg("object"), py::arg("debug") = false)
      .def("refine_anng", &Index::refineANNG, py::arg("epsilon") = 0.1, py::arg("expected_accuracy") = 0.0,
           py::arg("num_of_edges") = 0, py::arg("num_of_explored_edges") = INT_MIN, py::arg("batch_size") = 10000)
      .def("set", &Index::set, py::arg("num_of_search_objects") = 0, py::arg("search_radius") = -FLT_MAX,
           py::arg("epsilon") = -FLT_MAX, py::arg("edge_size") = INT_MIN, py::arg("expected_accuracy") = -FLT_MAX)
      .def("export_index", &Index::exportIndex, py::arg("path"))
      .def("import_index", &Index::importIndex, py::arg("path"));

  py::class_<Optimizer>(m, "Optimizer")
      .def(py::init<int, int, int, int, float, float, float, float, double, double, bool>(), py::arg("num_of_outgoings") = -1,
           py::arg("num_of_incomings") = -1, py::arg("num_of_queries") = -1, py::arg("num_of_objects") = -1,
           py::arg("low_accuracy_from") = -1.0, py::arg("low_accuracy_to") = -1.0, py::arg("high_accuracy_from") = -1.0,
           py::arg("high_accuracy_to") = -1.0, py::arg("gt_epsilon") = -DBL_MAX, py::arg("margin") = -1.0,
           py::arg("log_disabled") = false)
      .def("execute", &Optimizer::execute, py::arg("in_path"), py::arg("out_path"))
      .def("adjust_search_coefficients", &Optimizer::adjustSearchCoefficients, py::arg("path"))
      .def("set", &Optimizer::set, py::arg("num_of_outgoings") = -1, py::arg("num_of_incomings") = -1,
           py::arg("num_of_queries") = -1, py::arg("num_of_objects") = -1, py::arg("low_accuracy_from") = -1.0,
           py::arg("low_accuracy_to") = -1.0, py::arg("high_accuracy_from") = -1.0, py::arg("high_accuracy_to") = -1.0,
           py::arg("gt_epsilon") = -DBL_MAX, py::arg("margin") = -1.0)
      .def("set_processing_modes", &Optimizer::setProcessingModes, py::arg("shortcut_reduction") = true,
           py::arg("search_parameter_optimization") = true, py::arg("prefetch_parameter_optimization") = true,
           py::arg("accuracy_table_generation") = true)
      .def("optimize_search_parameters", &Optimizer::optimizeSearchParameters, py::arg("path"))
      .def("optimize_number_of_edges_for_anng", &Optimizer::optimizeNumberOfEdgesForANNG, py::arg("path"),
           py::arg("num_of_queries") = -1, py::arg("num_of_results") = -1, py::arg("num_of_threads") = -1,
           py::arg("target_accuracy") = -1, py::arg("target_num_of_objects") = -1, py::arg("num_of_sample_objects") = -1,
           py::arg("max_num_of_edges") = -1);

  py::class_<QuantizedIndex>(m, "QuantizedIndex")
      .def(py::init<const std::string &, size_t, bool, bool, bool>(), py::arg("path"), py::arg("max_no_of_edges") = 128,
           py::arg("zero_based_numbering") = true, py::arg("tree_disabled") = false, py::arg("log_disabled") = false);
}

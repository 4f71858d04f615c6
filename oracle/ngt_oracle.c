/*
 * ngt_oracle.c -- CPU restatement of NGT v1.13.8's hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the B200 engine. It is NOT product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (ngt_b200/, libngtgpu.so) never links, imports or calls anything in oracle/.
 *
 * Parity pin: every function here is checked (tests/test_oracle_pin.py, `-m "not gpu"`) against
 *   (1) the reference's only known-answer listing, bin/ngt/README.md:254-323, through
 *       tests/golden/readme_kat.json, and
 *   (2) outputs of the reference itself compiled from /root/reference into oracle/_ref/
 *       (oracle/Makefile, oracle/ref_shim.cpp), committed as tests/golden/*.npz by
 *       tests/golden/make_golden.py.
 *
 * All file:line citations are relative to /root/reference/lib/NGT/.
 *
 * Written from the reference's published behaviour; no reference source is copied. The lane
 * structure of the float kernels is restated as "W independent partial sums, folded pairwise"
 * because that is what decides the float rounding of the reference's SIMD code
 * (PrimitiveComparator.h:143-198: W=16 for AVX-512, 8 for AVX2, 4 for SSE).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

/* ---- enums mirror ObjectSpace.h:166-186 ------------------------------------------------ */
enum { NGTO_L1 = 0, NGTO_L2 = 1, NGTO_HAMMING = 2, NGTO_ANGLE = 3, NGTO_COSINE = 4,
       NGTO_NORMALIZED_ANGLE = 5, NGTO_NORMALIZED_COSINE = 6, NGTO_JACCARD = 7,
       NGTO_NORMALIZED_L2 = 9 };
enum { NGTO_UINT8 = 1, NGTO_FLOAT = 2 };

/* SIMD shape of the reference build being restated. Default = the oracle/_ref build
 * (-march=x86-64-v3 => NGT_AVX2, defines.h.in:47-58; gcc contracts add(mul) into FMA). */
static int g_lanes = 8;
static int g_fma = 1;
void ngto_set_simd(int lanes, int fma) { g_lanes = lanes; g_fma = fma; }

/* ObjectSpace.h:249 */
size_t ngto_padded_dimension(size_t dim) { return ((dim - 1) / 16 + 1) * 16; }

static inline float acc_step(float acc, float x, float y) {
  /* acc + x*y with or without contraction */
  if (g_fma) return fmaf(x, y, acc);
  volatile float p = x * y;
  return acc + p;
}

/* Fold W lane sums the way the reference does: halves added until 4 lanes remain
 * (PrimitiveComparator.h:153-154,165), then f[0]+f[1]+f[2]+f[3] left to right in float (:193). */
static float fold_lanes_float(float *s, int w) {
  while (w > 4) {
    w /= 2;
    for (int i = 0; i < w; i++) s[i] = s[i] + s[i + w];
  }
  float r = s[0] + s[1];
  r = r + s[2];
  r = r + s[3];
  return r;
}

/* PrimitiveComparator.h:143-198 (float L2). size = padded dimension (multiple of 16). */
double ngto_l2_float(const float *a, const float *b, size_t size) {
  float s[16] = {0};
  int w = g_lanes;
  /* AVX2/SSE variants unroll but keep W accumulators; order per lane is index order. */
  for (size_t i = 0; i < size; i += w)
    for (int l = 0; l < w; l++) {
      float v = a[i + l] - b[i + l];
      s[l] = acc_step(s[l], v, v);
    }
  double sum = fold_lanes_float(s, w);
  return sqrt(sum);
}

/* PrimitiveComparator.h:200-223 (uint8 L2): exact integer squares accumulated in 4 float lanes
 * (exact while < 2^24), folded in float, remainder loop in double, sqrt in double. */
double ngto_l2_uint8(const uint8_t *a, const uint8_t *b, size_t size) {
  float s[4] = {0, 0, 0, 0};
  size_t i = 0;
  for (; i + 7 < size; i += 8) {
    for (int l = 0; l < 4; l++) {
      int d0 = (int)a[i + l] - (int)b[i + l];
      s[l] = s[l] + (float)(d0 * d0);
    }
    for (int l = 0; l < 4; l++) {
      int d1 = (int)a[i + 4 + l] - (int)b[i + 4 + l];
      s[l] = s[l] + (float)(d1 * d1);
    }
  }
  float f = s[0] + s[1];
  f = f + s[2];
  f = f + s[3];
  double sum = f;
  for (; i < size; i++) {
    int d = (int)a[i] - (int)b[i];
    sum += d * d;
  }
  return sqrt(sum);
}

/* PrimitiveComparator.h:340-353: popcount over 64-bit words, two per iteration. */
double ngto_hamming(const uint8_t *a, const uint8_t *b, size_t size) {
  size_t count = 0;
  for (size_t i = 0; i < size; i += 8) {
    uint64_t x, y;
    memcpy(&x, a + i, 8);
    memcpy(&y, b + i, 8);
    count += (size_t)__builtin_popcountll(x ^ y);
  }
  return (double)count;
}

/* PrimitiveComparator.h:446-477: lanes folded in float down to 4, the last 4 summed in double. */
double ngto_dot_float(const float *a, const float *b, size_t size) {
  float s[16] = {0};
  int w = g_lanes;
  for (size_t i = 0; i < size; i += w)
    for (int l = 0; l < w; l++) s[l] = acc_step(s[l], a[i + l], b[i + l]);
  while (w > 4) {
    w /= 2;
    for (int i = 0; i < w; i++) s[i] = s[i] + s[i + w];
  }
  return (double)s[0] + (double)s[1] + (double)s[2] + (double)s[3];
}

/* PrimitiveComparator.h:487-553: three accumulators, float fold, s / sqrt(na*nb) in double. */
double ngto_cosine_float(const float *a, const float *b, size_t size) {
  float na[16] = {0}, nb[16] = {0}, s[16] = {0};
  int w = g_lanes;
  for (size_t i = 0; i < size; i += w)
    for (int l = 0; l < w; l++) {
      na[l] = acc_step(na[l], a[i + l], a[i + l]);
      nb[l] = acc_step(nb[l], b[i + l], b[i + l]);
      s[l] = acc_step(s[l], a[i + l], b[i + l]);
    }
  double dna = fold_lanes_float(na, w);
  double dnb = fold_lanes_float(nb, w);
  double ds = fold_lanes_float(s, w);
  return ds / sqrt(dna * dnb);
}

static double clamp_acos(double c) { /* PrimitiveComparator.h:571-593 */
  if (c >= 1.0) return 0.0;
  if (c <= -1.0) return acos(-1.0);
  return acos(c);
}

/* The comparator a given (distance type, object type) resolves to:
 * ObjectSpaceRepository.h:346-441 (writable path) and Graph.h:290-350 (read-only path) agree. */
double ngto_distance(int dtype, int otype, const void *a, const void *b, size_t padded) {
  if (otype == NGTO_UINT8) {
    switch (dtype) {
      case NGTO_HAMMING: return ngto_hamming((const uint8_t *)a, (const uint8_t *)b, padded);
      default: return ngto_l2_uint8((const uint8_t *)a, (const uint8_t *)b, padded);
    }
  }
  const float *fa = (const float *)a, *fb = (const float *)b;
  switch (dtype) {
    case NGTO_ANGLE: return clamp_acos(ngto_cosine_float(fa, fb, padded));
    case NGTO_COSINE: return 1.0 - ngto_cosine_float(fa, fb, padded);         /* :639-642 */
    case NGTO_NORMALIZED_ANGLE: return clamp_acos(ngto_dot_float(fa, fb, padded));
    case NGTO_NORMALIZED_COSINE: {                                             /* :644-648 */
      double v = 1.0 - ngto_dot_float(fa, fb, padded);
      return v < 0.0 ? 0.0 : v;
    }
    case NGTO_NORMALIZED_L2: {                                                 /* :226-234 */
      double v = 2.0 - 2.0 * ngto_dot_float(fa, fb, padded);
      return v < 0.0 ? 0.0 : sqrt(v);
    }
    default: return ngto_l2_float(fa, fb, padded);
  }
}

/* ObjectSpace.h:251-266: float accumulator, divide each element; returns -1 on a zero vector
 * (the reference throws). */
int ngto_normalize(float *data, size_t dim) {
  float sum = 0.0f;
  for (size_t i = 0; i < dim; i++) sum += data[i] * data[i];
  if (sum == 0.0f) return -1;
  sum = sqrtf(sum);
  for (size_t i = 0; i < dim; i++) data[i] = data[i] / sum;
  return 0;
}

/* ---- (distance,id) ordering, Common.h:1946-1959 --------------------------------------- */
typedef struct { uint32_t id; float distance; } od_t;
static inline int od_less(od_t a, od_t b) {
  if (a.distance == b.distance) return a.id < b.id;
  return a.distance < b.distance;
}

/* binary heap, max-at-top when `maxheap`, min-at-top otherwise. The order is total on
 * (distance,id) with unique ids, so any correct heap pops the same sequence as
 * std::priority_queue does in the reference. */
typedef struct { od_t *v; size_t n, cap; int maxheap; } heap_t;
static int heap_before(const heap_t *h, od_t a, od_t b) { return h->maxheap ? od_less(b, a) : od_less(a, b); }
static void heap_push(heap_t *h, od_t x) {
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->v = (od_t *)realloc(h->v, h->cap * sizeof(od_t)); }
  size_t i = h->n++;
  while (i > 0) {
    size_t p = (i - 1) / 2;
    if (!heap_before(h, x, h->v[p])) break;
    h->v[i] = h->v[p];
    i = p;
  }
  h->v[i] = x;
}
static od_t heap_pop(heap_t *h) {
  od_t top = h->v[0];
  od_t x = h->v[--h->n];
  size_t i = 0;
  for (;;) {
    size_t c = 2 * i + 1;
    if (c >= h->n) break;
    if (c + 1 < h->n && heap_before(h, h->v[c + 1], h->v[c])) c++;
    if (!heap_before(h, h->v[c], x)) break;
    h->v[i] = h->v[c];
    i = c;
  }
  if (h->n) h->v[i] = x;
  return top;
}

static int od_cmp_qsort(const void *pa, const void *pb) {
  od_t a = *(const od_t *)pa, b = *(const od_t *)pb;
  return od_less(a, b) ? -1 : (od_less(b, a) ? 1 : 0);
}

/* ---- linearSearch, ObjectSpaceRepository.h:466-502 ------------------------------------ */
/* objects: (n+1) rows of `stride` bytes, row 0 is the dummy slot (Common.h:1704 repository is
 * 1-based); valid[id]==0 marks a removed/empty slot (skipped, :485). radius<0 disables the filter.
 * Output ascending by (distance,id) (ObjectSpace.h:49-57), returns number of results. */
size_t ngto_linear_search(int dtype, int otype, const void *objects, size_t stride, size_t n,
                          const uint8_t *valid, size_t padded, const void *query, double radius,
                          size_t k, uint32_t *out_ids, float *out_dists) {
  heap_t res = {0, 0, 0, 1};
  const uint8_t *base = (const uint8_t *)objects;
  for (size_t idx = 1; idx <= n; idx++) {
    if (valid && !valid[idx]) continue;
    float d = (float)ngto_distance(dtype, otype, query, base + idx * stride, padded);
    if (radius < 0.0 || d <= radius) {
      od_t o = {(uint32_t)idx, d};
      heap_push(&res, o);
      if (res.n > k) heap_pop(&res);
    }
  }
  size_t m = res.n;
  for (size_t i = m; i-- > 0;) {
    od_t o = heap_pop(&res);
    out_ids[i] = o.id;
    out_dists[i] = o.distance;
  }
  free(res.v);
  return m;
}

/* ---- getEdgeSize, Graph.h:675-692 ------------------------------------------------------ */
/* returns the per-node edge cap; INT_MAX means "all edges"; -1 signals invalid parameters. */
int64_t ngto_edge_size(int64_t sc_edge_size, int64_t prop_edge_size_for_search,
                       float exploration_coefficient, int64_t dyn_base, int64_t dyn_rate) {
  int64_t esize = sc_edge_size == -1 ? prop_edge_size_for_search : sc_edge_size;
  if (esize == 0) return INT_MAX;
  if (esize > 0) return esize;
  if (esize == -2) {
    double add = pow(10, (exploration_coefficient - 1.0) * (float)dyn_rate);
    if (add >= (double)INT_MAX) return INT_MAX;
    return (int64_t)(size_t)((double)dyn_base + add);
  }
  return -1;
}

/* ---- graph search, Graph.cpp:398-495 (read-only) == Graph.cpp:499-638 (writable) ------- */
/* graph: CSR over ids 0..n (row 0 empty): row_ptr[id]..row_ptr[id+1] into col[], each list
 * ascending by (distance,id) as stored in `grp`. seeds: explicit ids (Index.h:1140 takes them).
 * epsilon -> explorationCoefficient = (float)(e + 1.0) (Common.h:2041).
 * stats[0] = distance computations incl. seeds (Graph.cpp:289,464), stats[1] = adjacency entries
 * examined (Graph.cpp:590), stats[2] = nodes expanded. */
size_t ngto_graph_search(int dtype, int otype, const void *objects, size_t stride, size_t n,
                         size_t padded, const uint64_t *row_ptr, const uint32_t *col,
                         const void *query, const uint32_t *seeds, size_t nseeds, size_t k,
                         float epsilon, float radius_in, int64_t edge_size,
                         uint32_t *out_ids, float *out_dists, uint64_t *stats) {
  if (stats) stats[0] = stats[1] = stats[2] = 0;
  if (k == 0) return 0;                                      /* Index.h:1141-1144 */
  float coef = (float)(epsilon + 1.0);
  float radius = radius_in < 0.0f ? FLT_MAX : radius_in;     /* Capi.cpp:384-386 */
  const uint8_t *base = (const uint8_t *)objects;
  uint8_t *checked = (uint8_t *)calloc(n + 1, 1);            /* Graph.cpp:412, Graph.h:751-755 */
  heap_t unchecked = {0, 0, 0, 0}, results = {0, 0, 0, 1};
  od_t *sd = (od_t *)malloc((nseeds ? nseeds : 1) * sizeof(od_t));
  /* setupDistances, Graph.cpp:292-338 */
  for (size_t i = 0; i < nseeds; i++) {
    sd[i].id = seeds[i];
    sd[i].distance = (float)ngto_distance(dtype, otype, query, base + (size_t)seeds[i] * stride, padded);
  }
  if (stats) stats[0] += nseeds;
  /* setupSeeds, Graph.cpp:341-366 */
  qsort(sd, nseeds, sizeof(od_t), od_cmp_qsort);
  for (size_t i = 0; i < nseeds; i++) {
    if (results.n < k && sd[i].distance <= radius) heap_push(&results, sd[i]);
    else break;
  }
  if (results.n >= k) radius = results.v[0].distance;
  for (size_t i = 0; i < nseeds; i++) {
    checked[sd[i].id] = 1;
    heap_push(&unchecked, sd[i]);
  }
  float exploration_radius = coef * radius;                  /* Graph.cpp:420 */
  while (unchecked.n) {
    od_t target = heap_pop(&unchecked);
    if (target.distance > exploration_radius) break;         /* :433 */
    uint64_t b = row_ptr[target.id], e = row_ptr[target.id + 1];
    uint64_t deg = e - b;
    if ((int64_t)deg > edge_size) deg = (uint64_t)edge_size; /* :438 */
    if (stats) { stats[1] += deg; stats[2] += 1; }
    for (uint64_t j = b; j < b + deg; j++) {
      uint32_t nid = col[j];
      if (checked[nid]) continue;
      checked[nid] = 1;
      float d = (float)ngto_distance(dtype, otype, query, base + (size_t)nid * stride, padded);
      if (stats) stats[0] += 1;
      if (d <= exploration_radius) {                         /* :471-483 */
        od_t r = {nid, d};
        heap_push(&unchecked, r);
        if (d <= radius) {
          heap_push(&results, r);
          if (results.n >= k) {
            if (results.n > k) heap_pop(&results);
            radius = results.v[0].distance;
            exploration_radius = coef * radius;
          }
        }
      }
    }
  }
  size_t m = results.n;
  for (size_t i = m; i-- > 0;) {
    od_t o = heap_pop(&results);
    out_ids[i] = o.id;
    out_dists[i] = o.distance;
  }
  free(results.v); free(unchecked.v); free(sd); free(checked);
  return m;
}

/* ---- recall as the reference defines it, Optimizer.h:400,496-507 ----------------------- */
/* a returned row is relevant if its id is in the ground truth OR its distance <= the farthest
 * ground-truth distance (when that is > 0); accuracy = relevant / |ground truth|. */
double ngto_recall(const uint32_t *ids, const float *dists, size_t nres,
                   const uint32_t *gt_ids, const float *gt_dists, size_t ngt) {
  if (ngt == 0) return 0.0;
  double farthest = gt_dists[ngt - 1];
  size_t relevant = 0;
  for (size_t i = 0; i < nres; i++) {
    int hit = 0;
    for (size_t j = 0; j < ngt; j++) if (gt_ids[j] == ids[i]) { hit = 1; break; }
    if (hit) relevant++;
    else if (farthest > 0.0 && (double)dists[i] <= farthest) relevant++;
  }
  return (double)relevant / (double)ngt;
}

/* ---- batch drivers (OpenMP when compiled with -fopenmp) used by the cpu_baseline leg ---- */
void ngto_batch_linear_search(int dtype, int otype, const void *objects, size_t stride, size_t n,
                              const uint8_t *valid, size_t padded, const void *queries,
                              size_t qstride, size_t nq, double radius, size_t k,
                              uint32_t *out_ids, float *out_dists, uint32_t *out_counts) {
#pragma omp parallel for schedule(dynamic, 1)
  for (long q = 0; q < (long)nq; q++) {
    out_counts[q] = (uint32_t)ngto_linear_search(dtype, otype, objects, stride, n, valid, padded,
                                                 (const uint8_t *)queries + (size_t)q * qstride, radius, k,
                                                 out_ids + (size_t)q * k, out_dists + (size_t)q * k);
  }
}

void ngto_batch_graph_search(int dtype, int otype, const void *objects, size_t stride, size_t n,
                             size_t padded, const uint64_t *row_ptr, const uint32_t *col,
                             const void *queries, size_t qstride, size_t nq, const uint32_t *seeds,
                             size_t nseeds, size_t k, float epsilon, float radius, int64_t edge_size,
                             uint32_t *out_ids, float *out_dists, uint32_t *out_counts,
                             uint64_t *out_stats) {
#pragma omp parallel for schedule(dynamic, 8)
  for (long q = 0; q < (long)nq; q++) {
    uint64_t st[3];
    out_counts[q] = (uint32_t)ngto_graph_search(
        dtype, otype, objects, stride, n, padded, row_ptr, col,
        (const uint8_t *)queries + (size_t)q * qstride, seeds + (size_t)q * nseeds, nseeds, k, epsilon,
        radius, edge_size, out_ids + (size_t)q * k, out_dists + (size_t)q * k, st);
    if (out_stats) { out_stats[3 * q] = st[0]; out_stats[3 * q + 1] = st[1]; out_stats[3 * q + 2] = st[2]; }
  }
}

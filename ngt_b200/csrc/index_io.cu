// index_io.cu -- host-side readers/writers of NGT's on-disk `obj` and `grp` files (no device code).
//
// Layout (lib/NGT/Common.h:1776-1837 Repository::serialize, ObjectSpace.h:293-301, Graph.h:151-158,
// Common.h:706-712, :1960-1967):
//   obj = u64 slots, then per slot '-' (empty; slot 0 always) or '+' followed by dimension*sizeof(T) raw
//         bytes of the object (unpadded).
//   grp = u64 slots, then per slot '-' or '+', u32 degree, degree x {u32 id, f32 distance} in ascending
//         (distance,id) order; then u32 count and count x u16 `prevsize`.
// Everything is streamed through one large buffer instead of the reference's per-object `new` + stream
// reads, so a 1M x 128 index loads in well under a second of host time.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ngtgpu_internal.cuh"

namespace {
struct File {
  FILE *f = nullptr;
  explicit File(const char *path, const char *mode) { f = fopen(path, mode); }
  ~File() {
    if (f) fclose(f);
  }
};

struct Reader {
  FILE *f;
  std::vector<uint8_t> buf;
  size_t pos = 0, len = 0;
  explicit Reader(FILE *file) : f(file), buf((size_t)8 << 20) {}
  bool read(void *dst, size_t n) {
    uint8_t *d = static_cast<uint8_t *>(dst);
    while (n) {
      if (pos == len) {
        len = fread(buf.data(), 1, buf.size(), f);
        pos = 0;
        if (len == 0) return false;
      }
      size_t m = len - pos < n ? len - pos : n;
      if (d) {
        memcpy(d, buf.data() + pos, m);
        d += m;
      }
      pos += m;
      n -= m;
    }
    return true;
  }
};

struct Writer {
  FILE *f;
  std::vector<uint8_t> buf;
  size_t pos = 0;
  bool ok = true;
  explicit Writer(FILE *file) : f(file), buf((size_t)8 << 20) {}
  void flush() {
    if (pos && fwrite(buf.data(), 1, pos, f) != pos) ok = false;
    pos = 0;
  }
  void write(const void *src, size_t n) {
    const uint8_t *s = static_cast<const uint8_t *>(src);
    while (n) {
      if (pos == buf.size()) flush();
      size_t m = buf.size() - pos < n ? buf.size() - pos : n;
      memcpy(buf.data() + pos, s, m);
      pos += m;
      s += m;
      n -= m;
    }
  }
};
}  // namespace

// slots = repository size (n + 1); present = number of '+' records
extern "C" int ngtgpu_io_obj_info(const char *path, uint32_t record_bytes, uint64_t *slots, uint64_t *present) {
  File fp(path, "rb");
  if (!fp.f) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("cannot open ") + path);
  Reader r(fp.f);
  uint64_t s = 0;
  if (!r.read(&s, 8)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated obj file ") + path);
  uint64_t p = 0;
  for (uint64_t i = 0; i < s; i++) {
    char t = 0;
    if (!r.read(&t, 1)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated obj file ") + path);
    if (t == '+') {
      if (!r.read(nullptr, record_bytes)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated obj file ") + path);
      p++;
    } else if (t != '-') {
      NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("bad record marker in ") + path + " (wrong dimension or object type?)");
    }
  }
  *slots = s;
  if (present) *present = p;
  return NGTGPU_OK;
}

// rows: (slots-1) x record_bytes for ids 1..slots-1 (zero filled where empty); present: slots bytes.
extern "C" int ngtgpu_io_read_obj(const char *path, uint32_t record_bytes, void *rows, uint8_t *present) {
  File fp(path, "rb");
  if (!fp.f) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("cannot open ") + path);
  Reader r(fp.f);
  uint64_t s = 0;
  if (!r.read(&s, 8)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated obj file ") + path);
  uint8_t *out = static_cast<uint8_t *>(rows);
  for (uint64_t i = 0; i < s; i++) {
    char t = 0;
    if (!r.read(&t, 1)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated obj file ") + path);
    if (present) present[i] = t == '+';
    if (t == '+') {
      void *dst = i == 0 ? nullptr : out + (i - 1) * record_bytes;
      if (!r.read(dst, record_bytes)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated obj file ") + path);
    } else if (t == '-') {
      if (i) memset(out + (i - 1) * record_bytes, 0, record_bytes);
    } else {
      NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("bad record marker in ") + path);
    }
  }
  return NGTGPU_OK;
}

// rows: n x record_bytes for ids 1..n; present (nullable): n+1 bytes.
extern "C" int ngtgpu_io_write_obj(const char *path, uint32_t record_bytes, const void *rows, uint64_t n,
                                   const uint8_t *present) {
  File fp(path, "wb");
  if (!fp.f) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("cannot create ") + path);
  Writer w(fp.f);
  uint64_t s = n + 1;
  w.write(&s, 8);
  const char minus = '-', plus = '+';
  w.write(&minus, 1);
  const uint8_t *in = static_cast<const uint8_t *>(rows);
  for (uint64_t i = 1; i <= n; i++) {
    if (present && !present[i]) {
      w.write(&minus, 1);
      continue;
    }
    w.write(&plus, 1);
    w.write(in + (i - 1) * record_bytes, record_bytes);
  }
  w.flush();
  if (!w.ok) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("short write to ") + path);
  return NGTGPU_OK;
}

extern "C" int ngtgpu_io_grp_info(const char *path, uint64_t *slots, uint64_t *nnz) {
  File fp(path, "rb");
  if (!fp.f) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("cannot open ") + path);
  Reader r(fp.f);
  uint64_t s = 0, e = 0;
  if (!r.read(&s, 8)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated grp file ") + path);
  for (uint64_t i = 0; i < s; i++) {
    char t = 0;
    if (!r.read(&t, 1)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated grp file ") + path);
    if (t == '+') {
      uint32_t d = 0;
      if (!r.read(&d, 4) || !r.read(nullptr, (size_t)d * 8)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated grp file ") + path);
      e += d;
    } else if (t != '-') {
      NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("bad record marker in ") + path);
    }
  }
  *slots = s;
  *nnz = e;
  return NGTGPU_OK;
}

// row_ptr: slots+1 entries over ids 0..slots-1; col/dist: nnz; present (nullable): slots bytes.
extern "C" int ngtgpu_io_read_grp(const char *path, uint64_t *row_ptr, uint32_t *col, float *dist, uint8_t *present) {
  File fp(path, "rb");
  if (!fp.f) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("cannot open ") + path);
  Reader r(fp.f);
  uint64_t s = 0, e = 0;
  if (!r.read(&s, 8)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated grp file ") + path);
  std::vector<uint32_t> tmp;
  for (uint64_t i = 0; i < s; i++) {
    row_ptr[i] = e;
    char t = 0;
    if (!r.read(&t, 1)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated grp file ") + path);
    if (present) present[i] = t == '+';
    if (t != '+') continue;
    uint32_t d = 0;
    if (!r.read(&d, 4)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated grp file ") + path);
    tmp.resize((size_t)d * 2);
    if (d && !r.read(tmp.data(), (size_t)d * 8)) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("truncated grp file ") + path);
    for (uint32_t j = 0; j < d; j++) {
      col[e + j] = tmp[2 * j];
      if (dist) memcpy(&dist[e + j], &tmp[2 * j + 1], 4);
    }
    e += d;
  }
  row_ptr[s] = e;
  return NGTGPU_OK;
}

// row_ptr over ids 0..n (n+2 entries, as ngtgpu_index_set_graph takes it); present (nullable): n+1 bytes.
extern "C" int ngtgpu_io_write_grp(const char *path, uint64_t n, const uint64_t *row_ptr, const uint32_t *col,
                                   const float *dist, const uint8_t *present) {
  File fp(path, "wb");
  if (!fp.f) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("cannot create ") + path);
  Writer w(fp.f);
  uint64_t s = n + 1;
  w.write(&s, 8);
  const char minus = '-', plus = '+';
  w.write(&minus, 1);
  for (uint64_t i = 1; i <= n; i++) {
    if (present && !present[i]) {
      w.write(&minus, 1);
      continue;
    }
    w.write(&plus, 1);
    uint32_t d = (uint32_t)(row_ptr[i + 1] - row_ptr[i]);
    w.write(&d, 4);
    for (uint64_t j = row_ptr[i]; j < row_ptr[i + 1]; j++) {
      w.write(&col[j], 4);
      w.write(&dist[j], 4);
    }
  }
  // prevsize (Graph.h:151-154): one u16 per slot, zero = nothing inserted since the last truncation
  uint32_t cnt = (uint32_t)s;
  w.write(&cnt, 4);
  std::vector<uint16_t> zeros(s, 0);
  w.write(zeros.data(), zeros.size() * 2);
  w.flush();
  if (!w.ok) NGTGPU_FAIL(NGTGPU_ERR_INVALID, std::string("short write to ") + path);
  return NGTGPU_OK;
}

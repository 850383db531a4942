// Text formats of the reference harness.
//
//   `.mtx.rcm.csr` / `.mtx.csr`   helpers/converter.m:42-48 (writer), spmv-csr/spmv.c:11-57,
//                                 spmv-csrk/spmv.cpp:32-79, cuda-spmv-csr/spmv.cu:11-73 (readers)
//   `.csr3`                       reformat-csr-to-csr3/spmv-auto.cpp:30-65 (writer),
//                                 reformat-csr-to-csr3/stats.c:10-79 (reader)
//
// The reference readers call fscanf once per token (about 1 s per 5 M nonzeros).  Here the file
// is read in one piece and tokens are converted by a hand-rolled scanner: integers directly;
// decimals through the exact fast path (<= 15 significant digits and a power of ten <= 22 are
// both exact doubles, so one IEEE division gives the correctly rounded result -- the same value
// strtod / fscanf("%lf") return); anything else falls back to strtod / strtof.
#include "host_index.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace hspmv {

namespace {

struct Scanner {
  const char *p, *end;
  bool ok = true;

  void skip() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool next_int(int64_t &out) {
    skip();
    if (p >= end) return ok = false;
    bool neg = false;
    if (*p == '-' || *p == '+') neg = (*p++ == '-');
    if (p >= end || *p < '0' || *p > '9') return ok = false;
    int64_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
    out = neg ? -v : v;
    return true;
  }
  template <typename T> bool next_real(T &out) {
    skip();
    if (p >= end) return ok = false;
    const char *tok = p;
    bool neg = false;
    if (*p == '-' || *p == '+') neg = (*p++ == '-');
    uint64_t mant = 0;
    int digits = 0, frac = 0;
    bool any = false;
    while (p < end && *p >= '0' && *p <= '9') {
      if (mant || *p != '0') ++digits;
      if (digits <= 19) mant = mant * 10 + (uint64_t)(*p - '0');
      ++p;
      any = true;
    }
    if (p < end && *p == '.') {
      ++p;
      while (p < end && *p >= '0' && *p <= '9') {
        if (mant || *p != '0') ++digits;
        if (digits <= 19) mant = mant * 10 + (uint64_t)(*p - '0');
        ++frac;
        ++p;
        any = true;
      }
    }
    if (!any) return ok = false;
    const bool plain = !(p < end && (*p == 'e' || *p == 'E' || *p == 'n' || *p == 'N' ||
                                     *p == 'i' || *p == 'I'));
    static const double p10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                                 1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    if (plain && sizeof(T) == 8 && digits <= 15 && frac <= 22) {
      double v = (double)mant / p10[frac];
      out = (T)(neg ? -v : v);
      return true;
    }
    if (plain && sizeof(T) == 4 && mant < (1u << 24) && frac <= 10) {
      float v = (float)mant / (float)p10[frac]; // both exact floats: one correctly rounded division
      out = (T)(neg ? -v : v);
      return true;
    }
    // rare: exponents, long mantissas, inf/nan
    char buf[128];
    const char *q = tok;
    while (q < end && !(*q == ' ' || *q == '\n' || *q == '\t' || *q == '\r')) ++q;
    size_t len = (size_t)(q - tok);
    if (len >= sizeof(buf)) return ok = false;
    std::memcpy(buf, tok, len);
    buf[len] = 0;
    char *e = nullptr;
    out = sizeof(T) == 8 ? (T)std::strtod(buf, &e) : (T)std::strtof(buf, &e);
    if (e == buf) return ok = false;
    p = q;
    return true;
  }
};

bool slurp(const char *path, std::vector<char> &buf, std::string &err) {
  FILE *fp = std::fopen(path, "rb");
  if (!fp) {
    err = std::string("cannot open ") + path + ": " + std::strerror(errno);
    return false;
  }
  std::fseek(fp, 0, SEEK_END);
  long sz = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  if (sz < 0) sz = 0;
  buf.resize((size_t)sz);
  size_t got = sz ? std::fread(buf.data(), 1, (size_t)sz, fp) : 0;
  std::fclose(fp);
  if (got != (size_t)sz) {
    err = std::string("short read on ") + path;
    return false;
  }
  return true;
}

template <typename I> bool read_ints(Scanner &s, I *dst, int64_t cnt, int64_t base) {
  for (int64_t i = 0; i < cnt; ++i) {
    int64_t t;
    if (!s.next_int(t)) return false;
    dst[i] = (I)(t - base);
  }
  return true;
}

bool read_vals(Scanner &s, void *dst, int dtype, int64_t cnt) {
  if (dtype == HSPMV_F64) {
    double *d = (double *)dst;
    for (int64_t i = 0; i < cnt; ++i)
      if (!s.next_real(d[i])) return false;
  } else {
    float *d = (float *)dst;
    for (int64_t i = 0; i < cnt; ++i)
      if (!s.next_real(d[i])) return false;
  }
  return true;
}

// appends tokens into a large buffer that is flushed to the FILE in chunks
struct Writer {
  FILE *fp;
  std::vector<char> buf;
  size_t pos = 0;
  explicit Writer(FILE *f) : fp(f), buf(1 << 22) {}
  void room(size_t need) {
    if (pos + need > buf.size()) flush();
  }
  void flush() {
    if (pos) std::fwrite(buf.data(), 1, pos, fp);
    pos = 0;
  }
  void put_uint(uint64_t v) {
    room(24);
    char tmp[24];
    int n = 0;
    do {
      tmp[n++] = (char)('0' + v % 10);
      v /= 10;
    } while (v);
    while (n) buf[pos++] = tmp[--n];
    buf[pos++] = ' ';
  }
  void put_real6(double v) { // "%f " / "%.6f "
    room(400);
    pos += (size_t)std::snprintf(&buf[pos], 390, "%f ", v);
  }
  void put_raw(const char *s) {
    size_t n = std::strlen(s);
    room(n);
    std::memcpy(&buf[pos], s, n);
    pos += n;
  }
};

} // namespace

int read_csr_text(const char *path, int base, int dtype, int64_t *m, int64_t *n, int64_t *nnz,
                  int32_t **rp, int32_t **ci, void **val, std::string &err) {
  if (!path || !m || !n || !nnz || !rp || !ci || !val || (dtype != HSPMV_F32 && dtype != HSPMV_F64))
    return HSPMV_EINVAL;
  std::vector<char> buf;
  if (!slurp(path, buf, err)) return HSPMV_EIO;
  Scanner s{buf.data(), buf.data() + buf.size()};
  if (!s.next_int(*m) || !s.next_int(*n) || !s.next_int(*nnz) || *m < 0 || *nnz < 0) {
    err = std::string("bad header in ") + path;
    return HSPMV_EIO;
  }
  const size_t vs = dtype == HSPMV_F64 ? 8 : 4;
  *rp = (int32_t *)std::malloc(((size_t)*m + 1) * sizeof(int32_t));
  *ci = (int32_t *)std::malloc(((size_t)*nnz + 1) * sizeof(int32_t));
  *val = std::malloc(((size_t)*nnz + 1) * vs);
  if (!*rp || !*ci || !*val) {
    std::free(*rp); std::free(*ci); std::free(*val);
    *rp = *ci = nullptr; *val = nullptr;
    return HSPMV_ENOMEM;
  }
  if (!read_ints(s, *rp, *m + 1, base) || !read_ints(s, *ci, *nnz, base) ||
      !read_vals(s, *val, dtype, *nnz)) {
    err = std::string("truncated or malformed body in ") + path;
    std::free(*rp); std::free(*ci); std::free(*val);
    *rp = *ci = nullptr; *val = nullptr;
    return HSPMV_EIO;
  }
  return HSPMV_OK;
}

int write_csr_text(const char *path, int base, int dtype, int64_t m, int64_t n, int64_t nnz,
                   const int32_t *rp, const int32_t *ci, const void *val, std::string &err) {
  FILE *fp = std::fopen(path, "wb");
  if (!fp) {
    err = std::string("cannot create ") + path + ": " + std::strerror(errno);
    return HSPMV_EIO;
  }
  Writer w(fp);
  char head[96];
  std::snprintf(head, sizeof(head), "%lld %lld %lld\n", (long long)m, (long long)n, (long long)nnz);
  w.put_raw(head);
  for (int64_t i = 0; i <= m; ++i) w.put_uint((uint64_t)(rp[i] + base));
  w.put_raw("\n");
  for (int64_t i = 0; i < nnz; ++i) w.put_uint((uint64_t)(ci[i] + base));
  w.put_raw("\n");
  for (int64_t i = 0; i < nnz; ++i)
    w.put_real6(dtype == HSPMV_F64 ? ((const double *)val)[i] : (double)((const float *)val)[i]);
  w.put_raw("\n");
  w.flush();
  std::fclose(fp);
  return HSPMV_OK;
}

int read_csr3_text(const char *path, int dtype, int64_t *n_ssr, int64_t *n_sr, int64_t *m,
                   int64_t *n, int64_t *nnz, u32 **ssr, u32 **sr, u32 **rp, u32 **ci, void **val,
                   std::string &err) {
  if (!path || !n_ssr || !n_sr || !m || !n || !nnz || !ssr || !sr || !rp || !ci || !val)
    return HSPMV_EINVAL;
  std::vector<char> buf;
  if (!slurp(path, buf, err)) return HSPMV_EIO;
  Scanner s{buf.data(), buf.data() + buf.size()};
  if (!s.next_int(*n_ssr) || !s.next_int(*n_sr) || !s.next_int(*m) || !s.next_int(*n) ||
      !s.next_int(*nnz) || *n_ssr < 0 || *n_sr < 0 || *m < 0 || *nnz < 0) {
    err = std::string("bad .csr3 header in ") + path;
    return HSPMV_EIO;
  }
  const size_t vs = dtype == HSPMV_F64 ? 8 : 4;
  *ssr = (u32 *)std::malloc(((size_t)*n_ssr + 1) * 4);
  *sr = (u32 *)std::malloc(((size_t)*n_sr + 1) * 4);
  *rp = (u32 *)std::malloc(((size_t)*m + 1) * 4);
  *ci = (u32 *)std::malloc(((size_t)*nnz + 1) * 4);
  *val = std::malloc(((size_t)*nnz + 1) * vs);
  bool good = *ssr && *sr && *rp && *ci && *val;
  if (good)
    good = read_ints(s, *ssr, *n_ssr + 1, 0) && read_ints(s, *sr, *n_sr + 1, 0) &&
           read_ints(s, *rp, *m + 1, 0) && read_ints(s, *ci, *nnz, 0) &&
           read_vals(s, *val, dtype, *nnz);
  if (!good) {
    err = std::string("truncated or malformed .csr3 body in ") + path;
    std::free(*ssr); std::free(*sr); std::free(*rp); std::free(*ci); std::free(*val);
    *ssr = *sr = *rp = *ci = nullptr; *val = nullptr;
    return HSPMV_EIO;
  }
  return HSPMV_OK;
}

int write_csr3_text(const char *path, int dtype, int64_t n_ssr, int64_t n_sr, int64_t m,
                    int64_t n, int64_t nnz, const u32 *ssr, const u32 *sr, const u32 *rp,
                    const u32 *ci, const void *val, std::string &err) {
  FILE *fp = std::fopen(path, "wb");
  if (!fp) {
    err = std::string("cannot create ") + path + ": " + std::strerror(errno);
    return HSPMV_EIO;
  }
  Writer w(fp);
  char head[160];
  // "%ld %ld %ld %ld %ld \n" = numCoarsestRows numCoarserRows M N NNZ (spmv-auto.cpp:38-39)
  std::snprintf(head, sizeof(head), "%lld %lld %lld %lld %lld \n", (long long)n_ssr,
                (long long)n_sr, (long long)m, (long long)n, (long long)nnz);
  w.put_raw(head);
  for (int64_t i = 0; i <= n_ssr; ++i) w.put_uint(ssr[i]);
  for (int64_t i = 0; i <= n_sr; ++i) w.put_uint(sr[i]);
  for (int64_t i = 0; i <= m; ++i) w.put_uint(rp[i]);
  for (int64_t i = 0; i < nnz; ++i) w.put_uint(ci[i]);
  for (int64_t i = 0; i < nnz; ++i)
    w.put_real6(dtype == HSPMV_F64 ? ((const double *)val)[i] : (double)((const float *)val)[i]);
  w.flush();
  std::fclose(fp);
  return HSPMV_OK;
}

} // namespace hspmv

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
//
// the file is mmap'ed and parsed by all host threads.  The body is cut into
// one piece per thread at whitespace, a first pass counts the tokens of every piece, the prefix
// sum tells each piece which array slots its tokens belong to, and a second pass converts them in
// place.  The writers format blocks of 2 M entries in parallel and write them in order; "%f" is
// produced by an exact fixed-point fast path (FMA residual decides the rounding, anything within
// 1e-6 of a tie goes through snprintf), so the bytes are those of the reference's fprintf.
#include "host_index.h"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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

inline bool is_ws(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

// Whole file in memory: mmap when the path is a regular file, fread otherwise (pipes, /dev/stdin).
struct FileImage {
  const char *data = nullptr;
  size_t size = 0;
  void *map = nullptr;
  std::vector<char> heap;
  ~FileImage() {
    if (map) ::munmap(map, size);
  }
  bool open(const char *path, std::string &err) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) {
      err = std::string("cannot open ") + path + ": " + std::strerror(errno);
      return false;
    }
    struct stat st;
    if (::fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
      void *m = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        ::madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        map = m;
        data = (const char *)m;
        size = (size_t)st.st_size;
        ::close(fd);
        return true;
      }
    }
    char tmp[1 << 16];
    for (;;) {
      ssize_t got = ::read(fd, tmp, sizeof(tmp));
      if (got < 0) {
        err = std::string("read error on ") + path + ": " + std::strerror(errno);
        ::close(fd);
        return false;
      }
      if (got == 0) break;
      heap.insert(heap.end(), tmp, tmp + got);
    }
    ::close(fd);
    data = heap.data();
    size = heap.size();
    return true;
  }
};

// One array of the file body: `count` tokens converted into dst.
struct Section {
  void *dst;
  int kind; // 0: int32 minus base, 1: uint32, 2: double, 3: float
  int64_t count;
  int64_t base;
};

// Converts the tokens of [p,end) whose global token index starts at `k` into the sections.
bool convert_tokens(const char *p, const char *end, int64_t k, const Section *secs, int nsec, int64_t total) {
  Scanner s{p, end};
  int si = 0;
  int64_t sec_lo = 0;
  while (si < nsec && k >= sec_lo + secs[si].count) sec_lo += secs[si++].count;
  while (k < total) {
    s.skip();
    if (s.p >= end) break;
    const Section &sc = secs[si];
    const int64_t i = k - sec_lo;
    if (sc.kind <= 1) {
      int64_t t;
      if (!s.next_int(t)) return false;
      if (sc.kind == 0)
        ((int32_t *)sc.dst)[i] = (int32_t)(t - sc.base);
      else
        ((u32 *)sc.dst)[i] = (u32)(t - sc.base);
    } else if (sc.kind == 2) {
      if (!s.next_real(((double *)sc.dst)[i])) return false;
    } else {
      if (!s.next_real(((float *)sc.dst)[i])) return false;
    }
    if (s.p < end && !is_ws(*s.p)) return false; // "12abc", a decimal where an index belongs, ...
    ++k;
    while (si < nsec && k >= sec_lo + secs[si].count) sec_lo += secs[si++].count;
  }
  return true;
}

// Parses the body [p,end) -- p sits right behind the header -- into the sections, in parallel.
bool parse_sections(const char *p, const char *end, const Section *secs, int nsec) {
  int64_t total = 0;
  for (int i = 0; i < nsec; ++i) total += secs[i].count;
  const size_t len = (size_t)(end - p);
  int T = (int)std::min<size_t>((size_t)omp_get_max_threads(), len / ((size_t)1 << 20) + 1);
  if (T < 1) T = 1;
  std::vector<const char *> cut((size_t)T + 1);
  cut[0] = p;
  cut[(size_t)T] = end;
  for (int t = 1; t < T; ++t) {
    const char *c = p + len * (size_t)t / (size_t)T;
    if (c < cut[(size_t)t - 1]) c = cut[(size_t)t - 1];
    while (c < end && !is_ws(*c)) ++c; // never split a token: every cut sits on whitespace
    cut[(size_t)t] = c;
  }
  std::vector<int64_t> first((size_t)T + 1, 0);
#pragma omp parallel for num_threads(T) schedule(static, 1)
  for (int t = 0; t < T; ++t) {
    const char *c = cut[(size_t)t], *e = cut[(size_t)t + 1];
    int64_t cnt = 0;
    bool prev_ws = true; // every cut sits on whitespace (or right behind the header)
    for (; c < e; ++c) {
      const bool w = is_ws(*c);
      cnt += (prev_ws && !w);
      prev_ws = w;
    }
    first[(size_t)t + 1] = cnt;
  }
  for (int t = 0; t < T; ++t) first[(size_t)t + 1] += first[(size_t)t];
  if (first[(size_t)T] < total) return false; // truncated
  int bad = 0;
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(| : bad)
  for (int t = 0; t < T; ++t)
    if (first[(size_t)t] < total &&
        !convert_tokens(cut[(size_t)t], cut[(size_t)t + 1], first[(size_t)t], secs, nsec, total))
      bad |= 1;
  return !bad;
}

// ---- writers ----------------------------------------------------------------------------------
inline char *put_uint(char *o, uint64_t v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *o++ = tmp[--n];
  *o++ = ' ';
  return o;
}

// "%f " (= "%.6f "): printf rounds the exact binary value half-to-even at the 6th decimal.
// v * 1e6 = p + e exactly (FMA residual); n = rint(p); the exact scaled value is n + d with
// d = (p - n) + e.  Unless d is within 1e-6 of +-0.5, n is the rounded integer.
inline char *put_real6(char *o, double v) {
  if (std::fabs(v) < 4.0e9) {
    const double p = v * 1e6;
    const double e = std::fma(v, 1e6, -p);
    const double n = std::nearbyint(p);
    const double d = (p - n) + e;
    if (std::fabs(d) <= 0.499999) {
      uint64_t q = (uint64_t)std::fabs(n);
      if (std::signbit(v)) *o++ = '-';
      const uint64_t ip = q / 1000000u;
      uint32_t fp = (uint32_t)(q % 1000000u);
      char tmp[24];
      int k = 0;
      uint64_t t = ip;
      do {
        tmp[k++] = (char)('0' + t % 10);
        t /= 10;
      } while (t);
      while (k) *o++ = tmp[--k];
      *o++ = '.';
      for (int i = 5; i >= 0; --i) {
        o[i] = (char)('0' + fp % 10);
        fp /= 10;
      }
      o += 6;
      *o++ = ' ';
      return o;
    }
  }
  return o + std::snprintf(o, 390, "%f ", v);
}

struct Writer {
  FILE *fp;
  bool ok = true;
  explicit Writer(FILE *f) : fp(f) {}
  void put_raw(const char *s) {
    if (std::fputs(s, fp) < 0) ok = false;
  }
  // formats `cnt` entries through fmt(o, i) -> o' in parallel blocks, written in order
  template <typename F> void put_array(int64_t cnt, size_t max_token, F fmt) {
    const int64_t block = (int64_t)1 << 21;
    const int T = std::max(1, omp_get_max_threads());
    std::vector<std::vector<char>> bufs((size_t)T);
    std::vector<size_t> used((size_t)T);
    for (int64_t b0 = 0; b0 < cnt; b0 += block) {
      const int64_t b1 = std::min(cnt, b0 + block);
      const int64_t per = (b1 - b0 + T - 1) / T;
#pragma omp parallel for num_threads(T) schedule(static, 1)
      for (int t = 0; t < T; ++t) {
        const int64_t lo = std::min(b1, b0 + per * t), hi = std::min(b1, lo + per);
        std::vector<char> &buf = bufs[(size_t)t];
        if (buf.size() < (size_t)per * max_token + 512) buf.resize((size_t)per * max_token + 512);
        char *o = buf.data();
        for (int64_t i = lo; i < hi; ++i) {
          if ((size_t)(o - buf.data()) + 400 > buf.size()) { // a run of huge "%f" values
            const size_t at = (size_t)(o - buf.data());
            buf.resize(buf.size() * 2 + 1024);
            o = buf.data() + at;
          }
          o = fmt(o, i);
        }
        used[(size_t)t] = (size_t)(o - buf.data());
      }
      for (int t = 0; t < T; ++t)
        if (used[(size_t)t] && std::fwrite(bufs[(size_t)t].data(), 1, used[(size_t)t], fp) != used[(size_t)t]) ok = false;
    }
  }
  template <typename I> void put_uints(const I *v, int64_t cnt, int64_t add) {
    put_array(cnt, 12, [=](char *o, int64_t i) { return put_uint(o, (uint64_t)((int64_t)v[i] + add)); });
  }
  void put_reals(const void *v, int dtype, int64_t cnt) {
    if (dtype == HSPMV_F64) {
      const double *d = (const double *)v;
      put_array(cnt, 24, [=](char *o, int64_t i) { return put_real6(o, d[i]); });
    } else {
      const float *d = (const float *)v;
      put_array(cnt, 24, [=](char *o, int64_t i) { return put_real6(o, (double)d[i]); });
    }
  }
};

} // namespace

int read_csr_text(const char *path, int base, int dtype, int64_t *m, int64_t *n, int64_t *nnz,
                  int32_t **rp, int32_t **ci, void **val, std::string &err) {
  if (!path || !m || !n || !nnz || !rp || !ci || !val || (dtype != HSPMV_F32 && dtype != HSPMV_F64))
    return HSPMV_EINVAL;
  FileImage img;
  if (!img.open(path, err)) return HSPMV_EIO;
  Scanner s{img.data, img.data + img.size};
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
  const Section secs[3] = {{*rp, 0, *m + 1, base}, {*ci, 0, *nnz, base}, {*val, dtype == HSPMV_F64 ? 2 : 3, *nnz, 0}};
  if (!parse_sections(s.p, s.end, secs, 3)) {
    err = std::string("truncated or malformed body in ") + path;
    std::free(*rp); std::free(*ci); std::free(*val);
    *rp = *ci = nullptr; *val = nullptr;
    return HSPMV_EIO;
  }
  return HSPMV_OK;
}

// ---- binary side file ------------------------------------------------------------------------
namespace {
struct BinHeader { // 64 bytes
  char magic[8];   // "HSPMVBIN"
  int32_t version, dtype, base, pad;
  int64_t m, n, nnz;
  int64_t text_size, text_mtime_ns;
};

bool text_identity(const char *path, int64_t &size, int64_t &mtime_ns) {
  struct stat st;
  if (::stat(path, &st) != 0) return false;
  size = (int64_t)st.st_size;
  mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000LL + (int64_t)st.st_mtim.tv_nsec;
  return true;
}

bool read_all(int fd, void *dst, size_t bytes) {
  char *p = (char *)dst;
  while (bytes) {
    ssize_t got = ::read(fd, p, bytes < ((size_t)1 << 30) ? bytes : ((size_t)1 << 30));
    if (got <= 0) return false;
    p += got;
    bytes -= (size_t)got;
  }
  return true;
}

bool write_all(int fd, const void *src, size_t bytes) {
  const char *p = (const char *)src;
  while (bytes) {
    ssize_t put = ::write(fd, p, bytes < ((size_t)1 << 30) ? bytes : ((size_t)1 << 30));
    if (put <= 0) return false;
    p += put;
    bytes -= (size_t)put;
  }
  return true;
}
} // namespace

int read_csr_cached(const char *path, int base, int dtype, int64_t *m, int64_t *n, int64_t *nnz,
                    int32_t **rp, int32_t **ci, void **val, int *cache_hit, std::string &err) {
  if (!path || !m || !n || !nnz || !rp || !ci || !val || (dtype != HSPMV_F32 && dtype != HSPMV_F64))
    return HSPMV_EINVAL;
  if (cache_hit) *cache_hit = 0;
  const std::string side = std::string(path) + ".hspmvbin";
  const size_t vs = dtype == HSPMV_F64 ? 8 : 4;
  int64_t tsize = 0, tmtime = 0;
  if (!text_identity(path, tsize, tmtime)) {
    err = std::string("cannot open ") + path + ": " + std::strerror(errno);
    return HSPMV_EIO;
  }
  int fd = ::open(side.c_str(), O_RDONLY);
  if (fd >= 0) {
    BinHeader h;
    bool ok = read_all(fd, &h, sizeof(h)) && !std::memcmp(h.magic, "HSPMVBIN", 8) && h.version == 1 &&
              h.dtype == dtype && h.base == base && h.text_size == tsize && h.text_mtime_ns == tmtime &&
              h.m >= 0 && h.nnz >= 0;
    if (ok) {
      *rp = (int32_t *)std::malloc(((size_t)h.m + 1) * 4);
      *ci = (int32_t *)std::malloc(((size_t)h.nnz + 1) * 4);
      *val = std::malloc(((size_t)h.nnz + 1) * vs);
      ok = *rp && *ci && *val && read_all(fd, *rp, ((size_t)h.m + 1) * 4) &&
           read_all(fd, *ci, (size_t)h.nnz * 4) && read_all(fd, *val, (size_t)h.nnz * vs);
      if (ok) {
        ::close(fd);
        *m = h.m;
        *n = h.n;
        *nnz = h.nnz;
        if (cache_hit) *cache_hit = 1;
        return HSPMV_OK;
      }
      std::free(*rp); std::free(*ci); std::free(*val);
      *rp = *ci = nullptr; *val = nullptr;
    }
    ::close(fd); // stale or damaged: fall through to the text
  }
  int rc = read_csr_text(path, base, dtype, m, n, nnz, rp, ci, val, err);
  if (rc) return rc;
  // best effort: a read-only directory just means no cache
  const std::string tmp = side + ".tmp";
  fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd >= 0) {
    BinHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "HSPMVBIN", 8);
    h.version = 1;
    h.dtype = dtype;
    h.base = base;
    h.m = *m;
    h.n = *n;
    h.nnz = *nnz;
    h.text_size = tsize;
    h.text_mtime_ns = tmtime;
    const bool ok = write_all(fd, &h, sizeof(h)) && write_all(fd, *rp, ((size_t)*m + 1) * 4) &&
                    write_all(fd, *ci, (size_t)*nnz * 4) && write_all(fd, *val, (size_t)*nnz * vs);
    ::close(fd);
    if (!ok || ::rename(tmp.c_str(), side.c_str()) != 0) ::unlink(tmp.c_str());
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
  w.put_uints(rp, m + 1, base);
  w.put_raw("\n");
  w.put_uints(ci, nnz, base);
  w.put_raw("\n");
  w.put_reals(val, dtype, nnz);
  w.put_raw("\n");
  if (std::fclose(fp) != 0 || !w.ok) {
    err = std::string("write error on ") + path;
    return HSPMV_EIO;
  }
  return HSPMV_OK;
}

int read_csr3_text(const char *path, int dtype, int64_t *n_ssr, int64_t *n_sr, int64_t *m,
                   int64_t *n, int64_t *nnz, u32 **ssr, u32 **sr, u32 **rp, u32 **ci, void **val,
                   std::string &err) {
  if (!path || !n_ssr || !n_sr || !m || !n || !nnz || !ssr || !sr || !rp || !ci || !val)
    return HSPMV_EINVAL;
  FileImage img;
  if (!img.open(path, err)) return HSPMV_EIO;
  Scanner s{img.data, img.data + img.size};
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
  if (good) {
    const Section secs[5] = {{*ssr, 1, *n_ssr + 1, 0}, {*sr, 1, *n_sr + 1, 0}, {*rp, 1, *m + 1, 0},
                             {*ci, 1, *nnz, 0}, {*val, dtype == HSPMV_F64 ? 2 : 3, *nnz, 0}};
    good = parse_sections(s.p, s.end, secs, 5);
  }
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
  w.put_uints(ssr, n_ssr + 1, 0);
  w.put_uints(sr, n_sr + 1, 0);
  w.put_uints(rp, m + 1, 0);
  w.put_uints(ci, nnz, 0);
  w.put_reals(val, dtype, nnz);
  if (std::fclose(fp) != 0 || !w.ok) {
    err = std::string("write error on ") + path;
    return HSPMV_EIO;
  }
  return HSPMV_OK;
}

} // namespace hspmv

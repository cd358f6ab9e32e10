// Host-side text formatting of TREC run files (multi-threaded C++, part of libcmx.so).
//
// Replaces the per-line Python f-string loops of the reference run scripts, which become
// the dominant cost once the search itself takes a fraction of a second:
//   mono      onepass_dense_mix_run_custom_lang.py:879-888
//   bilingual onepass_bilingual_mix_hub_custom_lang.py:950-958 (raw) and :165-181 (collapse_run_max)
// Output bytes are identical to what those loops write for the same (D, I):
//   * f"{score:.4f}" / f"{sc:.6f}" of a float32: float32 * 10^decimals is exact in double,
//     so round-half-even on it (nearbyint) reproduces Python's correctly rounded formatting;
//   * collapse: max over the 6-decimal ROUNDED scores, first-seen order of bases, stable sort.
// This is byte/integer work on the CPU by nature (the text is 3x larger than the (D, I)
// arrays it is made from, so formatting on the GPU would only add PCIe traffic).
//
// Layout of the work: queries are split over threads; each thread formats its range into its
// own growable buffer with raw pointer writes (one capacity check per line, not per
// character) and prefetches the document-id strings of the hits a few lines ahead -- at
// 8.8 M documents the id table is ~130 MB and every hit is a cache miss.  The parts are then
// either copied (in parallel) into one malloc'ed buffer, or written straight into the run
// file with one pwrite per thread at its final offset (no intermediate copy at all).
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

#include "../../include/cmx.h"

namespace cmx {
void set_error(const char* fmt, ...);
}

namespace {

struct StrTable {
  const char* buf;
  const int64_t* off;  // n + 1 offsets
  int64_t n;
  inline int64_t len(int64_t i) const { return off[i + 1] - off[i]; }
  inline const char* ptr(int64_t i) const { return buf + off[i]; }
};

// growable output of one thread
struct Part {
  char* base = nullptr;
  size_t cap = 0, len = 0;
  bool failed = false;
  ~Part() { free(base); }
  Part() = default;
  Part(const Part&) = delete;
  Part& operator=(const Part&) = delete;
  // room for at least `extra` more bytes; returns the write cursor (nullptr when out of memory)
  inline char* need(size_t extra) {
    if (len + extra > cap) {
      size_t ncap = std::max(cap + cap / 2, len + extra + (1u << 16));
      char* nb = (char*)realloc(base, ncap);
      if (!nb) { failed = true; return nullptr; }
      base = nb;
      cap = ncap;
    }
    return base + len;
  }
  inline void done(char* cursor) { len = (size_t)(cursor - base); }
  // same with a live cursor: keeps `cursor` valid across a reallocation
  inline bool ensure(char*& cursor, size_t extra) {
    if ((size_t)(cursor - base) + extra <= cap) return true;
    done(cursor);
    cursor = need(extra);
    return cursor != nullptr;
  }
};

inline char* put_str(char* p, const char* s, size_t n) { memcpy(p, s, n); return p + n; }

inline char* put_uint(char* p, uint64_t u) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  while (n) *p++ = tmp[--n];
  return p;
}

inline char* put_int(char* p, int64_t v) {
  if (v < 0) { *p++ = '-'; return put_uint(p, (uint64_t)(-(v + 1)) + 1u); }
  return put_uint(p, (uint64_t)v);
}

constexpr int kMaxFixed = 64;  // longest f"{float32:.6f}": '-' + 39 digits + '.' + 6
// f"{float(x):.{dec}f}" for a float32 x (dec = 4 or 6)
inline char* put_fixed(char* p, float x, int dec) {
  const double scale = dec == 4 ? 1e4 : 1e6;
  const double ax = std::fabs((double)x);
  if (!std::isfinite(x) || ax >= 1e11) {
    if (std::isnan(x)) return put_str(p, "nan", 3);
    return p + snprintf(p, kMaxFixed, dec == 4 ? "%.4f" : "%.6f", (double)x);
  }
  const uint64_t n = (uint64_t)std::nearbyint(ax * scale);  // exact product, ties to even
  const uint64_t sc = (uint64_t)scale;
  if (std::signbit(x)) *p++ = '-';
  p = put_uint(p, n / sc);
  *p++ = '.';
  uint64_t f = n % sc;
  for (int i = dec - 1; i >= 0; --i) { p[i] = (char)('0' + f % 10); f /= 10; }
  return p + dec;
}

template <typename F>
void parallel_ranges(int64_t n, int nthreads, F fn) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = (int)std::max<int64_t>(1, n);
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) {
    const int64_t a = n * t / nthreads, b = n * (t + 1) / nthreads;
    th.emplace_back([=]() { fn(t, a, b); });
  }
  fn(0, 0, n / nthreads);  // the caller is thread 0
  for (auto& x : th) x.join();
}

int pick_threads(int nthreads, int64_t nq) {
  if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
  return (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, std::max<int64_t>(nq, 1)));
}

bool any_failed(const std::vector<Part>& parts) {
  for (auto& p : parts) if (p.failed) return true;
  return false;
}

// parts -> one malloc'ed buffer (parallel copy)
int gather_parts(const std::vector<Part>& parts, char** out, int64_t* out_len, const char* who) {
  if (any_failed(parts)) { cmx::set_error("%s: out of host memory", who); return CMX_ERR_NOMEM; }
  std::vector<size_t> at(parts.size() + 1, 0);
  for (size_t i = 0; i < parts.size(); ++i) at[i + 1] = at[i] + parts[i].len;
  char* buf = (char*)malloc(at.back() ? at.back() : 1);
  if (!buf) { cmx::set_error("%s: out of host memory", who); return CMX_ERR_NOMEM; }
  parallel_ranges((int64_t)parts.size(), (int)parts.size(), [&](int, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) memcpy(buf + at[(size_t)i], parts[(size_t)i].base, parts[(size_t)i].len);
  });
  *out = buf;
  *out_len = (int64_t)at.back();
  return CMX_OK;
}

// parts -> file (created / truncated), one pwrite stream per part at its final offset
int write_parts(const std::vector<Part>& parts, const char* path, int64_t* out_len, const char* who) {
  if (any_failed(parts)) { cmx::set_error("%s: out of host memory", who); return CMX_ERR_NOMEM; }
  std::vector<size_t> at(parts.size() + 1, 0);
  for (size_t i = 0; i < parts.size(); ++i) at[i + 1] = at[i] + parts[i].len;
  const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) { cmx::set_error("%s: cannot open %s: %s", who, path, strerror(errno)); return CMX_ERR_INVALID; }
  std::vector<int> err(parts.size(), 0);
  parallel_ranges((int64_t)parts.size(), (int)parts.size(), [&](int, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      const Part& p = parts[(size_t)i];
      size_t done = 0;
      while (done < p.len) {
        const ssize_t w = pwrite(fd, p.base + done, p.len - done, (off_t)(at[(size_t)i] + done));
        if (w < 0) { if (errno == EINTR) continue; err[(size_t)i] = errno; break; }
        done += (size_t)w;
      }
    }
  });
  const int cerr = close(fd) != 0 ? errno : 0;
  for (int e : err) if (e) { cmx::set_error("%s: write to %s failed: %s", who, path, strerror(e)); return CMX_ERR_INVALID; }
  if (cerr) { cmx::set_error("%s: close of %s failed: %s", who, path, strerror(cerr)); return CMX_ERR_INVALID; }
  *out_len = (int64_t)at.back();
  return CMX_OK;
}

constexpr int kAhead = 12;  // lines of look-ahead for the id-table prefetches

struct MonoArgs {
  const float* D; const int64_t* I; int64_t nq; int k;
  StrTable q, dt; const int64_t* doc_keys; bool have_docs; const char* tag;
};

inline int64_t mono_doc_pos(const MonoArgs& a, int64_t id) {
  if (!a.have_docs) return -1;
  if (a.doc_keys) {
    const int64_t* e = a.doc_keys + a.dt.n;
    const int64_t* it = std::lower_bound(a.doc_keys, e, id);
    return (it != e && *it == id) ? it - a.doc_keys : -1;
  }
  return (id >= 0 && id < a.dt.n) ? id : -1;
}

void format_mono(const MonoArgs& a, int nthreads, std::vector<Part>& parts) {
  const size_t tag_len = strlen(a.tag);
  parallel_ranges(a.nq, nthreads, [&](int t, int64_t q0, int64_t q1) {
    Part& out = parts[(size_t)t];
    const bool direct = a.have_docs && !a.doc_keys;  // id == position: prefetchable
    char* p = out.need((size_t)(q1 - q0) * (size_t)a.k * 44 + 64);  // typical line: 40-44 bytes
    if (!p) return;
    for (int64_t r = q0; r < q1; ++r) {
      const size_t qlen = (size_t)a.q.len(r);
      const size_t fixed = qlen + 4 + 1 + 12 + 1 + kMaxFixed + 1 + tag_len + 1;
      const int64_t* Ir = a.I + r * a.k;
      const float* Dr = a.D + r * a.k;
      const char* qs = a.q.ptr(r);
      for (int j = 0; j < a.k; ++j) {
        if (direct) {
          // two-level prefetch: the offset entry of a far hit, the id bytes of a nearer one
          if (j + kAhead < a.k) { const int64_t f = Ir[j + kAhead]; if (f >= 0 && f < a.dt.n) __builtin_prefetch(a.dt.off + f); }
          if (j + kAhead / 2 < a.k) { const int64_t f = Ir[j + kAhead / 2]; if (f >= 0 && f < a.dt.n) __builtin_prefetch(a.dt.buf + a.dt.off[f]); }
        }
        const int64_t id = Ir[j];
        const int64_t pos = mono_doc_pos(a, id);
        const size_t dlen = pos >= 0 ? (size_t)a.dt.len(pos) : 21;
        if (!out.ensure(p, fixed + dlen)) return;
        if (r != 0 || j != 0) *p++ = '\n';  // "\n".join(lines): no trailing newline
        p = put_str(p, qs, qlen);
        p = put_str(p, "\tQ0\t", 4);
        if (pos >= 0) p = put_str(p, a.dt.ptr(pos), dlen); else p = put_int(p, id);  // id_lookup.get(int(doc), str(doc))
        *p++ = '\t';
        p = put_uint(p, (uint64_t)(j + 1));
        *p++ = '\t';
        p = put_fixed(p, Dr[j], 4);
        *p++ = '\t';
        p = put_str(p, a.tag, tag_len);
      }
    }
    out.done(p);
  });
}

struct BiArgs {
  const float* D; const int64_t* I; int64_t nq; int k;
  StrTable q, dt, bt; const int32_t* base_code; const char* tag;
  // collapsed groups computed on the device (cmx_collapse_max), or NULL: group here
  const int32_t* pre_code = nullptr; const int64_t* pre_val6 = nullptr; const int32_t* pre_count = nullptr;
};

void format_bilingual(const BiArgs& a, int nthreads, std::vector<Part>& raw, std::vector<Part>& col) {
  const size_t tag_len = strlen(a.tag);
  parallel_ranges(a.nq, nthreads, [&](int t, int64_t q0, int64_t q1) {
    Part& rs = raw[(size_t)t];
    Part& cs = col[(size_t)t];
    char* p = rs.need((size_t)(q1 - q0) * (size_t)a.k * 52 + 64);
    char* c = cs.need((size_t)(q1 - q0) * (size_t)a.k * 40 + 64);
    if (!p || !c) return;
    // The collapsed run carries, per base id, the max over the scores AS THE RAW FILE PRINTS THEM.
    // |score| < 1e11: fixed point, score6 = |rounded score| * 1e6 (exact).  Larger finite float32 values
    // are integers, print as "<digits>.000000" and survive the text round trip unchanged: kept as is.
    struct Grp { int32_t code; int64_t score6; bool neg; bool big; float orig; };
    auto cmp = [](const Grp& x, const Grp& y) -> int {  // -1 / 0 / +1 like the parsed doubles compare
      if (!x.big && !y.big) {
        const int64_t vx = x.neg ? -x.score6 : x.score6, vy = y.neg ? -y.score6 : y.score6;
        return (vx > vy) - (vx < vy);
      }
      const double dx = x.big ? (double)x.orig : (x.neg ? -1.0 : 1.0) * (double)x.score6 / 1e6;
      const double dy = y.big ? (double)y.orig : (y.neg ? -1.0 : 1.0) * (double)y.score6 / 1e6;
      return (dx > dy) - (dx < dy);
    };
    std::vector<Grp> groups;
    // base code -> group slot: open addressing over thread-local arrays, emptied per query by bumping a stamp --
    // no allocation inside the loop (a std::unordered_map here made 8 threads slower than 1: one malloc per hit)
    size_t tab_size = 64;
    while (tab_size < 2 * (size_t)a.k) tab_size <<= 1;
    const size_t tab_mask = tab_size - 1;
    std::vector<int32_t> tab_code(tab_size);
    std::vector<int> tab_slot(tab_size);
    std::vector<uint32_t> tab_stamp(tab_size, 0u);
    uint32_t stamp = 0;
    std::vector<int> order;
    groups.reserve((size_t)a.k);
    order.reserve((size_t)a.k);
    for (int64_t r = q0; r < q1; ++r) {
      const size_t qlen = (size_t)a.q.len(r);
      const size_t fixed = qlen + 4 + 1 + 12 + 1 + kMaxFixed + 1 + tag_len + 1;
      const int64_t* Ir = a.I + r * a.k;
      const float* Dr = a.D + r * a.k;
      const char* qs = a.q.ptr(r);
      groups.clear();
      if (++stamp == 0u) { std::fill(tab_stamp.begin(), tab_stamp.end(), 0u); stamp = 1; }
      for (int j = 0; j < a.k; ++j) {
        if (j + kAhead < a.k) { const int64_t f = Ir[j + kAhead]; if (f >= 0 && f < a.dt.n) { __builtin_prefetch(a.dt.off + f); __builtin_prefetch(a.base_code + f); } }
        if (j + kAhead / 2 < a.k) { const int64_t f = Ir[j + kAhead / 2]; if (f >= 0 && f < a.dt.n) __builtin_prefetch(a.dt.buf + a.dt.off[f]); }
        const int64_t ix = Ir[j];
        if (ix < 0 || ix >= a.dt.n) continue;  // skipped hits keep their rank number
        const float sc = Dr[j];
        const size_t dlen = (size_t)a.dt.len(ix);
        if (!rs.ensure(p, fixed + dlen)) return;
        p = put_str(p, qs, qlen);
        p = put_str(p, " Q0 ", 4);
        p = put_str(p, a.dt.ptr(ix), dlen);
        *p++ = ' ';
        p = put_uint(p, (uint64_t)(j + 1));
        *p++ = ' ';
        p = put_fixed(p, sc, 6);
        *p++ = ' ';
        p = put_str(p, a.tag, tag_len);
        *p++ = '\n';
        if (a.pre_code) continue;  // the groups of this query come from the device
        // collapse on the value the raw file carries: the 6-decimal rounded score
        const double ax = std::fabs((double)sc);
        const bool small = std::isfinite(sc) && ax < 1e11;
        const bool neg = std::signbit(sc);  // the sign survives the text round trip even for -0.000000
        const int32_t code = a.base_code[ix];
        const Grp cand{code, small ? (int64_t)std::nearbyint(ax * 1e6) : 0, neg, !small, sc};
        size_t h = ((uint32_t)code * 2654435761u) & tab_mask;
        while (tab_stamp[h] == stamp && tab_code[h] != code) h = (h + 1) & tab_mask;
        if (tab_stamp[h] != stamp) {
          tab_stamp[h] = stamp;
          tab_code[h] = code;
          tab_slot[h] = (int)groups.size();
          groups.push_back(cand);
        } else {
          Grp& g = groups[(size_t)tab_slot[h]];
          if (cmp(cand, g) > 0) g = cand;
        }
      }
      if (a.pre_code) {
        const int cnt = a.pre_count[r];
        for (int gi = 0; gi < cnt; ++gi) {
          const int32_t code = a.pre_code[r * a.k + gi];
          const int64_t v6 = a.pre_val6[r * a.k + gi];
          groups.push_back(Grp{code, v6 < 0 ? -v6 : v6, v6 < 0, false, 0.f});
        }
      }
      order.resize(groups.size());
      for (size_t i = 0; i < groups.size(); ++i) order[i] = (int)i;
      if (!a.pre_code)
      std::stable_sort(order.begin(), order.end(),
                       [&](int x, int y) { return cmp(groups[(size_t)x], groups[(size_t)y]) > 0; });
      uint64_t rank = 1;
      for (int gi : order) {
        const Grp& g = groups[(size_t)gi];
        const size_t blen = (size_t)a.bt.len(g.code);
        if (!cs.ensure(c, qlen + 4 + blen + 1 + 12 + 1 + kMaxFixed + 15)) return;
        c = put_str(c, qs, qlen);
        c = put_str(c, " Q0 ", 4);
        c = put_str(c, a.bt.ptr(g.code), blen);
        *c++ = ' ';
        c = put_uint(c, rank++);
        *c++ = ' ';
        if (g.big) {
          c = put_fixed(c, g.orig, 6);
        } else {
          if (g.neg) *c++ = '-';
          c = put_uint(c, (uint64_t)(g.score6 / 1000000));
          *c++ = '.';
          int64_t f = g.score6 % 1000000;
          for (int i = 5; i >= 0; --i) { c[i] = (char)('0' + f % 10); f /= 10; }
          c += 6;
        }
        c = put_str(c, " bilingual-mix\n", 15);
      }
    }
    rs.done(p);
    cs.done(c);
  });
}

bool mono_args_ok(const float* D, const int64_t* I, int64_t nq, int k, const char* qids, const int64_t* qid_off, const char* tag) {
  return D && I && qids && qid_off && tag && nq >= 0 && k >= 1;
}

}  // namespace

extern "C" {

void cmx_free_text(char* p) { free(p); }

int cmx_trec_mono(const float* D, const int64_t* I, int64_t nq, int k, const char* qids, const int64_t* qid_off,
                  const char* docs, const int64_t* doc_off, const int64_t* doc_keys, int64_t ndocs, const char* tag,
                  int nthreads, char** out, int64_t* out_len) {
  if (!mono_args_ok(D, I, nq, k, qids, qid_off, tag) || !out || !out_len) {
    cmx::set_error("cmx_trec_mono: bad argument");
    return CMX_ERR_INVALID;
  }
  const MonoArgs a{D, I, nq, k, {qids, qid_off, nq}, {docs, doc_off, docs ? ndocs : 0}, doc_keys, docs != nullptr, tag};
  const int nt = pick_threads(nthreads, nq);
  std::vector<Part> parts((size_t)nt);
  format_mono(a, nt, parts);
  return gather_parts(parts, out, out_len, "cmx_trec_mono");
}

int cmx_trec_mono_file(const float* D, const int64_t* I, int64_t nq, int k, const char* qids, const int64_t* qid_off,
                       const char* docs, const int64_t* doc_off, const int64_t* doc_keys, int64_t ndocs,
                       const char* tag, int nthreads, const char* path, int64_t* out_len) {
  if (!mono_args_ok(D, I, nq, k, qids, qid_off, tag) || !path || !out_len) {
    cmx::set_error("cmx_trec_mono_file: bad argument");
    return CMX_ERR_INVALID;
  }
  const MonoArgs a{D, I, nq, k, {qids, qid_off, nq}, {docs, doc_off, docs ? ndocs : 0}, doc_keys, docs != nullptr, tag};
  const int nt = pick_threads(nthreads, nq);
  std::vector<Part> parts((size_t)nt);
  format_mono(a, nt, parts);
  return write_parts(parts, path, out_len, "cmx_trec_mono_file");
}

static bool bi_args_ok(const float* D, const int64_t* I, int64_t nq, int k, const char* qids, const int64_t* qid_off,
                       const char* docs, const int64_t* doc_off, const int32_t* base_code, const char* bases,
                       const int64_t* base_off, const char* tag) {
  return D && I && qids && qid_off && docs && doc_off && base_code && bases && base_off && tag && nq >= 0 && k >= 1;
}

int cmx_trec_bilingual(const float* D, const int64_t* I, int64_t nq, int k, const char* qids, const int64_t* qid_off,
                       const char* docs, const int64_t* doc_off, int64_t ndocs, const int32_t* base_code,
                       const char* bases, const int64_t* base_off, int64_t nbases, const char* tag, int nthreads,
                       char** raw_out, int64_t* raw_len, char** col_out, int64_t* col_len) {
  if (!bi_args_ok(D, I, nq, k, qids, qid_off, docs, doc_off, base_code, bases, base_off, tag) || !raw_out || !raw_len ||
      !col_out || !col_len) {
    cmx::set_error("cmx_trec_bilingual: bad argument");
    return CMX_ERR_INVALID;
  }
  const BiArgs a{D, I, nq, k, {qids, qid_off, nq}, {docs, doc_off, ndocs}, {bases, base_off, nbases}, base_code, tag};
  const int nt = pick_threads(nthreads, nq);
  std::vector<Part> raw((size_t)nt), col((size_t)nt);
  format_bilingual(a, nt, raw, col);
  *raw_out = *col_out = nullptr;
  int rc = gather_parts(raw, raw_out, raw_len, "cmx_trec_bilingual");
  if (rc == CMX_OK) rc = gather_parts(col, col_out, col_len, "cmx_trec_bilingual");
  if (rc != CMX_OK) { free(*raw_out); *raw_out = nullptr; }
  return rc;
}

int cmx_trec_bilingual_file(const float* D, const int64_t* I, int64_t nq, int k, const char* qids,
                            const int64_t* qid_off, const char* docs, const int64_t* doc_off, int64_t ndocs,
                            const int32_t* base_code, const char* bases, const int64_t* base_off, int64_t nbases,
                            const char* tag, int nthreads, const char* raw_path, const char* col_path,
                            int64_t* raw_len, int64_t* col_len) {
  if (!bi_args_ok(D, I, nq, k, qids, qid_off, docs, doc_off, base_code, bases, base_off, tag) || !raw_path || !col_path ||
      !raw_len || !col_len) {
    cmx::set_error("cmx_trec_bilingual_file: bad argument");
    return CMX_ERR_INVALID;
  }
  const BiArgs a{D, I, nq, k, {qids, qid_off, nq}, {docs, doc_off, ndocs}, {bases, base_off, nbases}, base_code, tag};
  const int nt = pick_threads(nthreads, nq);
  std::vector<Part> raw((size_t)nt), col((size_t)nt);
  format_bilingual(a, nt, raw, col);
  int rc = write_parts(raw, raw_path, raw_len, "cmx_trec_bilingual_file");
  if (rc == CMX_OK) rc = write_parts(col, col_path, col_len, "cmx_trec_bilingual_file");
  return rc;
}

int cmx_trec_bilingual_file_pre(const float* D, const int64_t* I, int64_t nq, int k, const char* qids,
                                const int64_t* qid_off, const char* docs, const int64_t* doc_off, int64_t ndocs,
                                const int32_t* base_code, const char* bases, const int64_t* base_off, int64_t nbases,
                                const int32_t* col_code, const int64_t* col_val6, const int32_t* col_count,
                                const char* tag, int nthreads, const char* raw_path, const char* col_path,
                                int64_t* raw_len, int64_t* col_len) {
  if (!bi_args_ok(D, I, nq, k, qids, qid_off, docs, doc_off, base_code, bases, base_off, tag) || !raw_path || !col_path ||
      !raw_len || !col_len || !col_code || !col_val6 || !col_count) {
    cmx::set_error("cmx_trec_bilingual_file_pre: bad argument");
    return CMX_ERR_INVALID;
  }
  for (int64_t r = 0; r < nq; ++r) {
    if (col_count[r] < 0 || col_count[r] > k) { cmx::set_error("cmx_trec_bilingual_file_pre: bad group count"); return CMX_ERR_INVALID; }
    for (int g = 0; g < col_count[r]; ++g)
      if (col_code[r * k + g] < 0 || col_code[r * k + g] >= nbases) { cmx::set_error("cmx_trec_bilingual_file_pre: bad base code"); return CMX_ERR_INVALID; }
  }
  BiArgs a{D, I, nq, k, {qids, qid_off, nq}, {docs, doc_off, ndocs}, {bases, base_off, nbases}, base_code, tag};
  a.pre_code = col_code;
  a.pre_val6 = col_val6;
  a.pre_count = col_count;
  const int nt = pick_threads(nthreads, nq);
  std::vector<Part> raw((size_t)nt), col((size_t)nt);
  format_bilingual(a, nt, raw, col);
  int rc = write_parts(raw, raw_path, raw_len, "cmx_trec_bilingual_file_pre");
  if (rc == CMX_OK) rc = write_parts(col, col_path, col_len, "cmx_trec_bilingual_file_pre");
  return rc;
}

}  // extern "C"

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
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
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
  inline void append(std::string& s, int64_t i) const { s.append(buf + off[i], (size_t)(off[i + 1] - off[i])); }
};

inline void append_int(std::string& s, int64_t v) {
  char tmp[24];
  int n = 0;
  bool neg = v < 0;
  uint64_t u = neg ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  if (neg) s.push_back('-');
  while (n) s.push_back(tmp[--n]);
}

// f"{float(x):.{dec}f}" for a float32 x (dec = 4 or 6)
inline void append_fixed(std::string& s, float x, int dec) {
  const double scale = dec == 4 ? 1e4 : 1e6;
  const double ax = std::fabs((double)x);
  if (!std::isfinite(x) || ax >= 1e11) {
    char tmp[96];
    if (std::isnan(x)) { s.append("nan"); return; }
    snprintf(tmp, sizeof(tmp), dec == 4 ? "%.4f" : "%.6f", (double)x);
    s.append(tmp);
    return;
  }
  const uint64_t n = (uint64_t)std::nearbyint(ax * scale);  // exact product, ties to even
  const uint64_t sc = (uint64_t)scale;
  if (std::signbit(x)) s.push_back('-');
  append_int(s, (int64_t)(n / sc));
  s.push_back('.');
  uint64_t f = n % sc;
  char tmp[8];
  for (int i = dec - 1; i >= 0; --i) { tmp[i] = (char)('0' + f % 10); f /= 10; }
  s.append(tmp, (size_t)dec);
}

template <typename F>
void parallel_ranges(int64_t n, int nthreads, F fn) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = (int)std::max<int64_t>(1, n);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    const int64_t a = n * t / nthreads, b = n * (t + 1) / nthreads;
    th.emplace_back([=]() { fn(t, a, b); });
  }
  for (auto& x : th) x.join();
}

char* join_parts(const std::vector<std::string>& parts, int64_t* out_len) {
  size_t total = 0;
  for (auto& p : parts) total += p.size();
  char* out = (char*)malloc(total ? total : 1);
  if (!out) return nullptr;
  size_t pos = 0;
  for (auto& p : parts) { memcpy(out + pos, p.data(), p.size()); pos += p.size(); }
  *out_len = (int64_t)total;
  return out;
}

}  // namespace

extern "C" {

void cmx_free_text(char* p) { free(p); }

int cmx_trec_mono(const float* D, const int64_t* I, int64_t nq, int k, const char* qids, const int64_t* qid_off,
                  const char* docs, const int64_t* doc_off, const int64_t* doc_keys, int64_t ndocs, const char* tag,
                  int nthreads, char** out, int64_t* out_len) {
  if (!D || !I || !qids || !qid_off || !tag || !out || !out_len || nq < 0 || k < 1) {
    cmx::set_error("cmx_trec_mono: bad argument");
    return CMX_ERR_INVALID;
  }
  const StrTable q{qids, qid_off, nq}, dt{docs, doc_off, ndocs};
  const std::string tail = std::string("\t") + tag;
  if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::string> parts((size_t)std::max<int64_t>(1, std::min<int64_t>(nthreads, std::max<int64_t>(nq, 1))));
  parallel_ranges(nq, (int)parts.size(), [&](int t, int64_t a, int64_t b) {
    std::string& s = parts[(size_t)t];
    s.reserve((size_t)((b - a) * k * 44));
    for (int64_t r = a; r < b; ++r) {
      for (int j = 0; j < k; ++j) {
        if (r != 0 || j != 0) s.push_back('\n');  // "\n".join(lines): no trailing newline
        q.append(s, r);
        s.append("\tQ0\t");
        const int64_t id = I[r * k + j];
        int64_t pos = -1;
        if (docs) {
          if (doc_keys) {
            const int64_t* e = doc_keys + ndocs;
            const int64_t* it = std::lower_bound(doc_keys, e, id);
            if (it != e && *it == id) pos = it - doc_keys;
          } else if (id >= 0 && id < ndocs) {
            pos = id;
          }
        }
        if (pos >= 0) dt.append(s, pos); else append_int(s, id);  // id_lookup.get(int(doc), str(doc))
        s.push_back('\t');
        append_int(s, j + 1);
        s.push_back('\t');
        append_fixed(s, D[r * k + j], 4);
        s.append(tail);
      }
    }
  });
  *out = join_parts(parts, out_len);
  if (!*out) { cmx::set_error("cmx_trec_mono: out of host memory"); return CMX_ERR_NOMEM; }
  return CMX_OK;
}

int cmx_trec_bilingual(const float* D, const int64_t* I, int64_t nq, int k, const char* qids, const int64_t* qid_off,
                       const char* docs, const int64_t* doc_off, int64_t ndocs, const int32_t* base_code,
                       const char* bases, const int64_t* base_off, int64_t nbases, const char* tag, int nthreads,
                       char** raw_out, int64_t* raw_len, char** col_out, int64_t* col_len) {
  if (!D || !I || !qids || !qid_off || !docs || !doc_off || !base_code || !bases || !base_off || !tag || !raw_out ||
      !raw_len || !col_out || !col_len || nq < 0 || k < 1) {
    cmx::set_error("cmx_trec_bilingual: bad argument");
    return CMX_ERR_INVALID;
  }
  const StrTable q{qids, qid_off, nq}, dt{docs, doc_off, ndocs}, bt{bases, base_off, nbases};
  const std::string tail = std::string(" ") + tag + "\n";
  if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
  const size_t np = (size_t)std::max<int64_t>(1, std::min<int64_t>(nthreads, std::max<int64_t>(nq, 1)));
  std::vector<std::string> raw(np), col(np);
  parallel_ranges(nq, (int)np, [&](int t, int64_t a, int64_t b) {
    std::string& rs = raw[(size_t)t];
    std::string& cs = col[(size_t)t];
    rs.reserve((size_t)((b - a) * k * 52));
    cs.reserve((size_t)((b - a) * k * 40));
    struct Grp { int32_t code; int64_t score6; bool neg; };  // score6 = |rounded score| * 1e6
    std::vector<Grp> groups;
    std::unordered_map<int32_t, int> slot;
    std::vector<int> order;
    for (int64_t r = a; r < b; ++r) {
      groups.clear();
      slot.clear();
      for (int j = 0; j < k; ++j) {
        const int64_t ix = I[r * k + j];
        if (ix < 0 || ix >= ndocs) continue;  // skipped hits keep their rank number
        const float sc = D[r * k + j];
        q.append(rs, r);
        rs.append(" Q0 ");
        dt.append(rs, ix);
        rs.push_back(' ');
        append_int(rs, j + 1);
        rs.push_back(' ');
        append_fixed(rs, sc, 6);
        rs.append(tail);
        // collapse on the value the raw file carries: the 6-decimal rounded score
        const double ax = std::fabs((double)sc);
        const bool fin = std::isfinite(sc) && ax < 1e11;
        const int64_t v6 = fin ? (int64_t)std::nearbyint(ax * 1e6) : (int64_t)9e18;
        const bool neg = std::signbit(sc);  // the sign survives the text round trip even for -0.000000
        const int32_t code = base_code[ix];
        auto it = slot.find(code);
        if (it == slot.end()) {
          slot.emplace(code, (int)groups.size());
          groups.push_back({code, v6, neg});
        } else {
          Grp& g = groups[(size_t)it->second];
          const int64_t cur = g.neg ? -g.score6 : g.score6, nv = neg ? -v6 : v6;
          if (nv > cur) { g.score6 = v6; g.neg = neg; }
        }
      }
      order.resize(groups.size());
      for (size_t i = 0; i < groups.size(); ++i) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const int64_t vx = groups[(size_t)x].neg ? -groups[(size_t)x].score6 : groups[(size_t)x].score6;
        const int64_t vy = groups[(size_t)y].neg ? -groups[(size_t)y].score6 : groups[(size_t)y].score6;
        return vx > vy;
      });
      int rank = 1;
      for (int gi : order) {
        const Grp& g = groups[(size_t)gi];
        q.append(cs, r);
        cs.append(" Q0 ");
        bt.append(cs, g.code);
        cs.push_back(' ');
        append_int(cs, rank++);
        cs.push_back(' ');
        if (g.neg) cs.push_back('-');
        append_int(cs, g.score6 / 1000000);
        cs.push_back('.');
        char tmp[8];
        int64_t f = g.score6 % 1000000;
        for (int i = 5; i >= 0; --i) { tmp[i] = (char)('0' + f % 10); f /= 10; }
        cs.append(tmp, 6);
        cs.append(" bilingual-mix\n");
      }
    }
  });
  *raw_out = join_parts(raw, raw_len);
  *col_out = join_parts(col, col_len);
  if (!*raw_out || !*col_out) { cmx::set_error("cmx_trec_bilingual: out of host memory"); return CMX_ERR_NOMEM; }
  return CMX_OK;
}

}  // extern "C"

// C ABI of libcmx.so (see include/cmx.h): index storage in HBM, search
// orchestration (slab schedule -> scoring kernel -> compaction), the vector-mix
// prologue and the shard merge.  Host-side control only; all arithmetic is in the
// kernels of prologue.cu / stream_score.cu / tc_score.cu / select.cu.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace cmx {

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};
int g_profiling = 0;
static int g_default_precision = CMX_PRECISION_RESCORE;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
}

struct DevGuard {
  int prev = -1;
  bool ok = true;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <typename T>
static int ensure_buf(T** ptr, int64_t* cap, int64_t need) {
  if (need <= *cap && *ptr) return CMX_OK;
  if (*ptr) { cudaFree(*ptr); *ptr = nullptr; *cap = 0; }
  if (need <= 0) need = 1;
  cudaError_t e = cudaMalloc((void**)ptr, (size_t)need * sizeof(T));
  if (e != cudaSuccess) {
    set_error("device allocation of %lld bytes failed: %s", (long long)(need * (int64_t)sizeof(T)), cudaGetErrorString(e));
    *ptr = nullptr;
    return CMX_ERR_NOMEM;
  }
  *cap = need;
  return CMX_OK;
}

constexpr int kMaxSlabEvents = 64;
constexpr int64_t kQueryChunk = 8192;  // queries per corpus pass (bounds workspace + Q L2 footprint)
constexpr int kMaxSub = 8;             // scoring launches the last slab of a rescore-mode search is cut into

}  // namespace cmx

using namespace cmx;

struct cmx_index {
  int d = 0, d_pad = 0, device = 0, sm_count = 148;
  int64_t n = 0, cap_rows = 0;
  float* X = nullptr;  // fp32 row store [cap_rows, d]
  // fp16 hi/lo planes for the tensor path [cap_rows, d_pad], built lazily
  __half* Bhi = nullptr;
  __half* Blo = nullptr;
  int64_t plane_cap = 0, plane_rows = 0;       // hi plane
  int64_t lo_cap = 0, lo_rows = 0;             // lo plane (split precision only)
  float plane_scale = 0.f;
  uint32_t absmax_bits = 0;  // max |x| over all finite stored elements
  uint32_t* absmax_dev = nullptr;   // [2]: absmax bits, max row norm bits
  float row_norm_max = 0.f;  // max ||row||_2 over finite rows (rescore-mode error bound)
  float row_resid_max = 0.f;  // max ||row - hi plane row||_2 over the rows split so far
  float norm_floor = 0.f, resid_floor = 0.f;  // lower limits set from outside (maxima over all shards)
  int precision = CMX_PRECISION_RESCORE;  // set from the process default at creation
  // search workspace
  SearchWs ws;
  int64_t tau_cap = 0, cnt_cap = 0, cand_cap_elems = 0, margin_cap = 0;
  int cand_cap_override = 0;
  float* q_dev = nullptr; int64_t q_cap = 0;        // staged / mixed queries
  float* p_dev = nullptr; int64_t p_cap = 0;        // staged P
  float* s_dev = nullptr; int64_t s_cap = 0;        // staged S
  __half* Qhi = nullptr; int64_t qhi_cap = 0;
  __half* Qlo = nullptr; int64_t qlo_cap = 0;
  uint32_t* q_absmax = nullptr;
  float* margin_buf = nullptr;
  unsigned long long* progress = nullptr;  // tile-progress counter of the tensor kernels
  // two-phase (sharded) search state between cmx_search_begin and cmx_search_end
  bool pending = false;  // cleared by every entry point that touches the workspace or the row store
  int64_t pend_nq = 0, pend_id_base = 0;
  int pend_k = 0;
  float* q_scale = nullptr;  // {scale, 1/scale}
  float* D_dev = nullptr; int64_t D_cap = 0;
  int64_t* I_dev = nullptr; int64_t I_cap = 0;
  uint8_t* flags_dev = nullptr; int64_t flags_cap = 0;
  cudaEvent_t ev[4 * kMaxSlabEvents];
  bool ev_ready = false;
  int timed_slabs = 0;  // slabs of an asynchronous pass whose event times have not been read yet
  // prescoring (exact scores ahead of time, beside the scoring kernel): second stream, launch-boundary
  // events, snapshots of the append cursors
  cudaStream_t side = nullptr;
  cudaEvent_t ev_sub[kMaxSub + 1], ev_side_done = nullptr;
  uint32_t* snap = nullptr; int64_t snap_cap = 0;  // [kMaxSub + 1, nq_pad]
  float* est_buf = nullptr; int64_t est_cap = 0;
  const float* pend_q = nullptr;  // queries of the pending two-phase search (device)
  bool pend_skip = false;
  float* bounds_dev = nullptr;    // {max row norm, max residual norm} over all shards, reduced on device
  bool bounds_from_dev = false;
  cmx_search_stats stats;
};

namespace cmx {

static int ensure_events(cmx_index* ix) {
  if (ix->ev_ready) return CMX_OK;
  for (int i = 0; i < 4 * kMaxSlabEvents; ++i) CMX_CUDA(cudaEventCreate(&ix->ev[i]));
  ix->ev_ready = true;
  return CMX_OK;
}

static int grow_store(cmx_index* ix, int64_t need_rows) {
  if (need_rows <= ix->cap_rows) return CMX_OK;
  int64_t new_cap = std::max<int64_t>(need_rows, ix->cap_rows + ix->cap_rows / 2);
  float* nx = nullptr;
  cudaError_t e = cudaMalloc((void**)&nx, (size_t)new_cap * ix->d * sizeof(float));
  if (e != cudaSuccess && new_cap > need_rows) {
    new_cap = need_rows;
    e = cudaMalloc((void**)&nx, (size_t)new_cap * ix->d * sizeof(float));
  }
  if (e != cudaSuccess) {
    set_error("cannot allocate %.2f GB for %lld rows x %d: %s", (double)new_cap * ix->d * 4 / 1e9,
              (long long)new_cap, ix->d, cudaGetErrorString(e));
    return CMX_ERR_NOMEM;
  }
  if (ix->n > 0) CMX_CUDA(cudaMemcpy(nx, ix->X, (size_t)ix->n * ix->d * sizeof(float), cudaMemcpyDeviceToDevice));
  if (ix->X) cudaFree(ix->X);
  ix->X = nx;
  ix->cap_rows = new_cap;
  // planes follow the store lazily
  if (ix->Bhi) { cudaFree(ix->Bhi); ix->Bhi = nullptr; }
  if (ix->Blo) { cudaFree(ix->Blo); ix->Blo = nullptr; }
  ix->plane_cap = ix->plane_rows = 0;
  ix->lo_cap = ix->lo_rows = 0;
  return CMX_OK;
}

// build / extend the fp16 operand planes so they cover rows [0, n): the hi plane always,
// the lo plane only for the split-precision scorer
static int ensure_planes(cmx_index* ix, bool need_lo, cudaStream_t st) {
  const float want_scale = host_scale_for_absmax_bits(ix->absmax_bits);
  const size_t bytes = (size_t)ix->cap_rows * ix->d_pad * sizeof(__half);
  if (ix->plane_scale != want_scale) {  // a later add() raised max|x|: re-split everything
    ix->plane_rows = 0;
    ix->lo_rows = 0;
    ix->plane_scale = want_scale;
    ix->row_resid_max = 0.f;
  }
  if (ix->plane_cap < ix->cap_rows || !ix->Bhi) {
    if (ix->Bhi) { cudaFree(ix->Bhi); ix->Bhi = nullptr; }
    if (cudaMalloc((void**)&ix->Bhi, bytes) != cudaSuccess) {
      cudaGetLastError();
      ix->Bhi = nullptr;
      ix->plane_cap = 0;
      set_error("cannot allocate %.2f GB for the fp16 operand plane (tensor path); shard the index over more GPUs",
                (double)bytes / 1e9);
      return CMX_ERR_NOMEM;
    }
    ix->plane_cap = ix->cap_rows;
    ix->plane_rows = 0;
  }
  if (need_lo && (ix->lo_cap < ix->cap_rows || !ix->Blo)) {
    if (ix->Blo) { cudaFree(ix->Blo); ix->Blo = nullptr; }
    if (cudaMalloc((void**)&ix->Blo, bytes) != cudaSuccess) {
      cudaGetLastError();
      ix->Blo = nullptr;
      ix->lo_cap = 0;
      set_error("cannot allocate %.2f GB for the fp16 lo plane (split-precision tensor path); use the rescore "
                "precision or shard the index over more GPUs", (double)bytes / 1e9);
      return CMX_ERR_NOMEM;
    }
    ix->lo_cap = ix->cap_rows;
    ix->lo_rows = 0;
  }
  // rows that still miss a plane (lo is written together with hi; hi-only rows are re-split for lo)
  const int64_t r0 = need_lo ? std::min(ix->plane_rows, ix->lo_rows) : ix->plane_rows;
  if (r0 < ix->n) {
    CMX_TRY(launch_split_planes(ix->X + r0 * ix->d, ix->n - r0, ix->d, ix->d_pad, nullptr, ix->plane_scale,
                                ix->Bhi + r0 * ix->d_pad, need_lo ? ix->Blo + r0 * ix->d_pad : nullptr, st));
    // what the hi plane alone loses of each new row (error bound of the one-pass scorer)
    CMX_CUDA(cudaMemsetAsync(ix->absmax_dev + 1, 0, sizeof(uint32_t), st));
    CMX_TRY(launch_row_resid_max(ix->X + r0 * ix->d, ix->Bhi + r0 * ix->d_pad, ix->n - r0, ix->d, ix->d_pad,
                                 1.0f / ix->plane_scale, ix->absmax_dev + 1, st));
    float res = 0.f;
    CMX_CUDA(cudaMemcpyAsync(&res, ix->absmax_dev + 1, sizeof(float), cudaMemcpyDeviceToHost, st));
    CMX_CUDA(cudaStreamSynchronize(st));
    ix->row_resid_max = std::max(ix->row_resid_max, res);
    ix->plane_rows = ix->n;
    if (need_lo) ix->lo_rows = ix->n;
  }
  return CMX_OK;
}

// the one-pass (rescore) arithmetic needs finite corpus maxima for its error bound; a corpus with a finite
// row whose fp32 sum of squares overflows reports +inf (row_norm_max_kernel) and uses split precision
static bool rescore_usable(const cmx_index* ix) {
  return ix->precision == CMX_PRECISION_RESCORE && ix->row_norm_max > 0.f && std::isfinite(ix->row_norm_max) &&
         std::isfinite(ix->norm_floor) && std::isfinite(ix->resid_floor);
}

static int pick_cap(const cmx_index* ix, int k) {
  int cap = ix->cand_cap_override;
  if (cap <= 0) cap = (k <= 1024) ? 8192 : 16384;
  int p = 256;
  while (p < cap) p <<= 1;
  cap = p;
  while (cap < 2 * k) cap <<= 1;
  return cap;
}

struct SlabPlan {
  std::vector<int64_t> rows;   // slab sizes, in order; slab 0 is the dense one
  std::vector<int> spec_rank;  // > 0: the slab runs under a speculative threshold, the spec_rank-th best score after
                               // the slab before it (published and later verified by the compactions); 0: planned slab
  int nspec() const { int c = 0; for (int r : spec_rank) c += r > 0; return c; }
};

// Geometric slab schedule: slab 0 fills the (empty) candidate buffers densely, every
// later slab is sized so that, for rows arriving in exchangeable order, the expected
// number of rows beating the stale threshold is half of the free room (cap - k).
// safe = worst-case schedule (every row may pass): rows <= cap - k per slab.
//
// Speculative slabs (spec = true; tensor path, whose block order makes every prefix a uniform
// sample).  Once `seen` rows are in, the k-th best of the first T >= seen rows is expected near the
// r0 = k seen/T -th best so far.  A slab covering rows [seen, T) may therefore be filtered at the deeper
// rank R = kSpecDepth r0: ~kSpecDepth k survivors per query whatever T is, while the chance that fewer
// than k of the T rows clear it is P(Poisson(r0) >= R).  The compaction after the slab VERIFIES the guess
// per query (k-th best >= threshold + margin); a miss only costs a rerun with spec = false.  Two uses:
//   final: T = N as soon as r0 >= kSpecMinRank -- the rest of the corpus in ONE slab;
//   mid:   otherwise the largest T with r0 = kSpecMinRank, if that beats the geometric slab -- at
//          k = 1000 the dense slab of 8 192 rows is followed by one slab of 335 k rows instead of
//          three geometric ones (20 k, 74 k, 262 k) and their compactions.
// kSpecDepth balances two tails (r0 >= 32, buffers of 8192 keys, k' = 1341 resident):
//   guess too HIGH (rerun):      P(Poisson(r0) >= 2.5 r0) < 1e-13 per query
//   guess too LOW (buffer overflow, rerun): the sample holds R rows above the threshold where the corpus rate
//       predicts >= 1.85 R:  P(Poisson(1.85 R) <= R) ~ 1e-9 at R = 80
// Round 1 used depth 3: overflow needed only 1.57 R, P ~ 2e-7 per (query, shard) -- an 8-GPU 11-alpha job
// (614 k pairs) hit it (shard 6, alpha 0.7, query 2722: 125 sample rows above the threshold where 197 were due,
// 7 292 survivors: scripts/diag_emulate.py).  The Poisson figures assume independent rows; if a query's best
// rows come in clumps of c adjacent rows (passages of one document) the effective rank is r0/c -- at r0 >= 32
// and c = 3 a miss is still < 1e-5 per query, and neither tail is ever wrong, only a second pass.
constexpr double kSpecDepth = 2.5;
constexpr double kSpecMinRank = 32.0;
static int g_speculate = 1;  // 0: planned geometric slabs only (experiments)

static SlabPlan plan_slabs_one(int64_t N, int k, int cap, int align, bool safe, bool spec, int k_out, bool mid,
                               int64_t first_rows = 0) {
  SlabPlan pl;
  int64_t seen = 0;
  const int64_t room = cap - k;
  auto round_dn = [&](int64_t v) { return std::max<int64_t>(align, v / align * align); };
  auto push = [&](int64_t rows, int rank) { pl.rows.push_back(rows); pl.spec_rank.push_back(rank); seen += rows; };
  while (seen < N) {
    const double g = (double)seen * (double)room / (2.0 * (double)k);  // geometric slab
    if (spec && !safe && g_speculate && seen > 0) {
      const int rank_min = (int)std::ceil(kSpecDepth * kSpecMinRank);
      const double r0 = (double)k * (double)seen / (double)N;  // expected rank of the final k-th best
      // final: worth it only if the geometric plan still needs two or more slabs
      // (a guess admits ~kSpecDepth * k rows on top of the k kept ones: only with a buffer that has room for them)
      if (r0 >= kSpecMinRank && std::ceil(kSpecDepth * r0) < (double)k_out && kSpecDepth * r0 < 0.75 * k && 4.5 * k <= cap &&
          (double)(N - seen) > g) {
        push(N - seen, (int)std::ceil(kSpecDepth * r0));
        break;
      }
      // mid: up to the row count T at which r0 = kSpecMinRank
      const int64_t T = (int64_t)((double)k * (double)seen / kSpecMinRank);
      if (mid && r0 < kSpecMinRank && rank_min < k_out && kSpecDepth * kSpecMinRank < 0.75 * k && 4.5 * k <= cap &&
          T - seen > (int64_t)(1.5 * g)) {
        push(std::min(round_dn(T - seen), N - seen), rank_min);
        continue;
      }
    }
    int64_t rows;
    if (seen == 0) {
      rows = round_dn(first_rows > 0 ? std::min<int64_t>(first_rows, cap) : cap);
      if (rows > cap) rows = cap;  // align > cap cannot happen (cap >= 256)
    } else if (safe) {
      // no more rows than a buffer has room for; a room smaller than the alignment unit is served
      // in pieces that never straddle a unit (the tensor kernels address whole 256-row blocks)
      rows = room >= align ? room / align * align : std::min<int64_t>(room, align - seen % align);
    } else {
      rows = round_dn((int64_t)std::min<double>(g, 4e18));
    }
    push(std::min(rows, N - seen), 0);
  }
  return pl;
}

// A speculative mid slab is used only where it saves launches.  With one, the dense first slab need not fill the
// candidate buffers: it only has to be a sample large enough for the mid slab's guess (its kSpecDepth * kSpecMinRank-th
// best), and dense rows are the expensive ones -- every score of the slab is written out and read back by the first
// compaction (0.7 ms for 8192 rows x 6980 queries against 0.06 ms of tensor work).  So among the plans with the fewest
// slabs the one with the smallest first slab wins; kDenseMinRows keeps the sample at eight or more blocks of the
// permuted order.
constexpr int64_t kDenseMinRows = 2048;
static int g_small_first = 1;  // 0: the first slab always fills the buffers (experiments)
static SlabPlan plan_slabs(int64_t N, int k, int cap, int align, bool safe, bool spec = false, int k_out = 0) {
  SlabPlan plain = plan_slabs_one(N, k, cap, align, safe, spec, k_out, false);
  if (!spec || safe || !g_speculate) return plain;
  SlabPlan mid = plan_slabs_one(N, k, cap, align, safe, spec, k_out, true);
  SlabPlan best = mid.rows.size() < plain.rows.size() ? mid : plain;
  if (!g_small_first || best.nspec() == 0) return best;
  const int64_t f_min = std::max<int64_t>(kDenseMinRows, ((int64_t)k * 3 / 2 + 255) / 256 * 256);
  for (int64_t f = f_min; f < cap; f += 256) {
    SlabPlan cand = plan_slabs_one(N, k, cap, align, safe, spec, k_out, true, f);
    if (cand.rows.size() <= best.rows.size() && cand.nspec() == (int)cand.rows.size() - 1) return cand;
  }
  return best;
}

// Block multiplier of the tensor path's processing order (TcParams::perm): an integer P near
// nblk / golden ratio, coprime to nblk, so that j -> (j * P) mod nblk is a permutation of the blocks
// whose prefixes are spread evenly over the corpus (three-distance theorem) -- thresholds learnt on
// the first slabs then hold for a corpus that is not stationary in file order (EN rows then ZH
// rows in the bilingual index; passages grouped by source document).  Rounding nblk/phi to an
// integer can ruin this: P/nblk is rational, and one large partial quotient a_i in its continued
// fraction makes prefixes longer than the convergent's denominator pile up next to earlier
// points (21346/34539 = [0;1,...,1,73,2]: beyond 233 blocks the new ones land 2 blocks from old
// ones).  So the candidates within +-256 of nblk/phi are ranked by their largest partial quotient.
static int g_block_order = 1;  // 0: file order (experiments / adversarial tests)
static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }
static uint64_t max_partial_quotient(uint64_t p, uint64_t q) {  // of p/q, 0 < p < q, skipping the leading 0
  uint64_t worst = 0;
  uint64_t num = q, den = p;  // q/p = a1 + ...
  int depth = 0;
  while (den) {
    const uint64_t a = num / den, r = num % den;
    // the last quotient of a finite expansion is an artefact of termination ([..., a] = [..., a-1, 1])
    if (r != 0 || depth == 0) worst = std::max(worst, a);
    else worst = std::max(worst, a - 1);
    num = den;
    den = r;
    ++depth;
  }
  return worst;
}
static uint64_t pick_perm(int64_t nblk) {
  if (!g_block_order || nblk < 3) return 1;
  const uint64_t n = (uint64_t)nblk;
  const uint64_t centre = std::max<uint64_t>(1, (uint64_t)((double)nblk * 0.6180339887498949));
  uint64_t best = 0, best_q = ~0ull, best_dist = ~0ull;
  const uint64_t lo = centre > 256 ? centre - 256 : 1, hi = std::min<uint64_t>(n - 1, centre + 256);
  for (uint64_t p = lo; p <= hi; ++p) {
    if (gcd_u64(p, n) != 1) continue;
    const uint64_t q = max_partial_quotient(p, n);
    const uint64_t dist = p > centre ? p - centre : centre - p;
    if (q < best_q || (q == best_q && dist < best_dist)) { best = p; best_q = q; best_dist = dist; }
  }
  return best ? best : n - 1;  // n - 1 is always coprime
}

// workspace + (tensor path) operand planes, query planes and rescore margins for queries q_d[0..nq)
static int prepare_pass(cmx_index* ix, const float* q_d, int64_t nq, int k, int path, bool rescore, cudaStream_t st) {
  const int cap = pick_cap(ix, k);
  const int64_t nq_pad = (nq + 127) / 128 * 128;
  CMX_TRY(ensure_buf(&ix->ws.tau, &ix->tau_cap, 2 * nq_pad));
  ix->ws.spec = ix->ws.tau + nq_pad;
  CMX_TRY(ensure_buf(&ix->ws.cnt, &ix->cnt_cap, nq_pad));
  CMX_TRY(ensure_buf(&ix->ws.cand, &ix->cand_cap_elems, nq * (int64_t)cap));
  if (!ix->ws.overflow) CMX_CUDA(cudaMalloc((void**)&ix->ws.overflow, sizeof(uint32_t)));
  ix->ws.cap = cap;
  ix->ws.margin = nullptr;
  CMX_TRY(launch_ws_init(ix->ws, nq, nq_pad, st));

  if (path == CMX_PATH_TENSOR) {
    CMX_TRY(ensure_planes(ix, !rescore, st));
    CMX_TRY(ensure_buf(&ix->Qhi, &ix->qhi_cap, nq_pad * (int64_t)ix->d_pad));
    if (!rescore) CMX_TRY(ensure_buf(&ix->Qlo, &ix->qlo_cap, nq_pad * (int64_t)ix->d_pad));
    if (!ix->q_absmax) CMX_CUDA(cudaMalloc((void**)&ix->q_absmax, sizeof(uint32_t)));
    if (!ix->q_scale) CMX_CUDA(cudaMalloc((void**)&ix->q_scale, 2 * sizeof(float)));
    if (!ix->progress) CMX_CUDA(cudaMalloc((void**)&ix->progress, sizeof(unsigned long long)));
    CMX_CUDA(cudaMemsetAsync(ix->q_absmax, 0, sizeof(uint32_t), st));
    CMX_TRY(launch_absmax(q_d, nq * (int64_t)ix->d, ix->q_absmax, st));
    CMX_TRY(launch_scale_from_absmax(ix->q_absmax, ix->q_scale, st));
    if (nq_pad > nq) {
      CMX_CUDA(cudaMemsetAsync(ix->Qhi + nq * (int64_t)ix->d_pad, 0, (size_t)(nq_pad - nq) * ix->d_pad * sizeof(__half), st));
      if (!rescore)
        CMX_CUDA(cudaMemsetAsync(ix->Qlo + nq * (int64_t)ix->d_pad, 0, (size_t)(nq_pad - nq) * ix->d_pad * sizeof(__half), st));
    }
    CMX_TRY(launch_split_planes(q_d, nq, ix->d, ix->d_pad, ix->q_scale, 1.0f, ix->Qhi, rescore ? nullptr : ix->Qlo, st));
    if (rescore) {
      // |approx - exact| <= eps(q), margin = 2 eps(q): see query_margin_kernel and DESIGN.md 4b.  gamma =
      // 2*d*2^-23 covers the fp32 accumulation error of the tensor core (measured: tests/test_gpu_eps_bound.py)
      // plus that of the exact rescoring chain.
      const float gamma = 2.0f * (float)ix->d * 1.1920929e-07f;
      CMX_TRY(ensure_buf(&ix->margin_buf, &ix->margin_cap, nq_pad));
      CMX_TRY(launch_query_margin(q_d, ix->Qhi, nq, ix->d, ix->d_pad, ix->q_scale, std::max(ix->row_norm_max, ix->norm_floor),
                                  std::max(ix->row_resid_max, ix->resid_floor), ix->bounds_from_dev ? ix->bounds_dev : nullptr,
                                  gamma, ix->margin_buf, st));
      ix->ws.margin = ix->margin_buf;
    }
  }
  return CMX_OK;
}

// ---- prescoring resources ----------------------------------------------------------------------
// Prescoring is OFF by default (mode 0).  Measured at C2 on power-capped B200s (profiles/r02_prescore_experiments.md):
// the chip runs at its power limit while scoring, so work moved beside the scoring kernel is paid for in clock --
// hiding all of the rescoring (depth 1.5) slows scoring by as much as it saves, the surest winners only (depth
// 0.75) nets ~1.5 ms of 112 -- inside the box-to-box noise, and it lengthens the scoring launches the roofline is
// computed on.  Kept selectable (cmx_debug_set_prescore) with its bit-identity test: on a part that is not
// power-limited the overlap is worth the full 5-6 ms.
// mode 1: on; 2: cut the last slab into launches but do not prescore (experiments).
static int g_prescore = 0;
static double g_prescore_depth = 0.75;  // est rank = depth x the expected rank of the final k-th best
constexpr int64_t kPrescoreMinRows = 2 << 20;  // shorter last slabs: not worth the extra launches and selects
static int g_prescore_max_sub = kMaxSub;
static int64_t g_prescore_min_rows = kPrescoreMinRows;
constexpr int64_t kSubRows = 180 * 1024;  // rows per scoring launch of a prescored last slab (>= 2 ms of tensor work)

static int ensure_side(cmx_index* ix) {
  if (ix->side) return CMX_OK;
  CMX_CUDA(cudaStreamCreateWithFlags(&ix->side, cudaStreamNonBlocking));
  for (int i = 0; i <= kMaxSub; ++i) CMX_CUDA(cudaEventCreateWithFlags(&ix->ev_sub[i], cudaEventDisableTiming));
  CMX_CUDA(cudaEventCreateWithFlags(&ix->ev_side_done, cudaEventDisableTiming));
  return CMX_OK;
}

// event times of the slabs of the last pass -> stats (the events must have completed)
static void collect_slab_times(cmx_index* ix) {
  for (int s = 0; s < ix->timed_slabs; ++s) {
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, ix->ev[4 * s + 0], ix->ev[4 * s + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
    if (cudaEventElapsedTime(&b, ix->ev[4 * s + 1], ix->ev[4 * s + 2]) != cudaSuccess) { cudaGetLastError(); continue; }
    ix->stats.score_ms += a;
    ix->stats.select_ms += b;
    if (getenv("CMX_DEBUG_SLABS")) fprintf(stderr, "[cmx] slab %d score %.3f ms select %.3f ms\n", s, a, b);
  }
  ix->timed_slabs = 0;
}

// one pass over the corpus for queries q_d[0..nq) (nq <= kQueryChunk)
// rescore = true: the tensor kernels run ONE fp16 MMA pass (approximate scores), the buffers keep
// everything within 2*eps(q) of the k-th best approximate score, and the survivors get exact fp32
// scores at the end; rescore = false: three-pass split precision, scores final as they come
// defer = true (two-phase sharded search): stop after the last compaction, leaving the candidate
// superset in the workspace; cmx_search_end rescoring follows once the shards have exchanged their
// k-th best approximate scores.  overflowed == NULL: asynchronous -- nothing is read back, the
// caller looks at ws.overflow (and the event times) later.
// est_scale: fraction of this shard's k best that is expected in the final answer (1 / number of
// shards): sets how deep prescore_kernel goes.
static int search_pass(cmx_index* ix, const float* q_d, int64_t nq, int k, float* D_d, int64_t* I_d,
                       int64_t id_base, int path, bool rescore, bool safe, bool speculate, cudaStream_t st,
                       unsigned* overflowed, bool defer = false, float est_scale = 1.0f) {
  if (path != CMX_PATH_TENSOR) rescore = false;
  CMX_NVTX("cmx:search_pass");
  {
    CMX_NVTX("cmx:prologue");
    CMX_TRY(prepare_pass(ix, q_d, nq, k, path, rescore, st));
  }
  const int cap = ix->ws.cap;
  const int64_t nq_pad = (nq + 127) / 128 * 128;

  const int align = (path == CMX_PATH_TENSOR) ? 256 : 32;
  // rescore mode keeps the margin band on top of the k best: plan for ~4/3 k resident candidates
  const int k_plan = rescore ? std::min(cap / 2, k + k / 3 + 8) : k;
  // tensor path: whole 256-row blocks in permuted order (the tail block's padding rows yield null keys)
  const int64_t nblk = (ix->n + 255) / 256;
  const uint64_t perm = pick_perm(nblk);
  const int64_t n_plan = (path == CMX_PATH_TENSOR) ? nblk * 256 : ix->n;
  SlabPlan pl = plan_slabs(n_plan, k_plan, cap, align, safe, speculate && path == CMX_PATH_TENSOR, k);
  const int nslabs = (int)pl.rows.size();
  const bool prof = g_profiling && !safe && nslabs <= kMaxSlabEvents;
  if (prof) CMX_TRY(ensure_events(ix));
  ix->timed_slabs = 0;

  // Prescoring: the last slab is cut into up to kMaxSub scoring launches; between them the append
  // cursors are snapshotted, and prescore_kernel -- on a second stream, beside the next scoring
  // launch -- gives exact scores to the likely winners found so far (select.cu).
  int nsub = 1;
  if (rescore && !safe && g_prescore && nslabs >= 2 && pl.rows[nslabs - 1] >= g_prescore_min_rows &&
      prescore_smem_bytes(ix->d) <= kPrescoreMaxSmem) {
    nsub = (int)std::min<int64_t>(g_prescore_max_sub, std::max<int64_t>(1, pl.rows[nslabs - 1] / kSubRows));
    CMX_TRY(ensure_side(ix));
    CMX_TRY(ensure_buf(&ix->snap, &ix->snap_cap, (int64_t)(kMaxSub + 1) * nq_pad));
    CMX_TRY(ensure_buf(&ix->est_buf, &ix->est_cap, nq_pad));
    ix->ws.est = ix->est_buf;
  } else {
    ix->ws.est = nullptr;
  }
  const bool prescoring = ix->ws.est != nullptr;

  auto score = [&](int s, int64_t row0, int64_t rows, int64_t seen_before) -> int {
    const int dense = (s == 0) ? 1 : 0;
    if (path == CMX_PATH_TENSOR)
      return launch_tensor_score(ix->Bhi, ix->Blo, ix->n, row0, rows, ix->d_pad, ix->Qhi, ix->Qlo, nq, nq_pad, ix->q_scale + 1,
                                 1.0f / ix->plane_scale, ix->ws, dense, perm, rescore ? 1 : 3,
                                 seen_before > 0 ? (double)(pl.spec_rank[s] > 0 ? pl.spec_rank[s] : k_plan) / (double)seen_before : 1.0,
                                 ix->progress, st, ix->sm_count);
    return launch_stream_score(ix->X, row0, rows, ix->d, q_d, (int)nq, ix->ws, 0, dense, row0, st, ix->sm_count);
  };

  int64_t seen = 0;
  for (int s = 0; s < nslabs; ++s) {
    const int64_t rows = pl.rows[s];
    const int last = (s == nslabs - 1) ? 1 : 0;
    if (prof) CMX_CUDA(cudaEventRecord(ix->ev[4 * s + 0], st));
    {
      CMX_NVTX("cmx:score");
      if (last && prescoring) {
        // candidates kept so far: exact scores while the first launch runs
        CMX_TRY(launch_snapshot_counts(ix->ws, nq, ix->snap, st));
        CMX_CUDA(cudaEventRecord(ix->ev_sub[0], st));
        CMX_CUDA(cudaStreamWaitEvent(ix->side, ix->ev_sub[0], 0));
        if (g_prescore == 1) CMX_TRY(launch_prescore(ix->X, ix->d, q_d, ix->ws, nq, nullptr, ix->snap, ix->side));
        const int64_t blocks = rows / 256;  // whole 256-row blocks (a speculative / geometric slab always is)
        int64_t b0 = 0;
        for (int i = 0; i < nsub; ++i) {
          const int64_t b1 = (i == nsub - 1) ? blocks : blocks * (i + 1) / nsub;
          const int64_t r0 = b0 * 256, r1 = (i == nsub - 1) ? rows : b1 * 256;
          if (r1 > r0) CMX_TRY(score(s, seen + r0, r1 - r0, seen));
          if (i < nsub - 1) {
            uint32_t* snap_hi = ix->snap + (int64_t)(i + 1) * nq_pad;
            CMX_TRY(launch_snapshot_counts(ix->ws, nq, snap_hi, st));
            CMX_CUDA(cudaEventRecord(ix->ev_sub[i + 1], st));
            CMX_CUDA(cudaStreamWaitEvent(ix->side, ix->ev_sub[i + 1], 0));
            if (g_prescore == 1) CMX_TRY(launch_prescore(ix->X, ix->d, q_d, ix->ws, nq, ix->snap + (int64_t)i * nq_pad, snap_hi, ix->side));
          }
          b0 = b1;
        }
        CMX_CUDA(cudaEventRecord(ix->ev_side_done, ix->side));
        CMX_CUDA(cudaStreamWaitEvent(st, ix->ev_side_done, 0));  // the compaction moves keys: prescoring must be done
      } else {
        CMX_TRY(score(s, seen, rows, seen));
      }
    }
    if (prof) CMX_CUDA(cudaEventRecord(ix->ev[4 * s + 1], st));
    if (s == 0) CMX_TRY(launch_set_counts(ix->ws, nq, (uint32_t)rows, st));
    const int spec_rank = last ? 0 : pl.spec_rank[s + 1];  // publish the guess for the next slab
    const int verify = pl.spec_rank[s] > 0 ? 1 : 0;        // this slab ran under a guess
    int est_rank = 0;
    if (prescoring && s == nslabs - 2) {
      // where the FINAL k-th best is expected among the rows seen so far (x est_scale when only that share of
      // this shard's best can make the global answer); 1.5x deeper: wasted prescoring is hidden, missed is not
      const double r0 = (double)k_plan * (double)(seen + rows) / (double)n_plan * (double)est_scale;
      est_rank = (int)std::min<double>(std::max(1.0, std::ceil(g_prescore_depth * r0)), (double)(k - 1));
    }
    {
      CMX_NVTX(last ? "cmx:compact+rescore" : "cmx:compact");
      if (rescore) {
        // keep the margin band, then (after the last slab) exact fp32 scores + exact top-k
        CMX_TRY(launch_compact(ix->ws, nq, k, 0, D_d, I_d, id_base, st, spec_rank, verify, est_rank));
        if (last && !defer) CMX_TRY(launch_rescore(ix->X, ix->d, q_d, ix->ws, nq, k, D_d, I_d, id_base, RescoreCut(), st));
      } else {
        CMX_TRY(launch_compact(ix->ws, nq, k, last, D_d, I_d, id_base, st, spec_rank, verify, 0));
      }
    }
    if (prof) CMX_CUDA(cudaEventRecord(ix->ev[4 * s + 2], st));
    seen += rows;
  }
  ix->stats.slabs += nslabs;
  ix->stats.score_launches += (path == CMX_PATH_TENSOR) ? nslabs - 1 + nsub : nslabs * (int)((nq + 7) / 8);
  ix->stats.select_launches += nslabs + ((rescore && !defer) ? 1 : 0);
  if (prof) ix->timed_slabs = nslabs;
  if (overflowed == nullptr) return CMX_OK;  // asynchronous: flags and times are read later
  uint32_t ovf = 0;
  CMX_CUDA(cudaMemcpyAsync(&ovf, ix->ws.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CMX_CUDA(cudaStreamSynchronize(st));
  *overflowed = ovf;
  collect_slab_times(ix);
  return CMX_OK;
}

static int fill_empty(float* D_d, int64_t* I_d, int64_t count, cudaStream_t st) {
  std::vector<float> d((size_t)count, CMX_NEG_PAD);
  std::vector<int64_t> i((size_t)count, -1);
  // cudaMemcpyDefault: the outputs may be device memory or device-mapped pinned host memory
  CMX_CUDA(cudaMemcpyAsync(D_d, d.data(), count * sizeof(float), cudaMemcpyDefault, st));
  CMX_CUDA(cudaMemcpyAsync(I_d, i.data(), count * sizeof(int64_t), cudaMemcpyDefault, st));
  CMX_CUDA(cudaStreamSynchronize(st));
  return CMX_OK;
}

// all pointers on the index's device
static int search_device(cmx_index* ix, const float* q_d, int64_t nq, int k, float* D_d, int64_t* I_d,
                         int64_t id_base, int path, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  if (ix->n == 0) return fill_empty(D_d, I_d, nq * (int64_t)k, st);
  // measured on B200 (profiles/): per 36 GB corpus sweep the fp32 stream scorer takes 4.9-5.3 ms
  // (1-4 queries, 7.0-7.3 TB/s), the split-precision tensor kernels 5.0-5.2 ms (up to 32 queries),
  // the one-pass tensor kernels of the rescore mode 2.5-2.8 ms (they read only the 18 GB hi plane)
  if (path == CMX_PATH_AUTO) {
    const bool rescore_ok = rescore_usable(ix);
    path = (rescore_ok || nq > 4) ? CMX_PATH_TENSOR : CMX_PATH_STREAM;
  }
  if (path == CMX_PATH_STREAM && (ix->d & 3) != 0) path = CMX_PATH_TENSOR;
  ix->stats.path = path;
  for (int64_t q0 = 0; q0 < nq; q0 += kQueryChunk) {
    const int64_t nqc = std::min<int64_t>(kQueryChunk, nq - q0);
    const bool rescore = path == CMX_PATH_TENSOR && rescore_usable(ix);
    // Attempts, cheapest first; stats.reruns = how many of them had to be repeated.  An attempt is a
    // precision (rescore: one fp16 pass + exact rescoring of the margin band; split: exact scores, no
    // band) and a slab schedule (0 = geometric slabs with a speculative last slab, 1 = geometric
    // slabs, 2 = worst-case-safe slabs: so small that no buffer can overflow, split precision only).
    //   band of some query wider than half its buffer (thousands of rows within 2*eps of the k-th
    //   score, e.g. near-duplicates)            -> same schedule in split precision
    //   buffer overflow / speculation not cleared -> next schedule (rows arrive in an order that is
    //   adversarial for thresholds learnt on a sample); rescore gives way to split before schedule 2
    bool resc = rescore;
    int sched = (path == CMX_PATH_TENSOR && g_speculate) ? 0 : 1;  // the stream scorer walks the file in order: no guesses
    unsigned ovf = 1;
    for (int attempt = 0; attempt < 6 && ovf; ++attempt) {
      ix->stats.reruns = std::max(ix->stats.reruns, attempt);
      CMX_TRY(search_pass(ix, q_d + q0 * ix->d, nqc, k, D_d + q0 * k, I_d + q0 * k, id_base, path, resc, sched == 2,
                          sched == 0, st, &ovf));
      if (!ovf) break;
      if (resc && ((ovf & CMX_OVF_BAND) || sched >= 1)) resc = false;  // split precision, same schedule
      else if (sched < 2) ++sched;
      else break;
    }
    if (ovf) { set_error("internal: candidate buffer overflow in safe mode"); return CMX_ERR_INTERNAL; }
  }
  return CMX_OK;
}

// Host outputs in page-locked, device-mapped memory (cudaHostAlloc / torch pin_memory) can be
// written by the final kernels directly: the (D, I) rows -- 84 MB at C2 -- then cross PCIe as
// posted writes WHILE the rescoring kernel is still gathering rows from HBM, instead of in a
// device-to-host copy after it.  Pageable buffers (numpy) keep the staged copy.
static int g_mapped_outputs = 1;  // 0: always stage (experiments)
static bool mapped_host_outputs(float* D, int64_t* I, float** Dm, int64_t** Im) {
  if (!g_mapped_outputs) return false;
  cudaPointerAttributes a, b;
  if (cudaPointerGetAttributes(&a, D) != cudaSuccess || cudaPointerGetAttributes(&b, I) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (a.type != cudaMemoryTypeHost || b.type != cudaMemoryTypeHost || !a.devicePointer || !b.devicePointer) return false;
  *Dm = (float*)a.devicePointer;
  *Im = (int64_t*)b.devicePointer;
  return true;
}

// ---- streaming loader ---------------------------------------------------------------------------
// fills dst[0 .. bytes) from fd at `off` with `nthreads` concurrent preads; returns false on a short read / error
static bool pread_parallel(int fd, char* dst, int64_t off, int64_t bytes, int nthreads) {
  nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, bytes / (1 << 20)));
  std::vector<std::thread> th;
  std::vector<int> ok((size_t)nthreads, 1);
  for (int t = 0; t < nthreads; ++t) {
    const int64_t a = bytes * t / nthreads, b = bytes * (t + 1) / nthreads;
    th.emplace_back([=, &ok]() {
      int64_t done = a;
      while (done < b) {
        const ssize_t r = pread(fd, dst + done, (size_t)std::min<int64_t>(b - done, 1 << 30), (off_t)(off + done));
        if (r <= 0) { ok[(size_t)t] = 0; return; }
        done += r;
      }
    });
  }
  for (auto& x : th) x.join();
  for (int v : ok)
    if (!v) return false;
  return true;
}

// dst[0 .. bytes) = src[0 .. bytes) with `nthreads` concurrent memcpys (pageable host memory -> page-locked staging)
static bool memcpy_parallel(char* dst, const char* src, int64_t bytes, int nthreads) {
  nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, bytes / (4 << 20)));
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    const int64_t a = bytes * t / nthreads, b = bytes * (t + 1) / nthreads;
    th.emplace_back([=]() { memcpy(dst + a, src + a, (size_t)(b - a)); });
  }
  for (auto& x : th) x.join();
  return true;
}

// Appends n rows produced chunk by chunk by `fill(buf, first_row, rows)` (host threads writing a page-locked
// staging buffer): three buffers, the producer one chunk ahead of the copy engine, max |x| and max row norm of each
// chunk computed behind its copy.  Shared by the file loader and by add() of large pageable host arrays.
struct StagingPool {
  static constexpr int NB = 3;
  static constexpr size_t kBytes = (size_t)64 << 20;
  std::mutex mu;
  char* buf[NB] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[NB] = {nullptr, nullptr, nullptr};
  cudaStream_t st = nullptr;
  cudaError_t ensure() {  // call with the device current and `mu` held
    if (st) return cudaSuccess;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int i = 0; i < NB && e == cudaSuccess; ++i) {
      e = cudaHostAlloc((void**)&buf[i], kBytes, cudaHostAllocPortable);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
      for (int i = 0; i < NB; ++i) {
        if (buf[i]) { cudaFreeHost(buf[i]); buf[i] = nullptr; }
        if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
      }
      if (st) { cudaStreamDestroy(st); st = nullptr; }
    }
    return e;
  }
};
static StagingPool g_staging[16];

template <typename Fill>
static int stream_rows_into_store(cmx_index* ix, int64_t n, Fill fill, const char* what, double* seconds_out) {
  const auto t_begin = std::chrono::steady_clock::now();
  const int64_t row_bytes = (int64_t)ix->d * 4;
  CMX_TRY(grow_store(ix, ix->n + n));
  constexpr int NB = StagingPool::NB;
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)StagingPool::kBytes / row_bytes);
  const int64_t nchunks = (n + chunk_rows - 1) / chunk_rows;
  // page-locked staging buffers, stream and events are created once per device and reused (allocating 192 MB of
  // pinned memory per call cost 0.3 s -- more than copying a 256 MB chunk); the lock serialises loads on one device
  StagingPool& pool = g_staging[ix->device % 16];
  std::lock_guard<std::mutex> lock(pool.mu);
  cudaError_t e = pool.ensure();
  if (e != cudaSuccess) { set_error("%s: staging setup failed: %s", what, cudaGetErrorString(e)); return CMX_ERR_CUDA; }
  char** buf = pool.buf;
  cudaEvent_t* ev = pool.ev;
  cudaStream_t st = pool.st;
  int rc = CMX_OK;
  auto cleanup = [&]() {};
  e = cudaMemsetAsync(ix->absmax_dev, 0, 2 * sizeof(uint32_t), st);
  if (e != cudaSuccess) { set_error("%s: setup failed: %s", what, cudaGetErrorString(e)); cleanup(); return CMX_ERR_CUDA; }
  float* dst0 = ix->X + ix->n * ix->d;
  double t_fill = 0.0;
  std::thread producer;
  bool fill_ok = true;
  auto start_fill = [&](int64_t c) {
    const int64_t r0 = c * chunk_rows, rows = std::min(chunk_rows, n - r0);
    producer = std::thread([&, c, r0, rows]() {
      const auto t0 = std::chrono::steady_clock::now();
      fill_ok = fill(buf[c % NB], r0, rows);
      t_fill += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  };
  start_fill(0);
  for (int64_t c = 0; c < nchunks && rc == CMX_OK; ++c) {
    producer.join();
    if (!fill_ok) { set_error("%s: short read (chunk %lld)", what, (long long)c); rc = CMX_ERR_INVALID; break; }
    if (c + 1 < nchunks) {
      // buffer (c+1) % NB was last used by chunk c+1-NB: its copy must have left the host
      if (c + 1 >= NB && cudaEventSynchronize(ev[(c + 1) % NB]) != cudaSuccess) { set_error("%s: event wait failed", what); rc = CMX_ERR_CUDA; break; }
      start_fill(c + 1);
    }
    const int64_t r0 = c * chunk_rows, rows = std::min(chunk_rows, n - r0);
    float* dst = dst0 + r0 * ix->d;
    e = cudaMemcpyAsync(dst, buf[c % NB], (size_t)(rows * row_bytes), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(ev[c % NB], st);
    if (e != cudaSuccess) { set_error("%s: H2D copy failed: %s", what, cudaGetErrorString(e)); rc = CMX_ERR_CUDA; break; }
    // max |x| and max row norm of the chunk (operand scale / error bound of the tensor path), behind the copy
    rc = launch_absmax(dst, rows * (int64_t)ix->d, ix->absmax_dev, st);
    if (rc == CMX_OK) rc = launch_row_norm_max(dst, rows, ix->d, ix->absmax_dev + 1, st);
  }
  if (producer.joinable()) producer.join();
  uint32_t bits[2] = {0, 0};
  if (rc == CMX_OK) {
    e = cudaMemcpyAsync(bits, ix->absmax_dev, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); rc = CMX_ERR_CUDA; }
  } else {
    cudaStreamSynchronize(st);
  }
  cleanup();
  if (rc != CMX_OK) return rc;
  ix->absmax_bits = std::max(ix->absmax_bits, bits[0]);
  float nrm;
  memcpy(&nrm, &bits[1], sizeof(float));
  ix->row_norm_max = std::max(ix->row_norm_max, nrm);
  ix->n += n;
  if (seconds_out) {
    seconds_out[0] = t_fill;
    seconds_out[1] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  }
  return CMX_OK;
}

static int default_io_threads() { return (int)std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency() / 2)); }

static void stats_begin(cmx_index* ix, int64_t nq) {
  memset(&ix->stats, 0, sizeof(ix->stats));
  ix->stats.nq = nq;
  ix->stats.ntotal = ix->n;
  ix->stats.launches = (int32_t)g_launches.load();
}
static void stats_end(cmx_index* ix) { ix->stats.launches = (int32_t)g_launches.load() - ix->stats.launches; }

}  // namespace cmx

// =============================== C ABI ===========================================
extern "C" {

const char* cmx_last_error(void) { return t_error.c_str(); }
int cmx_version(void) { return 100; }
uint64_t cmx_launch_count(void) { return g_launches.load(); }
int cmx_set_profiling(int on) { g_profiling = on ? 1 : 0; return CMX_OK; }

int cmx_device_count(int* out) {
  CMX_CHECK(out != nullptr, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out = 0;
    set_error("no usable CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return CMX_ERR_CUDA;
  }
  *out = n;
  return CMX_OK;
}

int cmx_index_create(int d, int device, cmx_index** out) {
  CMX_CHECK(out != nullptr, "null out");
  *out = nullptr;
  CMX_CHECK(d > 0 && d <= 65536, "bad dimension %d", d);
  int ndev = 0;
  CMX_TRY(cmx_device_count(&ndev));
  CMX_CHECK(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
  DevGuard g(device);
  CMX_CHECK(g.ok, "cannot select device %d", device);
  cudaDeviceProp prop;
  CMX_CUDA(cudaGetDeviceProperties(&prop, device));
  CMX_CHECK(prop.major == 10, "libcmx is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
  cmx_index* ix = new cmx_index();
  ix->d = d;
  ix->d_pad = (d + 63) / 64 * 64;
  ix->device = device;
  ix->sm_count = prop.multiProcessorCount;
  ix->precision = g_default_precision;
  memset(&ix->stats, 0, sizeof(ix->stats));
  cudaError_t e = cudaMalloc((void**)&ix->absmax_dev, 2 * sizeof(uint32_t));
  if (e != cudaSuccess) { delete ix; set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return CMX_ERR_NOMEM; }
  *out = ix;
  return CMX_OK;
}

int cmx_index_free(cmx_index* ix) {
  if (!ix) return CMX_OK;
  DevGuard g(ix->device);
  void* ptrs[] = {ix->X, ix->Bhi, ix->Blo, ix->absmax_dev, ix->ws.tau, ix->ws.cnt, ix->ws.cand, ix->ws.overflow,
                  ix->q_dev, ix->p_dev, ix->s_dev, ix->Qhi, ix->Qlo, ix->q_absmax, ix->q_scale, ix->margin_buf, ix->progress, ix->D_dev, ix->I_dev,
                  ix->flags_dev, ix->snap, ix->est_buf, ix->bounds_dev};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (ix->side) {
    cudaStreamDestroy(ix->side);
    for (int i = 0; i <= kMaxSub; ++i) cudaEventDestroy(ix->ev_sub[i]);
    cudaEventDestroy(ix->ev_side_done);
  }
  if (ix->ev_ready)
    for (int i = 0; i < 4 * kMaxSlabEvents; ++i) cudaEventDestroy(ix->ev[i]);
  delete ix;
  return CMX_OK;
}

int cmx_index_reserve(cmx_index* ix, int64_t n) {
  CMX_CHECK(ix != nullptr, "null index");
  CMX_CHECK(n >= 0 && n < (int64_t)0x7fffff00, "reserve: row count %lld out of range (max 2^31 rows per shard)", (long long)n);
  DevGuard g(ix->device);
  ix->pending = false;
  return grow_store(ix, n);
}

int cmx_index_add(cmx_index* ix, const float* x, int64_t n, int x_on_device) {
  CMX_CHECK(ix != nullptr, "null index");
  CMX_CHECK(n >= 0, "negative row count");
  if (n == 0) return CMX_OK;
  CMX_CHECK(x != nullptr, "null data");
  CMX_CHECK(ix->n + n < (int64_t)0x7fffff00, "index would exceed 2^31 rows per shard");
  DevGuard g(ix->device);
  ix->pending = false;
  if (!x_on_device && n * (int64_t)ix->d * 4 >= ((int64_t)256 << 20)) {
    // a large host array (index_cpu_to_gpu of a read_index'ed corpus): cudaMemcpy from pageable memory is staged by
    // the driver at a few GB/s; parallel memcpys into page-locked buffers overlapped with the H2D copies are not
    cudaPointerAttributes a;
    const bool pinned = cudaPointerGetAttributes(&a, x) == cudaSuccess && a.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned) {
      const int64_t row_bytes = (int64_t)ix->d * 4;
      const int nthreads = default_io_threads();
      const char* src = reinterpret_cast<const char*>(x);
      CMX_NVTX("cmx:add(host, staged)");
      return stream_rows_into_store(
          ix, n, [=](char* buf, int64_t r0, int64_t rows) { return memcpy_parallel(buf, src + r0 * row_bytes, rows * row_bytes, nthreads); },
          "add", nullptr);
    }
  }
  CMX_TRY(grow_store(ix, ix->n + n));
  float* dst = ix->X + ix->n * ix->d;
  CMX_CUDA(cudaMemcpy(dst, x, (size_t)n * ix->d * sizeof(float), x_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  // track max |x| (finite) for the fp16 operand scale of the tensor path
  CMX_CUDA(cudaMemset(ix->absmax_dev, 0, 2 * sizeof(uint32_t)));
  CMX_TRY(launch_absmax(dst, n * (int64_t)ix->d, ix->absmax_dev, 0));
  CMX_TRY(launch_row_norm_max(dst, n, ix->d, ix->absmax_dev + 1, 0));
  uint32_t bits[2] = {0, 0};
  CMX_CUDA(cudaMemcpy(bits, ix->absmax_dev, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  ix->absmax_bits = std::max(ix->absmax_bits, bits[0]);
  {
    float nrm;
    memcpy(&nrm, &bits[1], sizeof(float));
    ix->row_norm_max = std::max(ix->row_norm_max, nrm);
  }
  ix->n += n;
  return CMX_OK;
}

int cmx_index_add_from_file(cmx_index* ix, const char* path, int64_t offset, int64_t n, int nthreads, double* seconds_out) {
  CMX_CHECK(ix != nullptr && path != nullptr, "null argument");
  CMX_CHECK(offset >= 0 && n >= 0, "bad range");
  if (n == 0) return CMX_OK;
  CMX_CHECK(ix->n + n < (int64_t)0x7fffff00, "index would exceed 2^31 rows per shard");
  DevGuard g(ix->device);
  ix->pending = false;
  CMX_NVTX("cmx:add_from_file");
  const int fd = open(path, O_RDONLY);
  CMX_CHECK(fd >= 0, "cannot open %s: %s", path, strerror(errno));
  struct Closer { int fd; ~Closer() { close(fd); } } closer{fd};
  const int64_t row_bytes = (int64_t)ix->d * 4;
  {
    const off_t size = lseek(fd, 0, SEEK_END);
    CMX_CHECK(size >= 0 && offset + n * row_bytes <= (int64_t)size, "%s holds fewer than %lld rows of %d floats at offset %lld", path,
              (long long)n, ix->d, (long long)offset);
  }
  if (nthreads <= 0) nthreads = default_io_threads();
  return stream_rows_into_store(
      ix, n, [=](char* buf, int64_t r0, int64_t rows) { return pread_parallel(fd, buf, offset + r0 * row_bytes, rows * row_bytes, nthreads); },
      path, seconds_out);
}

int cmx_read_file(const char* path, int64_t offset, int64_t bytes, void* dst, int nthreads) {
  CMX_CHECK(path && dst && offset >= 0 && bytes >= 0, "bad argument");
  if (bytes == 0) return CMX_OK;
  const int fd = open(path, O_RDONLY);
  CMX_CHECK(fd >= 0, "cannot open %s: %s", path, strerror(errno));
  const bool ok = pread_parallel(fd, (char*)dst, offset, bytes, nthreads > 0 ? nthreads : default_io_threads());
  close(fd);
  CMX_CHECK(ok, "short read from %s (%lld bytes at offset %lld)", path, (long long)bytes, (long long)offset);
  return CMX_OK;
}

// rows[i] of `src` (same device) -> appended to `ix`: the device-side form of the reference's
// np.vstack([base_index.reconstruct(e[0]) for e in batch]) (onepass_bilingual_mix_hub_custom_lang.py:644)
int cmx_index_add_gather(cmx_index* ix, const cmx_index* src, const int64_t* rows, int64_t n) {
  CMX_CHECK(ix && src && (rows || n == 0), "null argument");
  CMX_CHECK(ix != src, "source and destination must differ");
  CMX_CHECK(ix->d == src->d && ix->device == src->device, "gather needs two indexes of one dimension on one device");
  if (n == 0) return CMX_OK;
  CMX_CHECK(ix->n + n < (int64_t)0x7fffff00, "index would exceed 2^31 rows per shard");
  for (int64_t i = 0; i < n; ++i)
    CMX_CHECK(rows[i] >= 0 && rows[i] < src->n, "gather: row %lld out of range (source has %lld)", (long long)rows[i], (long long)src->n);
  DevGuard g(ix->device);
  ix->pending = false;
  CMX_TRY(grow_store(ix, ix->n + n));
  int64_t* rows_d = nullptr;
  CMX_CUDA(cudaMalloc((void**)&rows_d, (size_t)n * sizeof(int64_t)));
  cudaError_t e = cudaMemcpy(rows_d, rows, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice);
  int rc = e == cudaSuccess ? CMX_OK : CMX_ERR_CUDA;
  float* dst = ix->X + ix->n * ix->d;
  if (rc == CMX_OK) rc = launch_gather_rows(src->X, rows_d, n, ix->d, dst, 0);
  if (rc == CMX_OK && cudaMemset(ix->absmax_dev, 0, 2 * sizeof(uint32_t)) != cudaSuccess) rc = CMX_ERR_CUDA;
  if (rc == CMX_OK) rc = launch_absmax(dst, n * (int64_t)ix->d, ix->absmax_dev, 0);
  if (rc == CMX_OK) rc = launch_row_norm_max(dst, n, ix->d, ix->absmax_dev + 1, 0);
  uint32_t bits[2] = {0, 0};
  if (rc == CMX_OK && cudaMemcpy(bits, ix->absmax_dev, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) rc = CMX_ERR_CUDA;
  cudaFree(rows_d);
  if (rc != CMX_OK) {
    if (rc == CMX_ERR_CUDA) set_error("gather failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
  }
  ix->absmax_bits = std::max(ix->absmax_bits, bits[0]);
  float nrm;
  memcpy(&nrm, &bits[1], sizeof(float));
  ix->row_norm_max = std::max(ix->row_norm_max, nrm);
  ix->n += n;
  return CMX_OK;
}

int cmx_index_reset(cmx_index* ix) {
  CMX_CHECK(ix != nullptr, "null index");
  ix->pending = false;
  ix->n = 0;
  ix->plane_rows = 0;
  ix->lo_rows = 0;
  ix->absmax_bits = 0;
  ix->row_norm_max = 0.f;
  ix->row_resid_max = 0.f;
  ix->norm_floor = ix->resid_floor = 0.f;
  return CMX_OK;
}

int cmx_index_error_bounds(cmx_index* ix, float* out2) {
  CMX_CHECK(ix && out2, "null argument");
  DevGuard g(ix->device);
  if (ix->n > 0) CMX_TRY(ensure_planes(ix, false, 0));
  out2[0] = std::max(ix->row_norm_max, ix->norm_floor);
  out2[1] = std::max(ix->row_resid_max, ix->resid_floor);
  return CMX_OK;
}

int cmx_index_raise_error_bounds(cmx_index* ix, const float* in2) {
  CMX_CHECK(ix && in2, "null argument");
  CMX_CHECK(in2[0] >= 0.f && in2[1] >= 0.f, "error bounds must be finite and non-negative");  // NaN fails too
  ix->norm_floor = std::max(ix->norm_floor, in2[0]);
  ix->resid_floor = std::max(ix->resid_floor, in2[1]);
  return CMX_OK;
}

int cmx_index_memory(const cmx_index* ix, int64_t* out4) {
  CMX_CHECK(ix && out4, "null argument");
  const int64_t h = (int64_t)sizeof(__half), f = (int64_t)sizeof(float);
  out4[0] = ix->cap_rows * ix->d * f;                                                    // fp32 row store
  out4[1] = (ix->Bhi ? ix->plane_cap : 0) * ix->d_pad * h + (ix->Blo ? ix->lo_cap : 0) * ix->d_pad * h;  // fp16 operand plane(s)
  out4[2] = ix->cand_cap_elems * 8 + ix->tau_cap * f + ix->cnt_cap * 4 + ix->margin_cap * f + ix->snap_cap * 4 + ix->est_cap * f +
            ix->q_cap * f + ix->p_cap * f + ix->s_cap * f + (ix->qhi_cap + ix->qlo_cap) * h + ix->D_cap * f + ix->I_cap * 8 +
            ix->flags_cap;                                                                 // search workspace
  out4[3] = ix->n * ix->d * f;                                                           // what a FAISS flat index of these rows holds
  return CMX_OK;
}

int cmx_index_ntotal(const cmx_index* ix, int64_t* out) {
  CMX_CHECK(ix && out, "null argument");
  *out = ix->n;
  return CMX_OK;
}
int cmx_index_dim(const cmx_index* ix, int* out) {
  CMX_CHECK(ix && out, "null argument");
  *out = ix->d;
  return CMX_OK;
}
int cmx_index_device(const cmx_index* ix, int* out) {
  CMX_CHECK(ix && out, "null argument");
  *out = ix->device;
  return CMX_OK;
}
int cmx_index_data(const cmx_index* ix, const float** out) {
  CMX_CHECK(ix && out, "null argument");
  *out = ix->X;
  return CMX_OK;
}

int cmx_index_reconstruct(const cmx_index* ix, int64_t i0, int64_t n, float* out, int out_on_device) {
  CMX_CHECK(ix && out, "null argument");
  CMX_CHECK(i0 >= 0 && n >= 0 && i0 + n <= ix->n, "reconstruct: rows [%lld, %lld) out of range (ntotal %lld)",
            (long long)i0, (long long)(i0 + n), (long long)ix->n);
  if (n == 0) return CMX_OK;
  DevGuard g(ix->device);
  CMX_CUDA(cudaMemcpy(out, ix->X + i0 * ix->d, (size_t)n * ix->d * sizeof(float),
                      out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  return CMX_OK;
}

int cmx_index_set_cand_capacity(cmx_index* ix, int cap) {
  CMX_CHECK(ix != nullptr, "null index");
  CMX_CHECK(cap >= 0 && cap <= 16384, "candidate capacity %d out of range [0, 16384]", cap);
  ix->cand_cap_override = cap;
  return CMX_OK;
}

int cmx_set_default_precision(int mode) {
  CMX_CHECK(mode == CMX_PRECISION_SPLIT || mode == CMX_PRECISION_RESCORE, "bad precision mode %d", mode);
  g_default_precision = mode;
  return CMX_OK;
}

int cmx_index_set_precision(cmx_index* ix, int mode) {
  CMX_CHECK(ix != nullptr, "null index");
  CMX_CHECK(mode == CMX_PRECISION_SPLIT || mode == CMX_PRECISION_RESCORE, "bad precision mode %d", mode);
  ix->precision = mode;
  return CMX_OK;
}

int cmx_index_last_stats(const cmx_index* cix, cmx_search_stats* out) {
  CMX_CHECK(cix && out, "null argument");
  cmx_index* ix = const_cast<cmx_index*>(cix);
  if (ix->timed_slabs > 0) {  // an asynchronous pass: its events complete with the work
    DevGuard g(ix->device);
    cudaEventSynchronize(ix->ev[4 * (ix->timed_slabs - 1) + 2]);
    collect_slab_times(ix);
  }
  *out = ix->stats;
  return CMX_OK;
}

int cmx_index_search(cmx_index* ix, const float* q, int64_t nq, int k, float* D, int64_t* I, int io_on_device,
                     int64_t id_base, int path, void* stream) {
  CMX_CHECK(ix != nullptr, "null index");
  CMX_CHECK(nq >= 0, "negative query count");
  CMX_CHECK(k >= 1 && k <= CMX_MAX_K, "k=%d out of range [1, %d]", k, CMX_MAX_K);
  CMX_CHECK(path >= 0 && path <= 2, "bad path selector %d", path);
  if (nq == 0) return CMX_OK;
  CMX_CHECK(q && D && I, "null buffer");
  DevGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  ix->pending = false;
  stats_begin(ix, nq);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profiling) {
    CMX_TRY(ensure_events(ix));
    e0 = ix->ev[4 * kMaxSlabEvents - 1];
    e1 = ix->ev[4 * kMaxSlabEvents - 2];
    CMX_CUDA(cudaEventRecord(e0, st));
  }
  const float* q_d = q;
  float* D_d = D;
  int64_t* I_d = I;
  bool staged_out = false;
  if (!io_on_device) {
    CMX_TRY(ensure_buf(&ix->q_dev, &ix->q_cap, nq * (int64_t)ix->d));
    CMX_CUDA(cudaMemcpyAsync(ix->q_dev, q, (size_t)nq * ix->d * sizeof(float), cudaMemcpyHostToDevice, st));
    q_d = ix->q_dev;
    if (!mapped_host_outputs(D, I, &D_d, &I_d)) {
      CMX_TRY(ensure_buf(&ix->D_dev, &ix->D_cap, nq * (int64_t)k));
      CMX_TRY(ensure_buf(&ix->I_dev, &ix->I_cap, nq * (int64_t)k));
      D_d = ix->D_dev;
      I_d = ix->I_dev;
      staged_out = true;
    }
  }
  CMX_TRY(search_device(ix, q_d, nq, k, D_d, I_d, id_base, path, st));
  if (staged_out) {
    CMX_CUDA(cudaMemcpyAsync(D, D_d, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    CMX_CUDA(cudaMemcpyAsync(I, I_d, (size_t)nq * k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  }
  if (g_profiling) CMX_CUDA(cudaEventRecord(e1, st));
  CMX_CUDA(cudaStreamSynchronize(st));
  if (g_profiling) cudaEventElapsedTime(&ix->stats.total_ms, e0, e1);
  stats_end(ix);
  return CMX_OK;
}

static int prepare_alphas(const double* alphas, int nA, std::vector<float>& w, std::vector<int>& mode) {
  w.resize(2 * (size_t)nA);
  mode.resize(nA);
  for (int a = 0; a < nA; ++a) {
    const double al = alphas[a];
    // reference: weights are Python doubles cast to fp32 by numpy (onepass_dense_mix_run_custom_lang.py:356)
    w[a] = (float)(1.0 - al);
    w[nA + a] = (float)al;
    int sel = 0;
    if (std::fabs(al) <= 1e-8) sel = 1;
    else if (std::fabs(al - 1.0) <= 1e-8) sel = 2;
    const int fb = (std::fabs(al) > 0.5) ? 2 : 1;
    mode[a] = sel | (fb << 4);
  }
  return CMX_OK;
}

// shared by cmx_mix_normalize, cmx_search_mixed and cmx_search_prepare: asynchronous (the weights are kernel arguments)
static int mix_on_device(const float* P_d, const float* S_d, int64_t nq, int d, const double* alphas, int nA,
                         float* out_d, uint8_t* flags_d, cudaStream_t st) {
  std::vector<float> w;
  std::vector<int> mode;
  prepare_alphas(alphas, nA, w, mode);
  CMX_NVTX("cmx:mix_normalize");
  return launch_mix_normalize(P_d, S_d, nq, d, w.data(), w.data() + nA, mode.data(), nA, out_d, flags_d, st);
}

int cmx_mix_normalize(const float* P, const float* S, int64_t nq, int d, const double* alphas, int nA, float* out,
                      uint8_t* flags, int io_on_device, int device, void* stream) {
  CMX_CHECK(nq >= 0 && d > 0 && nA >= 0, "bad shape");
  if (nq == 0 || nA == 0) return CMX_OK;
  CMX_CHECK(P && S && alphas && out, "null buffer");
  int ndev = 0;
  CMX_TRY(cmx_device_count(&ndev));
  CMX_CHECK(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
  DevGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t in_bytes = (size_t)nq * d * sizeof(float);
  const size_t out_elems = (size_t)nA * nq * d;
  float *P_d = nullptr, *S_d = nullptr, *out_d = nullptr;
  uint8_t* f_d = nullptr;
  int rc = CMX_OK;
  auto cleanup = [&]() {
    if (!io_on_device) { cudaFree(P_d); cudaFree(S_d); cudaFree(out_d); cudaFree(f_d); }
  };
#define MIX_CUDA(expr)                                                             \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                   \
      cleanup();                                                                   \
      return CMX_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)
  if (io_on_device) {
    P_d = const_cast<float*>(P);
    S_d = const_cast<float*>(S);
    out_d = out;
    f_d = flags;
  } else {
    MIX_CUDA(cudaMalloc((void**)&P_d, in_bytes));
    MIX_CUDA(cudaMalloc((void**)&S_d, in_bytes));
    MIX_CUDA(cudaMalloc((void**)&out_d, out_elems * sizeof(float)));
    MIX_CUDA(cudaMalloc((void**)&f_d, (size_t)nA * nq));
    MIX_CUDA(cudaMemcpyAsync(P_d, P, in_bytes, cudaMemcpyHostToDevice, st));
    MIX_CUDA(cudaMemcpyAsync(S_d, S, in_bytes, cudaMemcpyHostToDevice, st));
  }
  rc = mix_on_device(P_d, S_d, nq, d, alphas, nA, out_d, f_d, st);
  if (rc != CMX_OK) { cleanup(); return rc; }
  if (!io_on_device) {
    MIX_CUDA(cudaMemcpyAsync(out, out_d, out_elems * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (flags) MIX_CUDA(cudaMemcpyAsync(flags, f_d, (size_t)nA * nq, cudaMemcpyDeviceToHost, st));
  }
  MIX_CUDA(cudaStreamSynchronize(st));
#undef MIX_CUDA
  cleanup();
  return CMX_OK;
}

int cmx_search_mixed(cmx_index* ix, const float* P, const float* S, int64_t nq, const double* alphas, int nA, int k,
                     float* D, int64_t* I, uint8_t* flags, int io_on_device, int64_t id_base, int path, void* stream) {
  CMX_CHECK(ix != nullptr, "null index");
  CMX_CHECK(nq >= 0 && nA >= 0, "bad shape");
  CMX_CHECK(k >= 1 && k <= CMX_MAX_K, "k=%d out of range [1, %d]", k, CMX_MAX_K);
  CMX_CHECK(path >= 0 && path <= 2, "bad path selector %d", path);
  if (nq == 0 || nA == 0) return CMX_OK;
  CMX_CHECK(P && S && alphas && D && I, "null buffer");
  DevGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nqt = nq * (int64_t)nA;
  ix->pending = false;
  stats_begin(ix, nqt);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profiling) {
    CMX_TRY(ensure_events(ix));
    e0 = ix->ev[4 * kMaxSlabEvents - 1];
    e1 = ix->ev[4 * kMaxSlabEvents - 2];
    CMX_CUDA(cudaEventRecord(e0, st));
  }
  const float *P_d = P, *S_d = S;
  float* D_d = D;
  int64_t* I_d = I;
  CMX_TRY(ensure_buf(&ix->q_dev, &ix->q_cap, nqt * (int64_t)ix->d));
  CMX_TRY(ensure_buf(&ix->flags_dev, &ix->flags_cap, nqt));
  bool staged_out = false;
  if (!io_on_device) {
    CMX_TRY(ensure_buf(&ix->p_dev, &ix->p_cap, nq * (int64_t)ix->d));
    CMX_TRY(ensure_buf(&ix->s_dev, &ix->s_cap, nq * (int64_t)ix->d));
    CMX_CUDA(cudaMemcpyAsync(ix->p_dev, P, (size_t)nq * ix->d * sizeof(float), cudaMemcpyHostToDevice, st));
    CMX_CUDA(cudaMemcpyAsync(ix->s_dev, S, (size_t)nq * ix->d * sizeof(float), cudaMemcpyHostToDevice, st));
    P_d = ix->p_dev;
    S_d = ix->s_dev;
    if (!mapped_host_outputs(D, I, &D_d, &I_d)) {
      CMX_TRY(ensure_buf(&ix->D_dev, &ix->D_cap, nqt * (int64_t)k));
      CMX_TRY(ensure_buf(&ix->I_dev, &ix->I_cap, nqt * (int64_t)k));
      D_d = ix->D_dev;
      I_d = ix->I_dev;
      staged_out = true;
    }
  }
  uint8_t* f_d = (io_on_device && flags) ? flags : ix->flags_dev;
  CMX_TRY(mix_on_device(P_d, S_d, nq, ix->d, alphas, nA, ix->q_dev, f_d, st));
  CMX_TRY(search_device(ix, ix->q_dev, nqt, k, D_d, I_d, id_base, path, st));
  if (staged_out) {
    CMX_CUDA(cudaMemcpyAsync(D, D_d, (size_t)nqt * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    CMX_CUDA(cudaMemcpyAsync(I, I_d, (size_t)nqt * k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  }
  if (!io_on_device && flags) CMX_CUDA(cudaMemcpyAsync(flags, f_d, (size_t)nqt, cudaMemcpyDeviceToHost, st));
  if (g_profiling) CMX_CUDA(cudaEventRecord(e1, st));
  CMX_CUDA(cudaStreamSynchronize(st));
  if (g_profiling) cudaEventElapsedTime(&ix->stats.total_ms, e0, e1);
  stats_end(ix);
  return CMX_OK;
}

// test hook: the (skip+1)-th cmx_search_begin from now publishes `bits` on top of its own status word
static std::atomic<int> g_inject_skip{-1};
static unsigned g_inject_bits = 0;

// ---- sharded (two-phase) search: asynchronous building blocks ------------------------------------
int cmx_search_prepare(cmx_index* ix, const float* P, const float* S, int64_t nq, const double* alphas, int nA,
                       const float** q_out, void* stream) {
  CMX_CHECK(ix != nullptr && q_out != nullptr, "null argument");
  CMX_CHECK(nq > 0 && nA > 0, "bad shape");
  CMX_CHECK(P && S && alphas, "null buffer");
  DevGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  ix->pending = false;
  const int64_t nqt = nq * (int64_t)nA;
  CMX_TRY(ensure_buf(&ix->q_dev, &ix->q_cap, nqt * (int64_t)ix->d));
  CMX_TRY(ensure_buf(&ix->flags_dev, &ix->flags_cap, nqt));
  CMX_TRY(mix_on_device(P, S, nq, ix->d, alphas, nA, ix->q_dev, ix->flags_dev, st));
  *q_out = ix->q_dev;
  return CMX_OK;
}

int cmx_index_export_bounds(cmx_index* ix, float* out2_dev, void* stream) {
  CMX_CHECK(ix && out2_dev, "null argument");
  DevGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (ix->n > 0 && ix->precision == CMX_PRECISION_RESCORE) CMX_TRY(ensure_planes(ix, false, st));
  return launch_store2(out2_dev, std::max(ix->row_norm_max, ix->norm_floor), std::max(ix->row_resid_max, ix->resid_floor), st);
}

int cmx_search_begin(cmx_index* ix, const float* q, int64_t nq, int k, int64_t id_base, const float* const* bounds_parts,
                     int nparts, float est_scale, float* scores_out, uint32_t* flag_out, void* stream) {
  CMX_CHECK(ix != nullptr, "null index");
  CMX_CHECK(nq > 0 && nq <= kQueryChunk, "two-phase search handles 1..%lld queries per call (got %lld): chunk the batch",
            (long long)kQueryChunk, (long long)nq);
  CMX_CHECK(k >= 1 && k <= CMX_MAX_K, "k=%d out of range [1, %d]", k, CMX_MAX_K);
  CMX_CHECK(q && scores_out && flag_out, "null buffer");
  CMX_CHECK(nparts >= 0 && nparts <= CMX_MAX_PEERS && (nparts == 0 || bounds_parts != nullptr), "bad bounds_parts");
  CMX_CHECK(est_scale > 0.f && est_scale <= 1.f, "est_scale must lie in (0, 1]");
  DevGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  stats_begin(ix, nq);
  ix->stats.path = CMX_PATH_TENSOR;
  ix->pending = false;
  if (!ix->ws.overflow) CMX_CUDA(cudaMalloc((void**)&ix->ws.overflow, sizeof(uint32_t)));
  // a shard that cannot run the one-pass arithmetic (split precision requested, non-finite error bound)
  // says so in its status word: every shard reads every word, so all of them fall back together
  const bool usable = ix->n == 0 || rescore_usable(ix);
  if (!usable || ix->n == 0) {
    CMX_TRY(launch_fill_f32(scores_out, nq * (int64_t)k, CMX_NEG_PAD, st));
    CMX_TRY(launch_publish_flag(nullptr, usable ? 0u : CMX_FLAG_NO_TWO_PHASE, flag_out, st));
    if (usable) {  // empty shard: no candidates, cmx_search_end writes padding
      const int64_t nq_pad = (nq + 127) / 128 * 128;
      CMX_TRY(ensure_buf(&ix->ws.cnt, &ix->cnt_cap, nq_pad));
      CMX_TRY(ensure_buf(&ix->ws.cand, &ix->cand_cap_elems, nq * (int64_t)pick_cap(ix, k)));
      ix->ws.cap = pick_cap(ix, k);
      ix->ws.margin = nullptr;
      CMX_CUDA(cudaMemsetAsync(ix->ws.cnt, 0, (size_t)nq_pad * sizeof(uint32_t), st));
      CMX_CUDA(cudaMemsetAsync(ix->ws.overflow, 0, sizeof(uint32_t), st));
    }
  } else {
    // the margin of every shard comes from the corpus maxima over ALL shards (read in place)
    if (nparts > 0) {
      if (!ix->bounds_dev) CMX_CUDA(cudaMalloc((void**)&ix->bounds_dev, 2 * sizeof(float)));
      CMX_TRY(launch_max_bounds(bounds_parts, nparts, ix->bounds_dev, st));
    }
    ix->bounds_from_dev = nparts > 0;
    const int rc = search_pass(ix, q, nq, k, nullptr, nullptr, id_base, CMX_PATH_TENSOR, true, false, true, st, nullptr, true, est_scale);
    ix->bounds_from_dev = false;
    CMX_TRY(rc);
    CMX_TRY(launch_export_scores(ix->ws, nq, k, scores_out, st));
    unsigned extra = 0u;
    if (g_inject_skip.load() >= 0 && g_inject_skip.fetch_sub(1) == 0) extra = g_inject_bits;
    CMX_TRY(launch_publish_flag(ix->ws.overflow, extra, flag_out, st));
  }
  ix->pending = true;
  ix->pend_skip = !usable;  // flagged: cmx_search_end is a no-op, the step is redone without the two-phase cut
  ix->pend_q = q;
  ix->pend_nq = nq;
  ix->pend_k = k;
  ix->pend_id_base = id_base;
  stats_end(ix);
  return CMX_OK;
}

int cmx_union_kth(const float* const* score_parts, int nparts, int64_t nq, int k, int64_t q0, int64_t q1,
                  float* const* kth_outs, int nouts, const uint32_t* const* flag_parts, int nflags, uint32_t* flag_any,
                  int device, void* stream) {
  CMX_CHECK(score_parts && kth_outs, "null pointer table");
  CMX_CHECK(k >= 1 && k <= CMX_MAX_K, "k=%d out of range [1, %d]", k, CMX_MAX_K);
  CMX_CHECK(q0 >= 0 && q0 <= q1 && q1 <= nq, "bad query slice");
  int ndev = 0;
  CMX_TRY(cmx_device_count(&ndev));
  CMX_CHECK(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
  DevGuard g(device);
  CMX_NVTX("cmx:union_kth");
  return launch_union_kth(score_parts, nparts, k, q0, q1, kth_outs, nouts, flag_parts, nflags, flag_any, (cudaStream_t)stream);
}

int cmx_search_end(cmx_index* ix, const float* const* kth_parts, int nparts, float* D, int64_t* I, void* stream) {
  CMX_CHECK(ix != nullptr && D && I, "null argument");
  CMX_CHECK(ix->pending, "cmx_search_end without a pending cmx_search_begin (a search / add / reset in between cancels it)");
  CMX_CHECK(nparts >= 0 && nparts <= CMX_MAX_PEERS && (nparts == 0 || kth_parts != nullptr), "bad kth_parts");
  DevGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  RescoreCut cut;
  cut.nparts = nparts;
  for (int i = 0; i < nparts; ++i) cut.kth[i] = kth_parts[i];
  const int launches0 = (int)g_launches.load();
  ix->pending = false;
  if (ix->pend_skip) return CMX_OK;
  {
    CMX_NVTX("cmx:rescore");
    CMX_TRY(launch_rescore(ix->X, ix->d, ix->pend_q, ix->ws, ix->pend_nq, ix->pend_k, D, I, ix->pend_id_base, cut, st));
  }
  ix->stats.launches += (int)g_launches.load() - launches0;
  ix->stats.select_launches += 1;
  return CMX_OK;
}

int cmx_peer_broadcast(const void* src, void* const* dsts, int ndst, int64_t bytes, int device, void* stream) {
  CMX_CHECK(bytes >= 0 && ndst >= 0 && ndst <= CMX_MAX_PEERS, "bad argument");
  if (bytes == 0 || ndst == 0) return CMX_OK;
  CMX_CHECK(src && dsts, "null pointer");
  CMX_CHECK((bytes & 15) == 0 && ((uintptr_t)src & 15) == 0, "peer broadcast works on 16-byte aligned blocks");
  for (int i = 0; i < ndst; ++i) CMX_CHECK(dsts[i] && ((uintptr_t)dsts[i] & 15) == 0, "bad destination %d", i);
  DevGuard g(device);
  CMX_NVTX("cmx:peer_broadcast");
  return launch_peer_broadcast(src, dsts, ndst, bytes, (cudaStream_t)stream);
}

int cmx_host_register(void* p, int64_t bytes, void** dev_ptr) {
  CMX_CHECK(p && bytes > 0 && dev_ptr, "bad argument");
  // memory that is already page-locked (cudaHostAlloc, torch pin_memory, an earlier registration) only needs its alias
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost && a.devicePointer != nullptr) {
    *dev_ptr = a.devicePointer;
    return CMX_OK;
  }
  cudaGetLastError();
  CMX_CUDA(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  CMX_CUDA(cudaHostGetDevicePointer(dev_ptr, p, 0));
  return CMX_OK;
}

int cmx_host_unregister(void* p) {
  CMX_CHECK(p != nullptr, "null pointer");
  cudaError_t e = cudaHostUnregister(p);
  if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); e = cudaSuccess; }
  CMX_CUDA(e);
  return CMX_OK;
}

int cmx_enable_peer_access(int device, int peer) {
  int ndev = 0;
  CMX_TRY(cmx_device_count(&ndev));
  CMX_CHECK(device >= 0 && device < ndev && peer >= 0 && peer < ndev, "device %d / peer %d out of range (have %d)", device, peer, ndev);
  if (device == peer) return CMX_OK;
  int can = 0;
  CMX_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
  CMX_CHECK(can, "device %d cannot map the memory of device %d (no NVLink / PCIe peer access)", device, peer);
  DevGuard g(device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  CMX_CUDA(e);
  return CMX_OK;
}

int cmx_merge_topk(const float* D_parts, const int64_t* I_parts, int nparts, int64_t nq, int k, float* D, int64_t* I,
                   int io_on_device, int device, void* stream) {
  CMX_CHECK(nparts >= 1 && nq >= 0, "bad shape");
  CMX_CHECK(k >= 1 && k <= CMX_MAX_K, "k=%d out of range [1, %d]", k, CMX_MAX_K);
  if (nq == 0) return CMX_OK;
  CMX_CHECK(D_parts && I_parts && D && I, "null buffer");
  int ndev = 0;
  CMX_TRY(cmx_device_count(&ndev));
  CMX_CHECK(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
  DevGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (io_on_device) {
    CMX_TRY(launch_merge(D_parts, I_parts, nparts, nq, k, D, I, st));
    CMX_CUDA(cudaStreamSynchronize(st));
    return CMX_OK;
  }
  const size_t pe = (size_t)nparts * nq * k, oe = (size_t)nq * k;
  float *Dp = nullptr, *Do = nullptr;
  int64_t *Ip = nullptr, *Io = nullptr;
  int rc = CMX_OK;
  cudaError_t e = cudaMalloc((void**)&Dp, pe * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&Ip, pe * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&Do, oe * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&Io, oe * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(Dp, D_parts, pe * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(Ip, I_parts, pe * sizeof(int64_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = launch_merge(Dp, Ip, nparts, nq, k, Do, Io, st);
    if (rc == CMX_OK) {
      e = cudaMemcpyAsync(D, Do, oe * sizeof(float), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(I, Io, oe * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
  }
  cudaFree(Dp); cudaFree(Ip); cudaFree(Do); cudaFree(Io);
  if (e != cudaSuccess) { set_error("merge: %s", cudaGetErrorString(e)); return CMX_ERR_CUDA; }
  return rc;
}

int cmx_merge_topk_peers(const float* const* D_parts, const int64_t* const* I_parts, int nparts, int64_t nq, int k,
                         int64_t q0, int64_t q1, float* const* D_outs, int64_t* const* I_outs, int nouts, int device,
                         void* stream) {
  CMX_CHECK(D_parts && I_parts && D_outs && I_outs, "null pointer table");
  CMX_CHECK(k >= 1 && k <= CMX_MAX_K, "k=%d out of range [1, %d]", k, CMX_MAX_K);
  CMX_CHECK(q0 >= 0 && q0 <= q1 && q1 <= nq, "bad query slice [%lld, %lld) of %lld", (long long)q0, (long long)q1, (long long)nq);
  int ndev = 0;
  CMX_TRY(cmx_device_count(&ndev));
  CMX_CHECK(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
  DevGuard g(device);
  return launch_merge_peers(D_parts, I_parts, nparts, k, q0, q1, D_outs, I_outs, nouts, (cudaStream_t)stream);
}

int cmx_collapse_max(const float* D, const int64_t* I, int64_t nq, int k, const int32_t* base_code, int64_t ndocs,
                     int32_t* col_code, int64_t* col_val6, int32_t* col_count, int* needs_host, int device, void* stream) {
  CMX_CHECK(nq >= 0 && k >= 1 && k <= CMX_MAX_K && ndocs >= 0, "bad shape");
  CMX_CHECK(needs_host != nullptr, "null status");
  *needs_host = 0;
  if (nq == 0) return CMX_OK;
  CMX_CHECK(D && I && base_code && col_code && col_val6 && col_count, "null buffer");
  int ndev = 0;
  CMX_TRY(cmx_device_count(&ndev));
  CMX_CHECK(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
  DevGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  CMX_NVTX("cmx:collapse_max");
  uint32_t* status = nullptr;
  CMX_CUDA(cudaMalloc((void**)&status, sizeof(uint32_t)));
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(uint32_t), st);
  int rc = e == cudaSuccess ? launch_collapse_max(D, I, nq, k, base_code, ndocs, col_code, col_val6, col_count, status, st) : CMX_ERR_CUDA;
  uint32_t h = 0;
  if (rc == CMX_OK) {
    e = cudaMemcpyAsync(&h, status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = CMX_ERR_CUDA;
  }
  cudaFree(status);
  if (rc == CMX_ERR_CUDA && e != cudaSuccess) set_error("collapse: %s", cudaGetErrorString(e));
  *needs_host = h ? 1 : 0;
  return rc;
}

/* test hook (not in cmx.h): tensor tile width 256 / 128 */
CMX_API int cmx_debug_set_tensor_tile(int bn) { set_tensor_tile(bn); return CMX_OK; }
CMX_API int cmx_debug_set_block_order(int on) { g_block_order = on ? 1 : 0; return CMX_OK; }
CMX_API int cmx_debug_set_speculate(int on) { g_speculate = on ? 1 : 0; return CMX_OK; }
CMX_API int cmx_debug_set_small_first(int on) { g_small_first = on ? 1 : 0; return CMX_OK; }
CMX_API int cmx_debug_set_mapped_outputs(int on) { g_mapped_outputs = on ? 1 : 0; return CMX_OK; }
/* test hooks (host logic, no GPU needed): the slab schedule of a tensor-path search and the block multiplier */
CMX_API int cmx_debug_plan_slabs(int64_t ntotal, int k, int cap, int rescore, int safe, int speculate, int64_t* rows_out,
                                 int max_slabs, int* nslabs, int* spec_slab, int* spec_rank) {
  CMX_CHECK(ntotal > 0 && k >= 1 && cap >= 2 * k && rows_out && nslabs && spec_slab && spec_rank, "bad argument");
  const int64_t nblk = (ntotal + 255) / 256;
  const int k_plan = rescore ? std::min(cap / 2, k + k / 3 + 8) : k;
  SlabPlan pl = plan_slabs(nblk * 256, k_plan, cap, 256, safe != 0, speculate != 0, k);
  *nslabs = (int)pl.rows.size();
  // the LAST speculative slab and its rank (-1 / 0: none); spec_rank_out of every slab via cmx_debug_plan_ranks
  *spec_slab = -1;
  *spec_rank = 0;
  for (int i = 0; i < *nslabs; ++i)
    if (pl.spec_rank[(size_t)i] > 0) { *spec_slab = i; *spec_rank = pl.spec_rank[(size_t)i]; }
  for (int i = 0; i < *nslabs && i < max_slabs; ++i) rows_out[i] = pl.rows[(size_t)i];
  return CMX_OK;
}
CMX_API int cmx_debug_plan_ranks(int64_t ntotal, int k, int cap, int rescore, int safe, int speculate, int* ranks_out, int max_slabs) {
  CMX_CHECK(ntotal > 0 && k >= 1 && cap >= 2 * k && ranks_out, "bad argument");
  const int64_t nblk = (ntotal + 255) / 256;
  const int k_plan = rescore ? std::min(cap / 2, k + k / 3 + 8) : k;
  SlabPlan pl = plan_slabs(nblk * 256, k_plan, cap, 256, safe != 0, speculate != 0, k);
  for (int i = 0; i < (int)pl.rows.size() && i < max_slabs; ++i) ranks_out[i] = pl.spec_rank[(size_t)i];
  return CMX_OK;
}
CMX_API int cmx_debug_inject_begin_status(int skip, unsigned bits) { g_inject_bits = bits; g_inject_skip.store(skip); return CMX_OK; }
CMX_API int cmx_debug_set_prescore(int mode) { g_prescore = mode; return CMX_OK; }
CMX_API int cmx_debug_set_prescore_min_rows(int64_t rows) { g_prescore_min_rows = rows < 0 ? kPrescoreMinRows : rows; return CMX_OK; }
CMX_API int cmx_debug_set_prescore_params(double depth, int pad_smem_bytes, int max_sub) {
  g_prescore_depth = depth > 0 ? depth : 0.75;
  set_prescore_pad(pad_smem_bytes);
  g_prescore_max_sub = max_sub >= 1 && max_sub <= kMaxSub ? max_sub : kMaxSub;
  return CMX_OK;
}
/* test hook: the APPROXIMATE scores the one-pass tensor scorer produces (rescore precision), for the `nrows` (a multiple
 * of 256, <= the candidate capacity) row positions starting at position `pos0` of the processing order.  q, scores_out
 * [nq, nrows], rows_out [nq, nrows] (corpus row of each score; -1 = padding) and margin_out [nq] (= 2 eps(q)) are device
 * buffers.  Lets a test measure |approx - exact| against eps(q) directly. */
CMX_API int cmx_debug_approx_scores(cmx_index* ix, const float* q, int64_t nq, int64_t pos0, int64_t nrows, float* scores_out,
                                    int64_t* rows_out, float* margin_out, void* stream) {
  CMX_CHECK(ix && q && scores_out && rows_out && margin_out, "null argument");
  CMX_CHECK(nq > 0 && nq <= kQueryChunk, "1..%lld queries", (long long)kQueryChunk);
  CMX_CHECK(ix->n > 0 && rescore_usable(ix), "needs a non-empty index in rescore precision");
  DevGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  ix->pending = false;
  const int k = 1;
  CMX_TRY(prepare_pass(ix, q, nq, k, CMX_PATH_TENSOR, true, st));
  const int cap = ix->ws.cap;
  const int64_t nblk = (ix->n + 255) / 256;
  CMX_CHECK(nrows > 0 && nrows <= cap && (nrows & 255) == 0 && (pos0 & 255) == 0 && pos0 + nrows <= nblk * 256,
            "bad position range [%lld, +%lld) (capacity %d, %lld positions)", (long long)pos0, (long long)nrows, cap, (long long)(nblk * 256));
  const int64_t nq_pad = (nq + 127) / 128 * 128;
  CMX_CUDA(cudaMemsetAsync(ix->ws.cand, 0, (size_t)nq * cap * sizeof(uint64_t), st));
  CMX_TRY(launch_tensor_score(ix->Bhi, ix->Blo, ix->n, pos0, nrows, ix->d_pad, ix->Qhi, ix->Qlo, nq, nq_pad, ix->q_scale + 1,
                              1.0f / ix->plane_scale, ix->ws, 1, pick_perm(nblk), 1, 1.0, ix->progress, st, ix->sm_count));
  for (int64_t qi = 0; qi < nq; ++qi)  // rows of the [nq, cap] buffer -> [nq, nrows]
    if (nrows == cap) { CMX_TRY(launch_decode_keys(ix->ws.cand, nq * (int64_t)cap, scores_out, rows_out, st)); break; }
    else CMX_TRY(launch_decode_keys(ix->ws.cand + qi * cap, nrows, scores_out + qi * nrows, rows_out + qi * nrows, st));
  CMX_CUDA(cudaMemcpyAsync(margin_out, ix->margin_buf, (size_t)nq * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CMX_CUDA(cudaStreamSynchronize(st));
  return CMX_OK;
}
CMX_API uint64_t cmx_debug_block_perm(int64_t nblk) { return pick_perm(nblk); }
CMX_API int cmx_debug_set_tensor_window(int w) { set_tensor_window(w); return CMX_OK; }
CMX_API int cmx_debug_set_tensor_small(int on) { set_tensor_small(on); return CMX_OK; }
CMX_API int cmx_debug_set_tensor_pair(int on) { set_tensor_pair(on); return CMX_OK; }
CMX_API int cmx_debug_set_tensor_flags(int f) { set_tensor_flags(f); return CMX_OK; }
CMX_API int cmx_debug_set_stream_variant(int v) { set_stream_variant(v); return CMX_OK; }

}  // extern "C"

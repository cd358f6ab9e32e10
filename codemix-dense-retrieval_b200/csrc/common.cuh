// Shared helpers for libcmx (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/cmx.h"

namespace cmx {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern int g_profiling;

#define CMX_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      cmx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                     __LINE__);                                                          \
      return (_e == cudaErrorMemoryAllocation) ? CMX_ERR_NOMEM : CMX_ERR_CUDA;           \
    }                                                                                    \
  } while (0)

#define CMX_CHECK(cond, ...)        \
  do {                              \
    if (!(cond)) {                  \
      cmx::set_error(__VA_ARGS__);  \
      return CMX_ERR_INVALID;       \
    }                               \
  } while (0)

#define CMX_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != CMX_OK) return _rc; \
  } while (0)

// after every kernel launch
#define CMX_LAUNCHED()                       \
  do {                                       \
    cmx::g_launches.fetch_add(1);            \
    CMX_CUDA(cudaGetLastError());            \
  } while (0)

// ---- candidate keys ---------------------------------------------------------
// A candidate is one 64-bit key: high word = score mapped to an unsigned that orders like the
// float; low word = bit 31 "score is approximate" | 31 bits of (0x7fffffff - row).  A
// DESCENDING sort of keys with equal flags yields score descending, row ascending.  The flag is
// set by the scoring kernels (fp16 one-pass scores of the rescore precision) and cleared once a
// key carries the exact fp32 score (prescore_kernel / rescore_kernel): selection may mix both
// kinds -- every key's score is within eps(q) of the exact one, which is all the margin proof
// of DESIGN.md 4b needs -- and the final lists hold exact keys only.  Rows are < 2^31 per shard.
// key 0 is below every real candidate.
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
constexpr uint32_t CMX_KEY_APPROX = 0x80000000u;
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return ((uint64_t)f32_to_ordered(score) << 32) | (uint64_t)(CMX_KEY_APPROX | (0x7fffffffu - row));
}
__host__ __device__ __forceinline__ uint64_t make_key_exact(float score, uint32_t row) {
  return ((uint64_t)f32_to_ordered(score) << 32) | (uint64_t)(0x7fffffffu - row);
}
// smallest key carrying `score` (every real key with that score compares >= it)
__host__ __device__ __forceinline__ uint64_t make_key_floor(float score) { return (uint64_t)f32_to_ordered(score) << 32; }
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ordered_to_f32((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return 0x7fffffffu - ((uint32_t)k & 0x7fffffffu); }
__host__ __device__ __forceinline__ bool key_is_exact(uint64_t k) { return ((uint32_t)k & CMX_KEY_APPROX) == 0u; }

#define CMX_NEG_PAD (-3.402823466e+38f) /* FAISS pads IP results with lowest float */

#define CMX_MAX_PEERS 16

// ---- search workspace (device) ----------------------------------------------
// bits of SearchWs::overflow[0]: why an attempt has to be repeated
constexpr unsigned CMX_OVF_BUFFER = 1u;  // more survivors than a candidate buffer holds (row order / bad guess)
constexpr unsigned CMX_OVF_BAND = 2u;    // rescore mode: the margin band of some query outgrew cap/2
constexpr unsigned CMX_OVF_SPEC = 4u;    // a speculative threshold was not cleared by the k-th best
constexpr unsigned CMX_FLAG_NO_TWO_PHASE = 0x100u;  // sharded search: this shard cannot run the one-pass arithmetic

// ---- NVTX ranges (visible in nsys / ncu --nvtx; no-ops without a profiler attached) ------------
struct NvtxRange {
  explicit NvtxRange(const char* name);
  ~NvtxRange();
};
#define CMX_NVTX_CAT2(a, b) a##b
#define CMX_NVTX_CAT(a, b) CMX_NVTX_CAT2(a, b)
#define CMX_NVTX(name) cmx::NvtxRange CMX_NVTX_CAT(_nvtx_, __LINE__)(name)

struct SearchWs {
  float* tau = nullptr;        // [nq_pad] running k-th best score per query (filter threshold)
  uint32_t* cnt = nullptr;     // [nq_pad] candidates appended per query (may exceed cap)
  uint64_t* cand = nullptr;    // [nq_pad, cap] candidate keys
  uint32_t* overflow = nullptr;// [1] set when some query's buffer overflowed
  float* margin = nullptr;     // [nq_pad] rescore mode: 2*eps(q), the slack kept below the k-th best
                               // APPROXIMATE score so that the exact top-k is provably contained; else 0
  float* spec = nullptr;       // [nq_pad] speculative threshold in force for the slab being scored (lowest-float = none)
  float* est = nullptr;        // [nq_pad] rescore mode: candidates scoring at least this are prescored (lowest-float = none)
  int cap = 0;
  int64_t nq_cap = 0;
};

// ---- kernels (launchers return CMX codes) -------------------------------------
constexpr int kMixMaxAlphas = 32;
struct MixParams {
  float w1[kMixMaxAlphas], w2[kMixMaxAlphas];
  int mode[kMixMaxAlphas];
};
// w1 / w2 / mode are HOST arrays [nA] (passed to the kernel by value; asynchronous)
int launch_mix_normalize(const float* P, const float* S, int64_t nq, int d, const float* w1,
                         const float* w2, const int* mode, int nA, float* out, uint8_t* flags,
                         cudaStream_t st);

int launch_absmax(const float* x, int64_t n, uint32_t* absmax_bits, cudaStream_t st);
// hi = f16(x*scale), lo = f16(x*scale - hi); planes have leading dim d_pad (zero padded)
int launch_split_planes(const float* x, int64_t rows, int d, int d_pad, const float* scale_dev,
                        float scale_host, __half* hi, __half* lo, cudaStream_t st);
// scale_out[0] = 2^e with absmax*2^e in [2^12,2^13); scale_out[1] = 1/scale
int launch_scale_from_absmax(const uint32_t* absmax_bits, float* scale_out, cudaStream_t st);
float host_scale_for_absmax_bits(uint32_t bits);
// max over rows of ||x||_2 (finite rows only), as float bits via atomicMax
int launch_row_norm_max(const float* x, int64_t rows, int d, uint32_t* max_bits, cudaStream_t st);
// max over rows of ||x - hi / scale||_2 (finite rows only), as float bits via atomicMax
int launch_row_resid_max(const float* x, const __half* hi, int64_t rows, int d, int d_pad, float inv_scale,
                         uint32_t* max_bits, cudaStream_t st);
// margin[q] = 2 eps(q), eps(q) = ||q - qh|| xmax + (||q|| + ||q - qh||) xres + gamma ||q|| xmax
// bounds_dev (may be NULL): {xmax, xres} over all shards on the device; the larger of each pair is used
int launch_query_margin(const float* Q, const __half* Qhi, int64_t nq, int d, int d_pad, const float* q_scale,
                        float xmax, float xres, const float* bounds_dev, float gamma, float* margin, cudaStream_t st);
// small asynchronous plumbing of the sharded search (prologue.cu)
int launch_store2(float* out2, float a, float b, cudaStream_t st);
int launch_fill_f32(float* out, int64_t n, float v, cudaStream_t st);
int launch_publish_flag(const uint32_t* overflow, uint32_t extra, uint32_t* flag_out, cudaStream_t st);
int launch_max_bounds(const float* const* parts, int nparts, float* out2, cudaStream_t st);
int launch_peer_broadcast(const void* src, void* const* dsts, int ndst, int64_t bytes, cudaStream_t st);
// collapse-by-base-id, fuse = max (collapse.cu): per query the (base code, 6-decimal value) groups in final order
int launch_collapse_max(const float* D, const int64_t* I, int64_t nq, int k, const int32_t* base_code, int64_t ndocs,
                        int32_t* out_code, int64_t* out_val6, int32_t* out_count, uint32_t* status, cudaStream_t st);
int launch_gather_rows(const float* X, const int64_t* rows_dev, int64_t n, int d, float* dst, cudaStream_t st);
int launch_decode_keys(const uint64_t* keys, int64_t n, float* scores, int64_t* rows, cudaStream_t st);

int launch_stream_score(const float* X, int64_t row0, int64_t nrows, int d, const float* Q,
                        int nq, const SearchWs& ws, int64_t q0, int dense, int64_t dense_row0,
                        cudaStream_t st, int sm_count);

void set_stream_variant(int v);  // tuning experiments
int tensor_path_available();
void set_tensor_tile(int bn);  // 256 (default) or 128 corpus rows per tile
void set_tensor_flags(int f);  // tuning experiments (cache hints)
void set_tensor_small(int on);  // corpus-as-M kernel for nq <= 64
void set_tensor_pair(int on);  // CTA-pair (cta_group::2) scorer for nq > 128
// passes = 3: split precision (hi*hi + hi*lo + lo*hi, fp32-faithful scores);
// passes = 1: hi*hi only (approximate filter scores; Blo/Qlo may be NULL) for the rescore mode
// row0 / nrows are row POSITIONS in processing order (whole 256-row blocks); position block j is
// corpus block (j * perm) mod ceil(plane_rows / 256) -- perm = 1 is file order
int launch_tensor_score(const __half* Bhi, const __half* Blo, int64_t plane_rows, int64_t row0,
                        int64_t nrows, int d_pad, const __half* Qhi, const __half* Qlo,
                        int64_t nq, int64_t nq_pad, const float* q_inv_scale_dev, float b_inv_scale,
                        const SearchWs& ws, int dense, uint64_t perm, int passes, double expected_pass_rate,
                        unsigned long long* progress, cudaStream_t st, int sm_count);
void set_tensor_window(int w);  // progress throttle slack in round-robin iterations (0 = off)

int launch_ws_init(const SearchWs& ws, int64_t nq, int64_t nq_pad, cudaStream_t st);
int launch_set_counts(const SearchWs& ws, int64_t nq, uint32_t value, cudaStream_t st);
// sort candidates, keep top-k, refresh tau; final=1 also writes D/I (+id_base, padded)
// spec_rank > 0: additionally publish the speculative threshold of the remaining corpus (the
// spec_rank-th best score so far, minus the margin); verify = 1: the slab just compacted ran under a
// speculative threshold -- flag `overflow` for any query whose k-th best does not clear it
// est_rank > 0: also publish ws.est = the est_rank-th best score so far (prescore_kernel's threshold)
int launch_compact(const SearchWs& ws, int64_t nq, int k, int final_pass, float* D, int64_t* I,
                   int64_t id_base, cudaStream_t st, int spec_rank = 0, int verify = 0, int est_rank = 0);
// snap[q] = cnt[q] (taken between two scoring launches of one slab)
int launch_snapshot_counts(const SearchWs& ws, int64_t nq, uint32_t* snap, cudaStream_t st);
// exact fp32 scores, in place, for the candidates in buffer positions [lo[q], hi[q]) (lo == NULL: from 0)
// that reach ws.est[q]; meant to run on a second stream beside the scoring kernel
constexpr int kPrescoreMaxSmem = 16 * 1024;  // the query copy must fit next to the resident scoring CTA
int prescore_smem_bytes(int d);
void set_prescore_pad(int bytes);
int launch_prescore(const float* X, int d, const float* Q, const SearchWs& ws, int64_t nq, const uint32_t* lo,
                    const uint32_t* hi, cudaStream_t st);
// rescore mode: exact fp32 scores of every surviving candidate from the fp32 row store, then
// the exact top-k (score desc, row asc) -> D, I
struct RescoreCut {
  int nparts = 0;                    // 0: no global cut (single shard)
  const float* kth[CMX_MAX_PEERS];   // per shard: k-th best approximate score per query (device, maybe peer memory)
};
int launch_kth_approx(const SearchWs& ws, int64_t nq, float* out, cudaStream_t st);
int launch_export_scores(const SearchWs& ws, int64_t nq, int k, float* out, cudaStream_t st);
// flags (may be NULL): per-shard status words, OR-ed into *flag_any by block 0
int launch_union_kth(const float* const* parts, int nparts, int k, int64_t q0, int64_t q1, float* const* outs, int nouts,
                     const uint32_t* const* flags, int nflags, uint32_t* flag_any, cudaStream_t st);
int launch_rescore(const float* X, int d, const float* Q, const SearchWs& ws, int64_t nq, int k, float* D,
                   int64_t* I, int64_t id_base, const RescoreCut& cut, cudaStream_t st);
int launch_merge(const float* D_parts, const int64_t* I_parts, int nparts, int64_t nq, int k,
                 float* D, int64_t* I, cudaStream_t st);

int launch_merge_peers(const float* const* D_parts, const int64_t* const* I_parts, int nparts, int k,
                       int64_t q0, int64_t q1, float* const* D_outs, int64_t* const* I_outs, int nouts,
                       cudaStream_t st);

}  // namespace cmx

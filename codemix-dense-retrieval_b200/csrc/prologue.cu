// Vector-mix prologue and fp16 hi/lo operand split (sm_100a).
//
// mix_normalize_kernel restates safe_mix (reference
// onepass_dense_mix_run_custom_lang.py:342-377) for a whole [nA, nq] batch in one
// launch: one warp per output row, HBM-bound (12*d bytes per row).
#include "common.cuh"

namespace cmx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ bool is_finite_f(float v) {
  return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
}

// mode[a]: bits 0-1: 0 = mix, 1 = copy primary, 2 = copy secondary;
//          bits 4-5: fallback row on non-finite output (1 = primary, 2 = secondary)
// The weights travel as kernel arguments (MixParams, up to 32 alphas per launch): nothing is
// staged through device memory, so the launch needs no host synchronisation.
__global__ void __launch_bounds__(256)
mix_normalize_kernel(const float* __restrict__ P, const float* __restrict__ S, int64_t nq, int d,
                     const MixParams mp, int nA, float* __restrict__ out, uint8_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)nA * nq) return;
  const int a = (int)(row / nq);
  const int64_t q = row - (int64_t)a * nq;
  const float* p = P + q * d;
  const float* s = S + q * d;
  float* o = out + row * d;
  const int mode = mp.mode[a];
  const int sel = mode & 3;
  uint8_t flag = 0;
  const bool vec = (d & 3) == 0;

  if (sel == 0) {
    const float w1 = mp.w1[a], w2 = mp.w2[a];
    float ss = 0.f;
    if (vec) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      const float4* s4 = reinterpret_cast<const float4*>(s);
      for (int i = lane; i < (d >> 2); i += 32) {
        float4 a4 = p4[i], b4 = s4[i];
        float m0 = __fadd_rn(__fmul_rn(w1, a4.x), __fmul_rn(w2, b4.x));
        float m1 = __fadd_rn(__fmul_rn(w1, a4.y), __fmul_rn(w2, b4.y));
        float m2 = __fadd_rn(__fmul_rn(w1, a4.z), __fmul_rn(w2, b4.z));
        float m3 = __fadd_rn(__fmul_rn(w1, a4.w), __fmul_rn(w2, b4.w));
        ss = fmaf(m0, m0, ss); ss = fmaf(m1, m1, ss); ss = fmaf(m2, m2, ss); ss = fmaf(m3, m3, ss);
      }
    } else {
      for (int i = lane; i < d; i += 32) {
        float m = __fadd_rn(__fmul_rn(w1, p[i]), __fmul_rn(w2, s[i]));
        ss = fmaf(m, m, ss);
      }
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    const bool denom_nan = (ss != ss);
    bool bad = denom_nan;
    if (vec) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      const float4* s4 = reinterpret_cast<const float4*>(s);
      float4* o4 = reinterpret_cast<float4*>(o);
      for (int i = lane; i < (d >> 2); i += 32) {
        float4 a4 = p4[i], b4 = s4[i], r;
        r.x = __fdiv_rn(__fadd_rn(__fmul_rn(w1, a4.x), __fmul_rn(w2, b4.x)), denom);
        r.y = __fdiv_rn(__fadd_rn(__fmul_rn(w1, a4.y), __fmul_rn(w2, b4.y)), denom);
        r.z = __fdiv_rn(__fadd_rn(__fmul_rn(w1, a4.z), __fmul_rn(w2, b4.z)), denom);
        r.w = __fdiv_rn(__fadd_rn(__fmul_rn(w1, a4.w), __fmul_rn(w2, b4.w)), denom);
        bad |= !(is_finite_f(r.x) && is_finite_f(r.y) && is_finite_f(r.z) && is_finite_f(r.w));
        o4[i] = r;
      }
    } else {
      for (int i = lane; i < d; i += 32) {
        float r = __fdiv_rn(__fadd_rn(__fmul_rn(w1, p[i]), __fmul_rn(w2, s[i])), denom);
        bad |= !is_finite_f(r);
        o[i] = r;
      }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (bad) {
      const int fb = (mode >> 4) & 3;
      const float* src = (fb == 2) ? s : p;
      __syncwarp();
      for (int i = lane; i < d; i += 32) o[i] = src[i];
      flag = (uint8_t)fb;
    }
  } else {
    const float* src = (sel == 2) ? s : p;
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* o4 = reinterpret_cast<float4*>(o);
      for (int i = lane; i < (d >> 2); i += 32) o4[i] = s4[i];
    } else {
      for (int i = lane; i < d; i += 32) o[i] = src[i];
    }
  }
  if (flags != nullptr && lane == 0) flags[row] = flag;
}

int launch_mix_normalize(const float* P, const float* S, int64_t nq, int d, const float* w1,
                         const float* w2, const int* mode, int nA, float* out, uint8_t* flags,
                         cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  const int warps = 8;
  for (int a0 = 0; a0 < nA; a0 += kMixMaxAlphas) {
    const int na = nA - a0 < kMixMaxAlphas ? nA - a0 : kMixMaxAlphas;
    MixParams mp;
    for (int a = 0; a < na; ++a) { mp.w1[a] = w1[a0 + a]; mp.w2[a] = w2[a0 + a]; mp.mode[a] = mode[a0 + a]; }
    const int64_t rows = (int64_t)na * nq;
    const int64_t blocks = (rows + warps - 1) / warps;
    mix_normalize_kernel<<<(unsigned)blocks, warps * 32, 0, st>>>(P, S, nq, d, mp, na, out + (int64_t)a0 * nq * d,
                                                                  flags ? flags + (int64_t)a0 * nq : nullptr);
    CMX_LAUNCHED();
  }
  return CMX_OK;
}

// ---- absmax over finite values ------------------------------------------------
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ out_bits) {
  uint32_t m = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t b = __float_as_uint(x[i]) & 0x7fffffffu;
    if ((b & 0x7f800000u) != 0x7f800000u) m = max(m, b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(out_bits, m);
}

int launch_absmax(const float* x, int64_t n, uint32_t* absmax_bits, cudaStream_t st) {
  if (n == 0) return CMX_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, absmax_bits);
  CMX_LAUNCHED();
  return CMX_OK;
}

// scale = 2^(12 - floor(log2(absmax))) so that absmax*scale lies in [2^12, 2^13):
// 8x head-room below the fp16 maximum, and the lo parts of typical elements stay
// fp16-normal.  scale_out = {scale, 1/scale}
__host__ __device__ inline float scale_for_absmax_bits(uint32_t bits) {
  int e = (int)((bits >> 23) & 0xffu);
  if (bits == 0u || e == 0) return 1.0f;  // all-zero (or denormal-only) data
  int se = 12 - (e - 127);
  if (se > 100) se = 100;
  if (se < -100) se = -100;
  uint32_t sb = (uint32_t)(se + 127) << 23;
#ifdef __CUDA_ARCH__
  return __uint_as_float(sb);
#else
  union { float f; uint32_t u; } c; c.u = sb; return c.f;
#endif
}

__global__ void scale_from_absmax_kernel(const uint32_t* __restrict__ bits, float* __restrict__ out) {
  float s = scale_for_absmax_bits(bits[0]);
  out[0] = s;
  out[1] = 1.0f / s;
}

int launch_scale_from_absmax(const uint32_t* absmax_bits, float* scale_out, cudaStream_t st) {
  scale_from_absmax_kernel<<<1, 1, 0, st>>>(absmax_bits, scale_out);
  CMX_LAUNCHED();
  return CMX_OK;
}

float host_scale_for_absmax_bits(uint32_t bits) { return scale_for_absmax_bits(bits); }

// ---- row norms (rescore mode error bound) ----------------------------------------
// one warp per row; rows holding inf / NaN elements are ignored
__global__ void __launch_bounds__(256)
row_norm_max_kernel(const float* __restrict__ x, int64_t rows, int d, uint32_t* __restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float best = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* row = x + r * d;
    float ss = 0.f;
    bool finite = true;
    for (int i = lane; i < d; i += 32) { const float v = row[i]; finite &= is_finite_f(v); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    finite = __all_sync(0xffffffffu, finite);
    const float nrm = sqrtf(ss);
    if (is_finite_f(nrm)) best = fmaxf(best, nrm);
    // a FINITE row whose fp32 sum of squares overflows (||x|| > 1.8e19): the error bound of the
    // one-pass scorer cannot be stated in fp32 -- report +inf, the index then uses split precision
    else if (finite) best = __int_as_float(0x7f800000);
  }
  if (lane == 0 && best > 0.f) atomicMax(max_bits, __float_as_uint(best));
}

int launch_row_norm_max(const float* x, int64_t rows, int d, uint32_t* max_bits, cudaStream_t st) {
  if (rows == 0) return CMX_OK;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  row_norm_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, rows, d, max_bits);
  CMX_LAUNCHED();
  return CMX_OK;
}

// max over rows of || x - hi / scale ||_2: what the one-pass scorer loses of a corpus row (fp16
// rounding, underflow of tiny elements -- whatever the cause, this is the true residual)
__global__ void __launch_bounds__(256)
row_resid_max_kernel(const float* __restrict__ x, const __half* __restrict__ hi, int64_t rows, int d, int d_pad,
                     float inv_scale, uint32_t* __restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float best = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* row = x + r * d;
    const __half* hrow = hi + r * d_pad;
    float ss = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float v = row[i] - __half2float(hrow[i]) * inv_scale;  // exact: scale is a power of two
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    if (is_finite_f(nrm)) best = fmaxf(best, nrm);
  }
  if (lane == 0 && best > 0.f) atomicMax(max_bits, __float_as_uint(best));
}

int launch_row_resid_max(const float* x, const __half* hi, int64_t rows, int d, int d_pad, float inv_scale,
                         uint32_t* max_bits, cudaStream_t st) {
  if (rows == 0) return CMX_OK;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  row_resid_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, hi, rows, d, d_pad, inv_scale, max_bits);
  CMX_LAUNCHED();
  return CMX_OK;
}

// Rescore-mode margin.  With q = qh + ql and x = xh + xl (qh, xh the fp16 planes in true units):
//   q.x - qh.xh = ql.x + qh.xl,   |ql.x + qh.xl| <= ||ql|| ||x|| + ||qh|| ||xl||     (Cauchy-Schwarz)
// and the fp32 accumulation of the tensor core plus that of the exact rescoring chain adds at
// most gamma ||q|| ||x||.  eps(q) = ||ql|| X + (||q|| + ||ql||) R + gamma ||q|| X with X, R the
// corpus maxima of ||x|| and ||xl||; margin = 2 eps.  The residual norms are measured, not
// bounded by 2^-11 ||.||: typical fp16 rounding loses ~0.4 of the worst case.
__global__ void __launch_bounds__(256)
query_margin_kernel(const float* __restrict__ Q, const __half* __restrict__ Qhi, int64_t nq, int d, int d_pad,
                    const float* __restrict__ q_scale, float xmax, float xres, const float* __restrict__ bounds,
                    float gamma, float* __restrict__ margin) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  const float* row = Q + q * d;
  const __half* hrow = Qhi + q * d_pad;
  const float inv_scale = q_scale[1];
  if (bounds != nullptr) {  // sharded search: corpus maxima over ALL shards, reduced on the device
    xmax = fmaxf(xmax, bounds[0]);
    xres = fmaxf(xres, bounds[1]);
  }
  float ss = 0.f, rr = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = row[i];
    const float r = v - __half2float(hrow[i]) * inv_scale;
    ss = fmaf(v, v, ss);
    rr = fmaf(r, r, rr);
  }
  ss = warp_sum(ss);
  rr = warp_sum(rr);
  // a non-finite query gives a non-finite margin: every score passes the filter, the buffers
  // overflow and the search falls back to the split-precision path
  if (lane == 0) {
    const float nq2 = sqrtf(ss) * 1.001f, rq = sqrtf(rr) * 1.001f;  // 1.001: rounding of these very sums
    margin[q] = 2.0f * (rq * xmax + (nq2 + rq) * xres + gamma * nq2 * xmax) * 1.001f;
  }
}

int launch_query_margin(const float* Q, const __half* Qhi, int64_t nq, int d, int d_pad, const float* q_scale,
                        float xmax, float xres, const float* bounds_dev, float gamma, float* margin, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  query_margin_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(Q, Qhi, nq, d, d_pad, q_scale, xmax, xres, bounds_dev, gamma, margin);
  CMX_LAUNCHED();
  return CMX_OK;
}

// ---- fp32 -> fp16 hi/lo planes --------------------------------------------------
// x*scale = hi + lo + O(2^-22 |x*scale|) with hi, lo fp16; scale is a power of two.
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ x, int64_t rows, int d, int d_pad,
                    const float* __restrict__ scale_dev, float scale_host,
                    __half* __restrict__ hi, __half* __restrict__ lo) {
  const float scale = scale_dev ? scale_dev[0] : scale_host;
  const int groups = d_pad >> 3;  // 8 elements per thread
  const int64_t total = rows * (int64_t)groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / groups;
    const int c0 = (int)(t - r * groups) << 3;
    float v[8];
    const float* src = x + r * d + c0;
    if ((d & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + 4 * j < d) f = *reinterpret_cast<const float4*>(src + 4 * j);
        v[4 * j + 0] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < d) ? src[j] : 0.f;
    }
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sv = v[j] * scale;
      __half hh = __float2half_rn(sv);
      float res = sv - __half2float(hh);
      // inf/NaN inputs: keep them in hi, zero in lo (inf - inf would poison lo with NaN anyway)
      h[j] = hh;
      l[j] = __float2half_rn(res);
    }
    *reinterpret_cast<uint4*>(hi + r * d_pad + c0) = *reinterpret_cast<uint4*>(h);
    if (lo != nullptr) *reinterpret_cast<uint4*>(lo + r * d_pad + c0) = *reinterpret_cast<uint4*>(l);
  }
}

int launch_split_planes(const float* x, int64_t rows, int d, int d_pad, const float* scale_dev,
                        float scale_host, __half* hi, __half* lo, cudaStream_t st) {
  if (rows == 0) return CMX_OK;
  const int64_t total = rows * (int64_t)(d_pad >> 3);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  split_planes_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, rows, d, d_pad, scale_dev, scale_host, hi, lo);
  CMX_LAUNCHED();
  return CMX_OK;
}

// ---- small device-side plumbing of the sharded search (all asynchronous) ---------------------------
__global__ void store2_kernel(float* out, float a, float b) { out[0] = a; out[1] = b; }
int launch_store2(float* out2, float a, float b, cudaStream_t st) {
  store2_kernel<<<1, 1, 0, st>>>(out2, a, b);
  CMX_LAUNCHED();
  return CMX_OK;
}

__global__ void fill_f32_kernel(float* out, int64_t n, float v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}
int launch_fill_f32(float* out, int64_t n, float v, cudaStream_t st) {
  if (n <= 0) return CMX_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fill_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, n, v);
  CMX_LAUNCHED();
  return CMX_OK;
}

// flag_out[0] = (overflow ? overflow[0] : 0) | extra  -- the shard's status word of this step
__global__ void publish_flag_kernel(const uint32_t* overflow, uint32_t extra, uint32_t* flag_out) {
  flag_out[0] = (overflow ? overflow[0] : 0u) | extra;
}
int launch_publish_flag(const uint32_t* overflow, uint32_t extra, uint32_t* flag_out, cudaStream_t st) {
  publish_flag_kernel<<<1, 1, 0, st>>>(overflow, extra, flag_out);
  CMX_LAUNCHED();
  return CMX_OK;
}

// out2 = element-wise max over the shards' {max row norm, max residual norm} (peer reads)
struct BoundsArgs { const float* parts[CMX_MAX_PEERS]; };
__global__ void max_bounds_kernel(const BoundsArgs a, int nparts, float* out2) {
  float x = 0.f, r = 0.f;
  for (int g = 0; g < nparts; ++g) {
    // NaN-safe: a non-finite bound must survive the reduction (the margins then become non-finite,
    // everything passes the filter, the buffers overflow and the step falls back)
    const float xg = a.parts[g][0], rg = a.parts[g][1];
    x = (xg != xg || xg > x) ? xg : x;
    r = (rg != rg || rg > r) ? rg : r;
  }
  out2[0] = x;
  out2[1] = r;
}
int launch_max_bounds(const float* const* parts, int nparts, float* out2, cudaStream_t st) {
  BoundsArgs a;
  for (int g = 0; g < nparts; ++g) a.parts[g] = parts[g];
  max_bounds_kernel<<<1, 1, 0, st>>>(a, nparts, out2);
  CMX_LAUNCHED();
  return CMX_OK;
}

// one read of src, one 128-bit store into each destination (peer memory over NVLink)
struct BcastArgs { uint4* dst[CMX_MAX_PEERS]; };
__global__ void __launch_bounds__(256) peer_broadcast_kernel(const uint4* __restrict__ src, const BcastArgs a, int ndst, int64_t n16) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = src[i];
    for (int g = 0; g < ndst; ++g) a.dst[g][i] = v;
  }
}
int launch_peer_broadcast(const void* src, void* const* dsts, int ndst, int64_t bytes, cudaStream_t st) {
  BcastArgs a;
  for (int g = 0; g < ndst; ++g) a.dst[g] = reinterpret_cast<uint4*>(dsts[g]);
  const int64_t n16 = bytes / 16;
  int64_t blocks = (n16 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  peer_broadcast_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), a, ndst, n16);
  CMX_LAUNCHED();
  return CMX_OK;
}

// dst[i, :] = X[rows[i], :]  (one warp per row, 128-bit when d % 4 == 0)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ X, const int64_t* __restrict__ rows, int64_t n,
                                                          int d, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
    const float* s = X + rows[i] * d;
    float* o = dst + i * d;
    if ((d & 3) == 0) {
      for (int j = lane; j < (d >> 2); j += 32) reinterpret_cast<float4*>(o)[j] = reinterpret_cast<const float4*>(s)[j];
    } else {
      for (int j = lane; j < d; j += 32) o[j] = s[j];
    }
  }
}
int launch_gather_rows(const float* X, const int64_t* rows_dev, int64_t n, int d, float* dst, cudaStream_t st) {
  if (n <= 0) return CMX_OK;
  int64_t blocks = (n + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, rows_dev, n, d, dst);
  CMX_LAUNCHED();
  return CMX_OK;
}

// test hook: candidate keys -> (score, row)
__global__ void decode_keys_kernel(const uint64_t* __restrict__ keys, int64_t n, float* __restrict__ scores, int64_t* __restrict__ rows) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = keys[i];
    scores[i] = k ? key_score(k) : CMX_NEG_PAD;
    rows[i] = k ? (int64_t)key_row(k) : -1;
  }
}
int launch_decode_keys(const uint64_t* keys, int64_t n, float* scores, int64_t* rows, cudaStream_t st) {
  if (n <= 0) return CMX_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  decode_keys_kernel<<<(unsigned)blocks, 256, 0, st>>>(keys, n, scores, rows);
  CMX_LAUNCHED();
  return CMX_OK;
}

}  // namespace cmx

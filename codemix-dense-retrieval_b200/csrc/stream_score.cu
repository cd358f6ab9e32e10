// CUDA-core fp32 streaming scorer for small query batches (HBM-bound).
//
// Replaces the nq<20 SIMD path / small-batch GPU path behind index.search
// (reference: onepass_dense_run.py:427,460 issue nq=1 searches; the vector-mix
// scripts issue per-qblock searches, onepass_dense_mix_run_custom_lang.py:878).
//
// Layout: the fp32 row store X [N, d] is read exactly once per group of up to 8
// queries with 128-bit non-allocating loads; the queries sit in shared memory;
// every warp owns R consecutive rows per iteration and keeps R*B fp32
// accumulators; a transposing butterfly reduces them so that one lane ends up
// with each (row, query) score.  Scores are exact fp32 FMA chains.  A score that
// beats the query's threshold tau is appended to the query's candidate buffer
// (select.cu turns the buffers into top-k).  Algorithmic bytes per launch:
// 4*nrows*d (+ 4*B*d for the queries).
#include "common.cuh"

namespace cmx {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

constexpr int kStreamThreads = 256;

// B queries per pass, R rows per warp iteration, U 128-bit loads per row in flight,
// MINB resident CTAs per SM the register budget is sized for
template <int B, int R, int U, int MINB>
__global__ void __launch_bounds__(kStreamThreads, MINB)
stream_score_kernel(const float* __restrict__ X, int64_t row0, int64_t nrows, int d,
                    const float* __restrict__ Q, int nq, const float* __restrict__ tau,
                    uint32_t* __restrict__ cnt, uint64_t* __restrict__ cand, int cap, int dense,
                    int64_t dense_row0) {
  extern __shared__ __align__(16) float Qs[];  // [B][d]
  constexpr int NV = R * B;
  static_assert(NV <= 32 && (NV & (NV - 1)) == 0, "R*B must be a power of two <= 32");
  const int lane = threadIdx.x & 31;
  const int d4 = d >> 2;

  for (int i = threadIdx.x; i < B * d4; i += blockDim.x) {
    const int qi = i / d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qi < nq) v = reinterpret_cast<const float4*>(Q)[i];
    reinterpret_cast<float4*>(Qs)[i] = v;
  }
  __syncthreads();

  // which (row-in-group, query) this lane holds after the butterfly
  int vi = 0;
  {
    int n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      if (n > 1) {
        n >>= 1;
        if (lane & off) vi += n;
      }
    }
  }
  constexpr int kPlainSteps = (NV == 32) ? 0 : (NV == 16) ? 1 : (NV == 8) ? 2 : (NV == 4) ? 3 : (NV == 2) ? 4 : 5;
  const bool emitter = (lane & ((1 << kPlainSteps) - 1)) == 0;
  const int my_r = vi / B;
  const int my_q = vi % B;
  const float my_tau = (my_q < nq) ? tau[my_q] : __int_as_float(0x7f800000);

  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t groups = (nrows + R - 1) / R;
  const float4* Qs4 = reinterpret_cast<const float4*>(Qs);

  for (int64_t g = warp_id; g < groups; g += warps_total) {
    const float4* xr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int64_t rr = g * R + r;
      if (rr >= nrows) rr = nrows - 1;  // clamp; result discarded below
      xr[r] = reinterpret_cast<const float4*>(X + (row0 + rr) * (int64_t)d);
    }
    // packed fp32 FMA (FFMA2, sm_100): two independent partial sums per (row, query)
    float2 acc2[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc2[i] = make_float2(0.f, 0.f);

    for (int j0 = 0; j0 < d4; j0 += 32 * U) {
      float4 x[R][U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j0 + u * 32 + lane;
#pragma unroll
        for (int r = 0; r < R; ++r)
          x[r][u] = (jj < d4) ? ldg_stream(xr[r] + jj) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int jj = j0 + u * 32 + lane;
        if (jj >= d4) jj = d4 - 1;  // x is zero there; keep the smem read in bounds
#pragma unroll
        for (int qi = 0; qi < B; ++qi) {
          const float4 q4 = Qs4[qi * d4 + jj];
          const float2 qa = make_float2(q4.x, q4.y), qb = make_float2(q4.z, q4.w);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float2 a = acc2[r * B + qi];
            a = __ffma2_rn(make_float2(x[r][u].x, x[r][u].y), qa, a);
            a = __ffma2_rn(make_float2(x[r][u].z, x[r][u].w), qb, a);
            acc2[r * B + qi] = a;
          }
        }
      }
    }
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = acc2[i].x + acc2[i].y;

    // transposing butterfly: NV partial sums x 32 lanes -> one total per lane
    {
      int n = NV;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
          const int half = n >> 1;
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const float send = upper ? acc[i] : acc[i + half];
            const float keep = upper ? acc[i + half] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
          n = half;
        } else {
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
        }
      }
    }

    const int64_t rr = g * R + my_r;
    if (emitter && rr < nrows && my_q < nq) {
      const float s = acc[0];
      const int64_t grow = row0 + rr;
      if (dense) {
        // first slab: every row is a candidate; rows that could never enter a FAISS
        // heap (NaN, -inf, lowest float) are stored as the null key
        const uint64_t key = (s > CMX_NEG_PAD) ? make_key(s, (uint32_t)grow) : 0ull;
        cand[(int64_t)my_q * cap + (grow - dense_row0)] = key;
      } else if (s > my_tau) {
        const uint32_t pos = atomicAdd(&cnt[my_q], 1u);
        if (pos < (uint32_t)cap) cand[(int64_t)my_q * cap + pos] = make_key(s, (uint32_t)grow);
      }
    }
  }
}

static int g_stream_variant = 2;  // B=8: U=4 loads in flight, 2 CTAs/SM (no spills); measured best
void set_stream_variant(int v) { g_stream_variant = v; }

template <int B, int R, int U, int MINB>
static int launch_one(const float* X, int64_t row0, int64_t nrows, int d, const float* Q, int nq,
                      const SearchWs& ws, int64_t q0, int dense, int64_t dense_row0,
                      cudaStream_t st, int sm_count) {
  const size_t smem = (size_t)B * d * sizeof(float);
  CMX_CUDA(cudaFuncSetAttribute(stream_score_kernel<B, R, U, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t groups = (nrows + R - 1) / R;
  const int warps_per_block = kStreamThreads / 32;
  int64_t blocks = (groups + warps_per_block - 1) / warps_per_block;
  const int64_t max_blocks = (int64_t)sm_count * MINB;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  stream_score_kernel<B, R, U, MINB><<<(unsigned)blocks, kStreamThreads, smem, st>>>(
      X, row0, nrows, d, Q, nq, ws.tau + q0, ws.cnt + q0, ws.cand + q0 * (int64_t)ws.cap, ws.cap,
      dense, dense_row0);
  CMX_LAUNCHED();
  return CMX_OK;
}

// Scores rows [row0, row0+nrows) of X against queries Q[0..nq) (nq <= 8 per launch;
// larger nq is processed in groups, each group re-streaming the slab).
int launch_stream_score(const float* X, int64_t row0, int64_t nrows, int d, const float* Q,
                        int nq, const SearchWs& ws, int64_t q0, int dense, int64_t dense_row0,
                        cudaStream_t st, int sm_count) {
  if (nrows <= 0 || nq <= 0) return CMX_OK;
  CMX_CHECK((d & 3) == 0, "stream path needs d %% 4 == 0 (got d=%d)", d);
  int bmax = 8;
  while (bmax > 1 && (size_t)bmax * d * sizeof(float) > 64 * 1024) bmax >>= 1;
  CMX_CHECK((size_t)bmax * d * sizeof(float) <= 200 * 1024, "d=%d too large for the stream path", d);
  for (int g0 = 0; g0 < nq; g0 += bmax) {
    const int b = (nq - g0 < bmax) ? (nq - g0) : bmax;
    const float* Qg = Q + (int64_t)g0 * d;
    const int64_t qq = q0 + g0;
    if (b > 4) {
      if (g_stream_variant == 3) CMX_TRY((launch_one<8, 4, 2, 2>(X, row0, nrows, d, Qg, b, ws, qq, dense, dense_row0, st, sm_count)));
      else if (g_stream_variant == 1) CMX_TRY((launch_one<8, 2, 2, 3>(X, row0, nrows, d, Qg, b, ws, qq, dense, dense_row0, st, sm_count)));
      else if (g_stream_variant == 2) CMX_TRY((launch_one<8, 2, 4, 2>(X, row0, nrows, d, Qg, b, ws, qq, dense, dense_row0, st, sm_count)));
      else CMX_TRY((launch_one<8, 2, 4, 3>(X, row0, nrows, d, Qg, b, ws, qq, dense, dense_row0, st, sm_count)));
    } else if (b > 2) CMX_TRY((launch_one<4, 2, 4, 3>(X, row0, nrows, d, Qg, b, ws, qq, dense, dense_row0, st, sm_count)));
    else if (b > 1) CMX_TRY((launch_one<2, 4, 4, 3>(X, row0, nrows, d, Qg, b, ws, qq, dense, dense_row0, st, sm_count)));
    else CMX_TRY((launch_one<1, 4, 4, 3>(X, row0, nrows, d, Qg, b, ws, qq, dense, dense_row0, st, sm_count)));
  }
  return CMX_OK;
}

}  // namespace cmx

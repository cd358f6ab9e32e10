// k-selection: candidate-buffer compaction (radix select + small bitonic sort in shared
// memory), final (D, I) emission and the k-way merge of per-shard results.
//
// Replaces FAISS's blockSelect / heap pass behind index.search
// (reference call sites onepass_dense_mix_run_custom_lang.py:878,
// onepass_bilingual_mix_hub_custom_lang.py:950).  The scoring kernels never
// materialise the score matrix: they append (score,row) keys that beat the
// query's running threshold tau to a per-query buffer; this file turns the
// buffers into exact, deterministically ordered top-k lists.
//
// Keys are distinct 64-bit integers (score-ordered high word, ~row low word), so the
// k-th largest key is found exactly by an MSB-first radix select (8-bit digits, shared
// memory histogram, digits above the first bit where min and max differ are skipped);
// only the k survivors are ever sorted, and only in the final pass.
#include "common.cuh"

namespace cmx {

constexpr int kSelThreads = 512;

__global__ void ws_init_kernel(float* tau, float* spec, uint32_t* cnt, uint32_t* overflow, int64_t nq, int64_t nq_pad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) overflow[0] = 0;
  if (i < nq_pad) {
    // FAISS's IP heap starts at lowest-float and admits only strictly larger scores;
    // padded (non-existent) queries get +inf so nothing ever passes the filter.
    tau[i] = (i < nq) ? CMX_NEG_PAD : __int_as_float(0x7f800000);
    if (spec) spec[i] = CMX_NEG_PAD;
    cnt[i] = 0;
  }
}

int launch_ws_init(const SearchWs& ws, int64_t nq, int64_t nq_pad, cudaStream_t st) {
  int64_t blocks = (nq_pad + 255) / 256;
  if (blocks < 1) blocks = 1;
  ws_init_kernel<<<(unsigned)blocks, 256, 0, st>>>(ws.tau, ws.spec, ws.cnt, ws.overflow, nq, nq_pad);
  CMX_LAUNCHED();
  return CMX_OK;
}

__global__ void set_counts_kernel(uint32_t* cnt, int64_t nq, uint32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) cnt[i] = v;
}

int launch_set_counts(const SearchWs& ws, int64_t nq, uint32_t value, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  set_counts_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(ws.cnt, nq, value);
  CMX_LAUNCHED();
  return CMX_OK;
}

// in-place descending bitonic sort of keys[0..P) (P a power of two) by the whole CTA
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

struct SelectShared {
  uint32_t hist[256];
  uint64_t red_min[32];
  uint64_t red_max[32];
  uint64_t kmin, kmax;
  uint32_t bin, above, count, bin_count;
  uint64_t found;
};

// Exact k-th largest of the n DISTINCT keys in smem (1 <= k < n).  All threads of the
// CTA must call; returns the same value to every thread.
__device__ uint64_t block_kth_largest(const uint64_t* keys, int n, int k, SelectShared& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint64_t mn = ~0ull, mx = 0ull;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t v = keys[i];
    mn = min(mn, v);
    mx = max(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { sh.red_min[warp] = mn; sh.red_max[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = (lane < nwarps) ? sh.red_min[lane] : ~0ull;
    mx = (lane < nwarps) ? sh.red_max[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { sh.kmin = mn; sh.kmax = mx; }
  }
  __syncthreads();
  const uint64_t kmin = sh.kmin, kmax = sh.kmax;
  const uint64_t diff = kmin ^ kmax;
  if (diff == 0ull) return kmax;  // all keys equal: only possible for the null key
  const int top_byte = (63 - __clzll((long long)diff)) >> 3;
  // digits above top_byte are common to all keys
  uint64_t mask = (top_byte == 7) ? 0ull : (~0ull << ((top_byte + 1) * 8));
  uint64_t prefix = kmax & mask;
  uint32_t want = (uint32_t)k;  // rank (1 = largest) among the keys matching the prefix
  for (int shift = top_byte * 8; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh.hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t v = keys[i];
      if ((v & mask) == prefix) atomicAdd(&sh.hist[(uint32_t)(v >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns digits [8l, 8l+8); find the digit where the count from the top crosses `want`
      uint32_t h[8], s = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { h[j] = sh.hist[lane * 8 + j]; s += h[j]; }
      uint32_t incl = s;  // inclusive suffix sum over lanes (lane 31 = highest digits)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += t;
      }
      uint32_t running = incl - s;  // keys in digits above this lane's range
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (running < want && want <= running + h[j]) { sh.bin = lane * 8 + j; sh.above = running; sh.bin_count = h[j]; }
        running += h[j];
      }
    }
    __syncthreads();
    want -= sh.above;
    prefix |= (uint64_t)sh.bin << shift;
    mask |= 0xffull << shift;
    const bool unique = sh.bin_count == 1u;
    __syncthreads();
    if (unique && shift > 0) {
      // a single key carries this prefix: it is the k-th largest, no need to resolve lower digits
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t v = keys[i];
        if ((v & mask) == prefix) sh.found = v;
      }
      __syncthreads();
      return sh.found;
    }
  }
  return prefix;
}

// Appends the survivors to dst (smem or global), order arbitrary, and returns how many:
// the keys >= kth (exactly k of them when kth is the k-th largest: real keys are distinct;
// more when kth is a lowered cut-off), or -- when kth is the null key, i.e. fewer than k
// real candidates exist -- all non-null keys.
__device__ int block_partition(const uint64_t* keys, int n, uint64_t kth, uint64_t* dst, SelectShared& sh) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) sh.count = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const uint64_t v = (i < n) ? keys[i] : 0ull;
    const bool keep = (i < n) && ((kth != 0ull) ? (v >= kth) : (v != 0ull));
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    uint32_t wbase = 0;
    if (lane == 0 && ballot) wbase = atomicAdd(&sh.count, (uint32_t)__popc(ballot));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (keep) dst[wbase + __popc(ballot & ((1u << lane) - 1u))] = v;
  }
  __syncthreads();
  return (int)sh.count;
}

// One CTA per query.  Keeps the query's best k candidates at the front of its buffer and
// sets tau to the k-th best score.  Because later corpus slabs only hold larger row
// numbers, "score > tau" (strict) is then an exact filter under the (score desc, row
// asc) order.  final_pass additionally sorts the survivors and writes D / I.
// dynamic smem: cap keys, then next_pow2(k) keys for the final sort
__global__ void __launch_bounds__(kSelThreads)
compact_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt, float* __restrict__ tau,
               uint32_t* __restrict__ overflow, const float* __restrict__ margin, float* __restrict__ spec, int cap,
               int k, int final_pass, int spec_rank, int verify, float* __restrict__ D, int64_t* __restrict__ I,
               int64_t id_base) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  const int64_t q = blockIdx.x;
  const uint32_t n_raw = cnt[q];
  if (n_raw > (uint32_t)cap && threadIdx.x == 0) atomicOr(overflow, CMX_OVF_BUFFER);
  const int n = (int)min(n_raw, (uint32_t)cap);
  // Speculative threshold (see plan_slabs): the slab just scored was filtered with spec[q] > the
  // provably safe threshold.  That was valid iff the k-th best now clears it by the margin --
  // then everything within the margin band of the k-th best was above the filter.
  const float spec_old = (verify && spec) ? spec[q] : CMX_NEG_PAD;
  if (spec_old > CMX_NEG_PAD && n <= k && threadIdx.x == 0) {
    // a speculation is only published over >= k kept candidates, so n <= k means nothing new cleared
    // it: count that as a failed guess (conservative; the chunk is redone with the planned slabs)
    atomicOr(overflow, CMX_OVF_SPEC);
    spec[q] = CMX_NEG_PAD;
  }
  if (!final_pass && n <= k) return;  // nothing to drop yet; tau stays
  uint64_t* buf = cand + q * (int64_t)cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = buf[i];
  __syncthreads();
  uint64_t* top = keys + cap;  // survivors (final pass only)
  int kk = n;
  if (n > k) {
    const uint64_t kth = block_kth_largest(keys, n, k, sh);
    // rescore mode: scores are approximate (|approx - exact| <= margin/2), so everything within
    // `margin` below the k-th best approximate score is kept -- that provably contains the exact
    // top-k -- and the filter threshold is lowered by the same amount
    const float m = margin ? margin[q] : 0.f;
    uint64_t thr = kth;
    float new_tau = key_score(kth);
    if (m > 0.f && kth != 0ull) {
      new_tau = key_score(kth) - m;
      thr = make_key(new_tau, 0xffffffffu);  // smallest key carrying that score
    }
    if (spec_old > CMX_NEG_PAD && threadIdx.x == 0) {
      const bool cleared = kth != 0ull && (m > 0.f ? key_score(kth) >= spec_old + m : key_score(kth) > spec_old);
      if (!cleared) atomicOr(overflow, CMX_OVF_SPEC);
      spec[q] = CMX_NEG_PAD;
    }
    float spec_tau = CMX_NEG_PAD;
    if (spec_rank > 0 && spec_rank < k && kth != 0ull) {
      // the spec_rank-th best so far estimates (with a 3x safety factor, plan_slabs) where the k-th
      // best of the WHOLE corpus will be; `keys` still holds all n candidates
      __syncthreads();
      const uint64_t rth = block_kth_largest(keys, n, spec_rank, sh);
      if (rth != 0ull) spec_tau = key_score(rth) - m;
    }
    kk = block_partition(keys, n, thr, final_pass ? top : buf, sh);
    if (threadIdx.x == 0) {
      cnt[q] = (uint32_t)kk;
      if (kth != 0ull) tau[q] = new_tau;
      if (spec_tau > new_tau && kth != 0ull) {  // only a threshold above the safe one is a speculation
        tau[q] = spec_tau;
        spec[q] = spec_tau;
      }
      // rescore mode: the k best plus their margin band must fit half of the buffer (the other half
      // is the room the slab plan counts on, and rescore_kernel holds at most cap/2 keys)
      if (m > 0.f && kk > cap / 2) atomicOr(overflow, CMX_OVF_BAND);
    }
  } else if (final_pass) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) top[i] = keys[i];
    __syncthreads();
  }
  if (final_pass) {
    const int P = next_pow2(max(kk, 2));
    for (int i = kk + threadIdx.x; i < P; i += blockDim.x) top[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(top, P);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      float s = CMX_NEG_PAD;
      int64_t id = -1;
      if (i < kk && top[i] != 0ull) {
        s = key_score(top[i]);
        id = id_base + (int64_t)key_row(top[i]);
      }
      D[q * k + i] = s;
      I[q * k + i] = id;
    }
  }
}

// Rescore mode, last step.  One CTA per query: every surviving candidate (a superset of the
// exact top-k, selected on approximate fp16 tensor-core scores) gets its EXACT fp32 score
// from the fp32 row store, then the exact top-k under (score desc, row asc) is selected,
// sorted and written to D / I.
//  1. the candidates that pass the cut are gathered into shared memory (sharded search: the
//     cut is the GLOBAL k-th best approximate score minus the margin, so a shard keeps ~1/G
//     of its local band);
//  2. one warp per TWO candidate rows at a time (16 independent 128-bit loads in flight per
//     lane), fp32 FMA chain + shuffle reduction -- deterministic for a given (query, row);
//  3. radix select + bitonic sort of the k best exact keys.
// dynamic smem: cap/2 keys | next_pow2(k) keys | d floats (the query): 44 KB at k = 1000,
// d = 1024, so four 512-thread CTAs share an SM and the row gathers of one hide the select
// phase of another.
__device__ __forceinline__ float dot_row_pair_reduce(float acc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__global__ void __launch_bounds__(kSelThreads)
rescore_kernel(const float* __restrict__ X, int d, const float* __restrict__ Q, const uint64_t* __restrict__ cand,
               const uint32_t* __restrict__ cnt, uint32_t* __restrict__ overflow, const float* __restrict__ margin,
               const RescoreCut cut, int cap, int k, int topn, float* __restrict__ D, int64_t* __restrict__ I,
               int64_t id_base) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  __shared__ uint32_t s_m;
  const int half = cap >> 1;
  uint64_t* top = keys + half;
  float* qv = reinterpret_cast<float*>(top + topn);
  const int64_t q = blockIdx.x;
  const uint32_t n_raw = cnt[q];
  if (threadIdx.x == 0) {
    s_m = 0;
    if (n_raw > (uint32_t)cap) atomicOr(overflow, CMX_OVF_BUFFER);
  }
  const int n = (int)min(n_raw, (uint32_t)cap);
  for (int i = threadIdx.x; i < d; i += blockDim.x) qv[i] = Q[q * (int64_t)d + i];
  // sharded search: any lower bound of the GLOBAL k-th best approximate score is a valid cut
  float lowest = CMX_NEG_PAD;
  if (cut.nparts > 0) {
    float ak = cut.kth[0][q];
    for (int g = 1; g < cut.nparts; ++g) ak = fmaxf(ak, cut.kth[g][q]);
    lowest = ak - (margin ? margin[q] : 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint64_t* buf = cand + q * (int64_t)cap;
  // 1. gather (order is irrelevant: exact keys are distinct and the selection is a total order)
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const uint64_t key = i < n ? buf[i] : 0ull;
    const bool valid = key != 0ull && key_score(key) >= lowest;
    const uint32_t ballot = __ballot_sync(0xffffffffu, valid);
    uint32_t wbase = 0;
    if (lane == 0 && ballot) wbase = atomicAdd(&s_m, (uint32_t)__popc(ballot));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    const uint32_t pos = wbase + __popc(ballot & ((1u << lane) - 1u));
    if (valid && pos < (uint32_t)half) keys[pos] = key;
  }
  __syncthreads();
  if (s_m > (uint32_t)half && threadIdx.x == 0) atomicOr(overflow, CMX_OVF_BAND);  // compact_kernel flags this first
  const int m = (int)min(s_m, (uint32_t)half);
  // 2. exact scores, two rows per warp and iteration
  const bool vec = (d & 3) == 0;
  for (int i = 2 * warp; i < m; i += 2 * nwarps) {
    const bool two = i + 1 < m;
    const uint32_t row0 = key_row(keys[i]);
    const uint32_t row1 = two ? key_row(keys[i + 1]) : row0;
    const float* x0 = X + (int64_t)row0 * d;
    const float* x1 = X + (int64_t)row1 * d;
    float acc0 = 0.f, acc1 = 0.f;
    if (vec) {
      const float4* a4 = reinterpret_cast<const float4*>(x0);
      const float4* b4 = reinterpret_cast<const float4*>(x1);
      const float4* q4 = reinterpret_cast<const float4*>(qv);
#pragma unroll 4
      for (int j = lane; j < (d >> 2); j += 32) {
        const float4 a = a4[j], b = b4[j], c = q4[j];
        acc0 = fmaf(a.x, c.x, acc0); acc0 = fmaf(a.y, c.y, acc0); acc0 = fmaf(a.z, c.z, acc0); acc0 = fmaf(a.w, c.w, acc0);
        acc1 = fmaf(b.x, c.x, acc1); acc1 = fmaf(b.y, c.y, acc1); acc1 = fmaf(b.z, c.z, acc1); acc1 = fmaf(b.w, c.w, acc1);
      }
    } else {
      for (int j = lane; j < d; j += 32) {
        const float c = qv[j];
        acc0 = fmaf(x0[j], c, acc0);
        acc1 = fmaf(x1[j], c, acc1);
      }
    }
    acc0 = dot_row_pair_reduce(acc0);
    acc1 = dot_row_pair_reduce(acc1);
    __syncwarp();
    if (lane == 0) {
      keys[i] = acc0 > CMX_NEG_PAD ? make_key(acc0, row0) : 0ull;
      if (two) keys[i + 1] = acc1 > CMX_NEG_PAD ? make_key(acc1, row1) : 0ull;
    }
  }
  __syncthreads();
  // 3. exact top-k
  int kk = m;
  if (m > k) {
    const uint64_t kth = block_kth_largest(keys, m, k, sh);
    kk = block_partition(keys, m, kth, top, sh);
  } else {
    for (int i = threadIdx.x; i < m; i += blockDim.x) top[i] = keys[i];
    __syncthreads();
  }
  const int P = next_pow2(max(kk, 2));
  for (int i = kk + threadIdx.x; i < P; i += blockDim.x) top[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(top, P);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float s = CMX_NEG_PAD;
    int64_t id = -1;
    if (i < kk && top[i] != 0ull) {
      s = key_score(top[i]);
      id = id_base + (int64_t)key_row(top[i]);
    }
    D[q * k + i] = s;
    I[q * k + i] = id;
  }
}

static int pow2_at_least(int n) {
  int p = 2;
  while (p < n) p <<= 1;
  return p;
}

int launch_compact(const SearchWs& ws, int64_t nq, int k, int final_pass, float* D, int64_t* I,
                   int64_t id_base, cudaStream_t st, int spec_rank, int verify) {
  if (nq == 0) return CMX_OK;
  const size_t smem = ((size_t)ws.cap + pow2_at_least(k)) * sizeof(uint64_t);
  CMX_CUDA(cudaFuncSetAttribute(compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  compact_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(ws.cand, ws.cnt, ws.tau, ws.overflow, ws.margin,
                                                          ws.spec ? ws.spec : nullptr, ws.cap, k, final_pass,
                                                          ws.spec ? spec_rank : 0, ws.spec ? verify : 0, D, I, id_base);
  CMX_LAUNCHED();
  return CMX_OK;
}

// a_k[q] = k-th best approximate score of this shard so far = tau + margin
__global__ void kth_approx_kernel(const float* __restrict__ tau, const float* __restrict__ margin, int64_t nq,
                                  float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) out[i] = tau[i] + (margin ? margin[i] : 0.f);
}

// Export the shard's k best APPROXIMATE scores per query (order arbitrary, short lists padded with
// lowest-float) so that the shards can find the GLOBAL k-th best approximate score together.
__global__ void __launch_bounds__(256)
export_scores_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt, const float* __restrict__ tau,
                     const float* __restrict__ margin, int cap, int k, float* __restrict__ out) {
  __shared__ uint32_t s_n;
  const int64_t q = blockIdx.x;
  const int n = (int)min(cnt[q], (uint32_t)cap);
  const float ak = tau[q] + (margin ? margin[q] : 0.f);  // local k-th best approximate score (or lower)
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const uint64_t* buf = cand + q * (int64_t)cap;
  float* o = out + q * (int64_t)k;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t key = buf[i];
    if (key != 0ull) {
      const float s = key_score(key);
      if (s >= ak) {
        const uint32_t pos = atomicAdd(&s_n, 1u);
        if (pos < (uint32_t)k) o[pos] = s;
      }
    }
  }
  __syncthreads();
  for (int i = (int)min(s_n, (uint32_t)k) + threadIdx.x; i < k; i += blockDim.x) o[i] = CMX_NEG_PAD;
}

int launch_export_scores(const SearchWs& ws, int64_t nq, int k, float* out, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  export_scores_kernel<<<(unsigned)nq, 256, 0, st>>>(ws.cand, ws.cnt, ws.tau, ws.margin, ws.cap, k, out);
  CMX_LAUNCHED();
  return CMX_OK;
}

// GLOBAL k-th best approximate score of queries [q0, q1): the k-th largest of the union of every
// shard's exported list, read in place (peer memory), written into every shard's kth array.
struct UnionArgs {
  const float* parts[CMX_MAX_PEERS];
  float* outs[CMX_MAX_PEERS];
};

__global__ void __launch_bounds__(kSelThreads)
union_kth_kernel(const UnionArgs a, int nparts, int nouts, int64_t q0, int k) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  const int64_t q = q0 + blockIdx.x;
  const int n_all = nparts * k;
  for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
    const int g = i / k, pos = i - g * k;
    keys[i] = make_key(a.parts[g][q * k + pos], (uint32_t)i);  // distinct keys; padding sorts last
  }
  __syncthreads();
  float v;
  if (n_all > k) v = key_score(block_kth_largest(keys, n_all, k, sh));
  else {
    // single part: the k-th best is the minimum of the list
    float m = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) m = fminf(m, key_score(keys[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fminf(m, red[w]);
    v = m;
  }
  if (threadIdx.x == 0)
    for (int o = 0; o < nouts; ++o) a.outs[o][q] = v;
}

int launch_union_kth(const float* const* parts, int nparts, int k, int64_t q0, int64_t q1, float* const* outs, int nouts,
                     cudaStream_t st) {
  if (q1 <= q0) return CMX_OK;
  CMX_CHECK(nparts >= 1 && nparts <= CMX_MAX_PEERS && nouts >= 1 && nouts <= CMX_MAX_PEERS, "union_kth: at most %d parts",
            CMX_MAX_PEERS);
  const size_t smem = (size_t)nparts * k * sizeof(uint64_t);
  CMX_CHECK(smem <= 200 * 1024, "union_kth: nparts*k = %d too large", nparts * k);
  UnionArgs a;
  for (int g = 0; g < nparts; ++g) a.parts[g] = parts[g];
  for (int o = 0; o < nouts; ++o) a.outs[o] = outs[o];
  CMX_CUDA(cudaFuncSetAttribute(union_kth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  union_kth_kernel<<<(unsigned)(q1 - q0), kSelThreads, smem, st>>>(a, nparts, nouts, q0, k);
  CMX_LAUNCHED();
  return CMX_OK;
}

int launch_kth_approx(const SearchWs& ws, int64_t nq, float* out, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  kth_approx_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(ws.tau, ws.margin, nq, out);
  CMX_LAUNCHED();
  return CMX_OK;
}

int launch_rescore(const float* X, int d, const float* Q, const SearchWs& ws, int64_t nq, int k, float* D,
                   int64_t* I, int64_t id_base, const RescoreCut& cut, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  const int topn = pow2_at_least(k);
  const size_t smem = ((size_t)(ws.cap >> 1) + topn) * sizeof(uint64_t) + (size_t)d * sizeof(float);
  CMX_CHECK(smem <= 220 * 1024, "rescore: d=%d too large for the shared-memory query copy", d);
  CMX_CUDA(cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rescore_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(X, d, Q, ws.cand, ws.cnt, ws.overflow, ws.margin, cut, ws.cap, k,
                                                          topn, D, I, id_base);
  CMX_LAUNCHED();
  return CMX_OK;
}

// k-way merge: parts are individually sorted (score desc, row asc) and ordered by
// ascending row range, so (part, position) is the tie-break that reproduces the
// single-shard order exactly.  dynamic smem: nparts*k keys, then next_pow2(k) keys
__global__ void __launch_bounds__(kSelThreads)
merge_kernel(const float* __restrict__ Dp, const int64_t* __restrict__ Ip, int nparts, int64_t nq,
             int k, float* __restrict__ D, int64_t* __restrict__ I) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  __shared__ uint32_t s_valid;
  const int64_t q = blockIdx.x;
  const int n_all = nparts * k;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  // compact the valid (id >= 0) entries of all parts into keys[0..n)
  {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n_all; base += blockDim.x) {
      const int i = base + threadIdx.x;
      bool valid = false;
      uint64_t key = 0ull;
      if (i < n_all) {
        const int g = i / k, pos = i - g * k;
        const int64_t src = ((int64_t)g * nq + q) * k + pos;
        valid = Ip[src] >= 0;
        if (valid) key = make_key(Dp[src], (uint32_t)i);
      }
      const uint32_t ballot = __ballot_sync(0xffffffffu, valid);
      uint32_t wbase = 0;
      if (lane == 0 && ballot) wbase = atomicAdd(&s_valid, (uint32_t)__popc(ballot));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (valid) keys[wbase + __popc(ballot & ((1u << lane) - 1u))] = key;
    }
  }
  __syncthreads();
  const int n = (int)s_valid;
  uint64_t* top = keys + n_all;
  int kk = n;
  if (n > k) {
    const uint64_t kth = block_kth_largest(keys, n, k, sh);
    kk = block_partition(keys, n, kth, top, sh);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) top[i] = keys[i];
    __syncthreads();
  }
  const int P = next_pow2(max(kk, 2));
  for (int i = kk + threadIdx.x; i < P; i += blockDim.x) top[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(top, P);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float s = CMX_NEG_PAD;
    int64_t id = -1;
    if (i < kk) {
      const uint32_t src_i = key_row(top[i]);
      const int g = src_i / k, pos = src_i - g * k;
      const int64_t src = ((int64_t)g * nq + q) * k + pos;
      s = Dp[src];
      id = Ip[src];
    }
    D[q * k + i] = s;
    I[q * k + i] = id;
  }
}

// Fused exchange + merge over peer memory: part g's list lives in GPU g's memory and is read
// here directly through NVLink peer pointers (no all_gather, no staging copy); the merged
// rows of queries [q0, q1) are stored into every rank's output buffer (peer stores), so
// after one cross-rank barrier each rank holds the full result.
struct MergePeerArgs {
  const float* D[CMX_MAX_PEERS];
  const int64_t* I[CMX_MAX_PEERS];
  float* Do[CMX_MAX_PEERS];
  int64_t* Io[CMX_MAX_PEERS];
};

__global__ void __launch_bounds__(kSelThreads)
merge_peers_kernel(const MergePeerArgs a, int nparts, int nouts, int64_t q0, int k) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  __shared__ uint32_t s_valid;
  const int64_t q = q0 + blockIdx.x;
  const int n_all = nparts * k;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n_all; base += blockDim.x) {
      const int i = base + threadIdx.x;
      bool valid = false;
      uint64_t key = 0ull;
      if (i < n_all) {
        const int g = i / k, pos = i - g * k;
        const int64_t src = q * k + pos;
        valid = a.I[g][src] >= 0;
        if (valid) key = make_key(a.D[g][src], (uint32_t)i);
      }
      const uint32_t ballot = __ballot_sync(0xffffffffu, valid);
      uint32_t wbase = 0;
      if (lane == 0 && ballot) wbase = atomicAdd(&s_valid, (uint32_t)__popc(ballot));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (valid) keys[wbase + __popc(ballot & ((1u << lane) - 1u))] = key;
    }
  }
  __syncthreads();
  const int n = (int)s_valid;
  uint64_t* top = keys + n_all;
  int kk = n;
  if (n > k) {
    const uint64_t kth = block_kth_largest(keys, n, k, sh);
    kk = block_partition(keys, n, kth, top, sh);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) top[i] = keys[i];
    __syncthreads();
  }
  const int P = next_pow2(max(kk, 2));
  for (int i = kk + threadIdx.x; i < P; i += blockDim.x) top[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(top, P);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float s = CMX_NEG_PAD;
    int64_t id = -1;
    if (i < kk) {
      const uint32_t src_i = key_row(top[i]);
      const int g = src_i / k, pos = src_i - g * k;
      s = a.D[g][q * k + pos];
      id = a.I[g][q * k + pos];
    }
    for (int o = 0; o < nouts; ++o) {
      a.Do[o][q * k + i] = s;
      a.Io[o][q * k + i] = id;
    }
  }
}

int launch_merge_peers(const float* const* D_parts, const int64_t* const* I_parts, int nparts, int k,
                       int64_t q0, int64_t q1, float* const* D_outs, int64_t* const* I_outs, int nouts,
                       cudaStream_t st) {
  if (q1 <= q0) return CMX_OK;
  CMX_CHECK(nparts >= 1 && nparts <= CMX_MAX_PEERS && nouts >= 1 && nouts <= CMX_MAX_PEERS,
            "merge_peers: at most %d parts / outputs", CMX_MAX_PEERS);
  const size_t smem = ((size_t)nparts * k + pow2_at_least(k)) * sizeof(uint64_t);
  if (smem > 200 * 1024) {
    set_error("merge: nparts*k = %d too large for one shared-memory selection", nparts * k);
    return CMX_ERR_INVALID;
  }
  MergePeerArgs a;
  for (int g = 0; g < nparts; ++g) { a.D[g] = D_parts[g]; a.I[g] = I_parts[g]; }
  for (int o = 0; o < nouts; ++o) { a.Do[o] = D_outs[o]; a.Io[o] = I_outs[o]; }
  CMX_CUDA(cudaFuncSetAttribute(merge_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_peers_kernel<<<(unsigned)(q1 - q0), kSelThreads, smem, st>>>(a, nparts, nouts, q0, k);
  CMX_LAUNCHED();
  return CMX_OK;
}

int launch_merge(const float* D_parts, const int64_t* I_parts, int nparts, int64_t nq, int k,
                 float* D, int64_t* I, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  const size_t smem = ((size_t)nparts * k + pow2_at_least(k)) * sizeof(uint64_t);
  if (smem > 200 * 1024) {
    set_error("merge: nparts*k = %d too large for one shared-memory selection", nparts * k);
    return CMX_ERR_INVALID;
  }
  CMX_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(D_parts, I_parts, nparts, nq, k, D, I);
  CMX_LAUNCHED();
  return CMX_OK;
}

}  // namespace cmx

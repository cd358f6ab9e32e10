// k-selection: candidate-buffer compaction (bitonic sort in shared memory),
// final (D, I) emission and the k-way merge of per-shard results.
//
// Replaces FAISS's blockSelect / heap pass behind index.search
// (reference call sites onepass_dense_mix_run_custom_lang.py:878,
// onepass_bilingual_mix_hub_custom_lang.py:950).  The scoring kernels never
// materialise the score matrix: they append (score,row) keys that beat the
// query's running threshold tau to a per-query buffer; this file turns the
// buffers into exact, deterministically ordered top-k lists.
#include "common.cuh"

namespace cmx {

__global__ void ws_init_kernel(float* tau, uint32_t* cnt, uint32_t* overflow, int64_t nq, int64_t nq_pad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) overflow[0] = 0;
  if (i < nq_pad) {
    // FAISS's IP heap starts at lowest-float and admits only strictly larger scores;
    // padded (non-existent) queries get +inf so nothing ever passes the filter.
    tau[i] = (i < nq) ? CMX_NEG_PAD : __int_as_float(0x7f800000);
    cnt[i] = 0;
  }
}

int launch_ws_init(const SearchWs& ws, int64_t nq, int64_t nq_pad, cudaStream_t st) {
  int64_t blocks = (nq_pad + 255) / 256;
  if (blocks < 1) blocks = 1;
  ws_init_kernel<<<(unsigned)blocks, 256, 0, st>>>(ws.tau, ws.cnt, ws.overflow, nq, nq_pad);
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

// One CTA per query.  Sorts the query's candidates, keeps the best k at the front
// of its buffer and sets tau to the k-th best score.  Because later corpus slabs
// only hold larger row numbers, "score > tau" (strict) is then an exact filter
// under the (score desc, row asc) order.
__global__ void __launch_bounds__(512)
compact_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt, float* __restrict__ tau,
               uint32_t* __restrict__ overflow, int cap, int k, int final_pass,
               float* __restrict__ D, int64_t* __restrict__ I, int64_t id_base) {
  extern __shared__ __align__(16) uint64_t keys[];
  const int64_t q = blockIdx.x;
  const uint32_t n_raw = cnt[q];
  if (n_raw > (uint32_t)cap && threadIdx.x == 0) atomicExch(overflow, 1u);
  const int n = (int)min(n_raw, (uint32_t)cap);
  if (!final_pass && n <= k) return;  // nothing to drop yet; tau stays
  uint64_t* buf = cand + q * (int64_t)cap;
  const int P = next_pow2(max(n, 2));
  for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = (i < n) ? buf[i] : 0ull;
  __syncthreads();
  bitonic_sort_desc(keys, P);
  const int kk = min(n, k);
  for (int i = threadIdx.x; i < kk; i += blockDim.x) buf[i] = keys[i];
  if (threadIdx.x == 0) {
    cnt[q] = (uint32_t)kk;
    if (n >= k && keys[k - 1] != 0ull) tau[q] = key_score(keys[k - 1]);
  }
  if (final_pass) {
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      float s = CMX_NEG_PAD;
      int64_t id = -1;
      if (i < kk && keys[i] != 0ull) {
        s = key_score(keys[i]);
        id = id_base + (int64_t)key_row(keys[i]);
      }
      D[q * k + i] = s;
      I[q * k + i] = id;
    }
  }
}

int launch_compact(const SearchWs& ws, int64_t nq, int k, int final_pass, float* D, int64_t* I,
                   int64_t id_base, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  const size_t smem = (size_t)ws.cap * sizeof(uint64_t);
  CMX_CUDA(cudaFuncSetAttribute(compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  compact_kernel<<<(unsigned)nq, 512, smem, st>>>(ws.cand, ws.cnt, ws.tau, ws.overflow, ws.cap, k,
                                                  final_pass, D, I, id_base);
  CMX_LAUNCHED();
  return CMX_OK;
}

// k-way merge: parts are individually sorted (score desc, row asc) and ordered by
// ascending row range, so (part, position) is the tie-break that reproduces the
// single-shard order exactly.
__global__ void __launch_bounds__(512)
merge_kernel(const float* __restrict__ Dp, const int64_t* __restrict__ Ip, int nparts, int64_t nq,
             int k, float* __restrict__ D, int64_t* __restrict__ I) {
  extern __shared__ __align__(16) uint64_t keys[];
  const int64_t q = blockIdx.x;
  const int n = nparts * k;
  const int P = next_pow2(max(n, 2));
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t key = 0ull;
    if (i < n) {
      const int g = i / k, pos = i - g * k;
      const int64_t src = ((int64_t)g * nq + q) * k + pos;
      if (Ip[src] >= 0) key = make_key(Dp[src], (uint32_t)i);
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc(keys, P);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float s = CMX_NEG_PAD;
    int64_t id = -1;
    const uint64_t key = keys[i];
    if (i < n && key != 0ull) {
      const uint32_t src_i = key_row(key);
      const int g = src_i / k, pos = src_i - g * k;
      const int64_t src = ((int64_t)g * nq + q) * k + pos;
      s = Dp[src];
      id = Ip[src];
    }
    D[q * k + i] = s;
    I[q * k + i] = id;
  }
}

int launch_merge(const float* D_parts, const int64_t* I_parts, int nparts, int64_t nq, int k,
                 float* D, int64_t* I, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  int n = nparts * k;
  int P = 2;
  while (P < n) P <<= 1;
  const size_t smem = (size_t)P * sizeof(uint64_t);
  if (smem > 200 * 1024) {
    set_error("merge: nparts*k = %d too large for one shared-memory sort", n);
    return CMX_ERR_INVALID;
  }
  CMX_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_kernel<<<(unsigned)nq, 512, smem, st>>>(D_parts, I_parts, nparts, nq, k, D, I);
  CMX_LAUNCHED();
  return CMX_OK;
}

}  // namespace cmx

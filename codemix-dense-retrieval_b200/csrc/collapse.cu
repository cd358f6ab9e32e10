// Collapse-by-base-id with fuse = max, on the device.
//
// Replaces the grouping half of collapse_run_max (reference
// onepass_bilingual_mix_hub_custom_lang.py:165-181), which re-parses the raw run text:
// per query, group the hits by base = did.split('#')[0], keep the max of the scores AS THE
// RAW FILE PRINTS THEM (6 decimals), sort the groups by that value descending -- Python's
// stable sort, so equal values keep the order in which their bases first appeared -- and
// re-rank.  Here the derived id -> base id map is an int32 code per corpus row (base_code),
// so a query's (D, I) row becomes (base code, value) pairs without any text:
//   value = sign * rint(|score| * 1e6)   (float32 * 1e6 is exact in double; ties to even =
//           Python's correctly rounded f"{score:.6f}")
//   1. sort (code, position) ascending          -> groups contiguous, first-seen hit first
//   2. group heads take the max value of their group and the head's position
//   3. sort (value desc, first-seen position asc)
// One CTA per query, bitonic sorts of 64-bit keys in shared memory (k <= 2048).  The host
// formatter (trec_text.cpp) then writes the collapsed run from these lists and the raw run
// from (D, I): byte-identical to the reference's text round trip.  Anything the 64-bit keys
// cannot carry exactly -- non-finite scores, |score| >= 2^20, a negative zero -- sets a
// status word and the host grouping (which handles them) is used for that batch.
#include "common.cuh"

namespace cmx {

constexpr int kColThreads = 256;
constexpr long long kValBias = 1ll << 41;  // |value| < 2^40 * ... : value + bias fits 42 bits

__device__ __forceinline__ void bitonic_sort_u64(uint64_t* keys, int P, bool descending) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0) != descending;  // ascending run when !descending
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
}

// dynamic smem: P u64 keys | k i64 values
__global__ void __launch_bounds__(kColThreads)
collapse_max_kernel(const float* __restrict__ D, const int64_t* __restrict__ I, int k, int P,
                    const int32_t* __restrict__ base_code, int64_t ndocs, int32_t* __restrict__ out_code,
                    int64_t* __restrict__ out_val6, int32_t* __restrict__ out_count, uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint64_t ckeys[];
  long long* vals = reinterpret_cast<long long*>(ckeys + P);
  __shared__ int s_count;
  const int64_t q = blockIdx.x;
  const float* Dq = D + q * k;
  const int64_t* Iq = I + q * k;
  if (threadIdx.x == 0) s_count = 0;
  // 1. (code, position) keys; values as the raw file prints them
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    uint64_t key = ~0ull;  // invalid hits sort last
    if (j < k) {
      const int64_t ix = Iq[j];
      long long v = 0;
      if (ix >= 0 && ix < ndocs) {
        const float s = Dq[j];
        const double ax = fabs((double)s);
        if (!isfinite(s) || ax >= 1048576.0 || (s == 0.0f && signbit(s))) {
          atomicOr(status, 1u);  // the host grouping handles these exactly
        } else {
          v = (long long)rint(ax * 1e6);
          if (s < 0.0f) {
            if (v == 0) atomicOr(status, 1u);  // prints as "-0.000000": keep the sign on the host path
            v = -v;
          }
          key = ((uint64_t)(uint32_t)base_code[ix] << 32) | (uint32_t)j;
        }
      }
      vals[j] = v;
    }
    ckeys[j] = key;
  }
  __syncthreads();
  bitonic_sort_u64(ckeys, P, false);
  // 2. group heads: max value of the group, position of the first-seen hit
  uint64_t mine[8];  // this thread's P / blockDim.x <= 8 new keys (P <= 2048)
  int nm = 0;
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    const uint64_t key = ckeys[t];
    uint64_t out = 0ull;
    if (key != ~0ull) {
      const uint32_t code = (uint32_t)(key >> 32);
      const bool head = t == 0 || (uint32_t)(ckeys[t - 1] >> 32) != code;
      if (head) {
        const int pos0 = (int)(uint32_t)key;
        long long best = vals[pos0];
        for (int u = t + 1; u < P; ++u) {
          const uint64_t ku = ckeys[u];
          if (ku == ~0ull || (uint32_t)(ku >> 32) != code) break;
          const long long vu = vals[(int)(uint32_t)ku];
          best = vu > best ? vu : best;
        }
        out = ((uint64_t)(best + kValBias) << 12) | (uint64_t)(4095 - pos0);
        atomicAdd(&s_count, 1);
      }
    }
    mine[nm++] = out;
  }
  __syncthreads();  // every thread has read the neighbours it needs
  nm = 0;
  for (int t = threadIdx.x; t < P; t += blockDim.x) ckeys[t] = mine[nm++];
  __syncthreads();
  // 3. value descending, first-seen position ascending
  bitonic_sort_u64(ckeys, P, true);
  const int count = s_count;
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    int32_t code = -1;
    long long v6 = 0;
    if (r < count) {
      const uint64_t key = ckeys[r];
      const int pos0 = 4095 - (int)(key & 4095ull);
      v6 = (long long)(key >> 12) - kValBias;
      code = base_code[Iq[pos0]];
    }
    out_code[q * k + r] = code;
    out_val6[q * k + r] = v6;
  }
  if (threadIdx.x == 0) out_count[q] = count;
}

int launch_collapse_max(const float* D, const int64_t* I, int64_t nq, int k, const int32_t* base_code, int64_t ndocs,
                        int32_t* out_code, int64_t* out_val6, int32_t* out_count, uint32_t* status, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  int P = 2;
  while (P < k) P <<= 1;
  CMX_CHECK(P <= 2048, "collapse: k=%d exceeds 2048", k);
  const size_t smem = (size_t)P * sizeof(uint64_t) + (size_t)k * sizeof(long long);
  collapse_max_kernel<<<(unsigned)nq, kColThreads, smem, st>>>(D, I, k, P, base_code, ndocs, out_code, out_val6, out_count, status);
  CMX_LAUNCHED();
  return CMX_OK;
}

}  // namespace cmx

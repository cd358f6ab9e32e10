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
// rescore_kernel: 52 registers = 2 CTAs per SM.  Forcing 3 (40 registers, fewer loads in flight per warp) measured
// 0.5 ms SLOWER per C2 step in a same-box A/B (5.9 vs 6.4 ms for the final compaction + rescoring).
#ifndef RESCORE_MIN_CTAS
#define RESCORE_MIN_CTAS 2
#endif

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

constexpr int kSelBins = 512;      // value bins of the first selection step
constexpr int kSelListMax = 256;   // keys of the crossing bin that are ranked directly (they reuse hist)
struct SelectShared {
  uint32_t hist[kSelBins];
  uint64_t red_min[16];  // one per warp (kSelThreads = 512)
  uint64_t red_max[16];
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
  // Fast path.  Candidate scores are bell-shaped, so the leading radix digits (sign, exponent) put
  // nearly all keys into one or two bins -- thousands of shared-memory atomics on one address, and
  // four or five passes before the k-th is isolated.  Binning the SCORE linearly between the smallest
  // and the largest spreads the keys over all kSelBins bins; float subtract, multiply and truncation
  // are monotone, so bin order agrees with key order and the k-th largest key lies in the bin where
  // the count from the top crosses k.  That bin's keys (a few dozen) are then ranked against each
  // other on the full 64 bits.  Degenerate inputs (non-finite or padded scores, heavy ties: more than
  // kSelListMax keys in the crossing bin) take the radix passes below.
  {
    const float smin = key_score(kmin), smax = key_score(kmax);
    const float scale = (float)kSelBins / (smax - smin);
    if (isfinite(smin) && isfinite(smax) && isfinite(scale) && scale > 0.f) {
      for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) sh.hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int b = min(kSelBins - 1, (int)((key_score(keys[i]) - smin) * scale));
        atomicAdd(&sh.hist[b], 1u);
      }
      __syncthreads();
      if (warp == 0) {
        constexpr int kPer = kSelBins / 32;  // lane l owns bins [kPer l, kPer (l + 1))
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) s += sh.hist[lane * kPer + j];
        uint32_t incl = s;  // inclusive suffix sum over lanes (lane 31 = largest scores)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_down_sync(0xffffffffu, incl, o);
          if (lane + o < 32) incl += t;
        }
        uint32_t running = incl - s;
        if (running < (uint32_t)k && (uint32_t)k <= incl) {  // the crossing bin is one of this lane's
          for (int j = kPer - 1; j >= 0; --j) {
            const uint32_t h = sh.hist[lane * kPer + j];
            if ((uint32_t)k <= running + h) { sh.bin = lane * kPer + j; sh.above = running; sh.bin_count = h; break; }
            running += h;
          }
        }
        if (lane == 0) sh.count = 0;
      }
      __syncthreads();
      const int b_k = (int)sh.bin, m = (int)sh.bin_count, want_in = k - (int)sh.above;
      __syncthreads();  // hist is reused as the key list below
      if (m <= kSelListMax) {
        uint64_t* list = reinterpret_cast<uint64_t*>(sh.hist);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const uint64_t v = keys[i];
          if (min(kSelBins - 1, (int)((key_score(v) - smin) * scale)) == b_k) list[atomicAdd(&sh.count, 1u)] = v;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += blockDim.x) {
          const uint64_t v = list[t];
          int larger = 0;
          for (int j = 0; j < m; ++j) larger += list[j] > v;
          if (larger == want_in - 1) sh.found = v;  // keys are distinct: exactly one thread
        }
        __syncthreads();
        const uint64_t found = sh.found;
        __syncthreads();  // callers reuse sh (and call again) right away
        return found;
      }
    }
  }
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
// INPLACE additionally packs the survivors to the front of `keys` itself.  Survivor number p lands
// at keys[p] with p < (elements visited so far), so the only hazard is a slot of the chunk being
// visited that another warp has not read yet: one barrier per chunk between the reads and the
// writes removes it.
template <bool INPLACE = false>
__device__ int block_partition(uint64_t* keys, int n, uint64_t kth, uint64_t* dst, SelectShared& sh) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) sh.count = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const uint64_t v = (i < n) ? keys[i] : 0ull;
    if (INPLACE) __syncthreads();
    const bool keep = (i < n) && ((kth != 0ull) ? (v >= kth) : (v != 0ull));
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    uint32_t wbase = 0;
    if (lane == 0 && ballot) wbase = atomicAdd(&sh.count, (uint32_t)__popc(ballot));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (keep) {
      const uint32_t pos = wbase + __popc(ballot & ((1u << lane) - 1u));
      dst[pos] = v;
      if (INPLACE) keys[pos] = v;
    }
  }
  __syncthreads();
  return (int)sh.count;
}

// ---- list load: one bulk copy by the TMA engine ------------------------------------------------
// A per-thread "keys[i] = buf[i]" loop keeps ONE 8-byte load in flight per thread (12 KB per SM with three
// CTAs of 512 threads): ncu showed compact_kernel latency-bound on it (a quarter of all stall samples on the
// first shared-memory store, DRAM at 1 TB/s).  One cp.async.bulk per CTA puts the whole list in flight.
__device__ __forceinline__ uint32_t sel_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load_begin(uint64_t* smem_dst, const uint64_t* gsrc, uint32_t nbytes, uint64_t* bar) {
  const uint32_t b = sel_smem_u32(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nbytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sel_smem_u32(smem_dst)), "l"(gsrc), "r"(nbytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_load_wait(uint64_t* bar) {
  const uint32_t b = sel_smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b) : "memory");
  } while (!ok);
}

// One CTA per query.  Keeps the query's best k candidates (plus, in rescore mode, the margin
// band below them) at the front of its buffer and sets the filter threshold tau for the next
// slab.  The tensor kernels visit the corpus in a permuted block order, so rows of later slabs
// are NOT larger than the ones seen: the filter is exact because of the margin (rescore mode:
// everything within 2 eps of the running k-th best approximate score is kept); with
// a zero margin (split precision, stream path in file order) the strict "score > tau" admits no
// row that ties the current k-th score, which at an exact tie across the k-th place yields one
// of the tied rows, not necessarily the smallest id ("Ties").
// final_pass additionally sorts the survivors and writes D / I.
// spec_rank / est_rank > 0 publish two order statistics of the survivors for the NEXT slab:
//   spec[q] (and tau[q]) = score of rank spec_rank - margin: the speculative filter threshold
//   est[q]               = score of rank est_rank: candidates at or above it will very likely be in
//                          the final top-k, prescore_kernel gives them exact scores while the next
//                          slab is still being scored
// dynamic smem: cap keys, then next_pow2(k) keys for the final sort
// 3 CTAs per SM (72 KB of shared memory each): the register budget must allow it too -- at 46 registers only two
// fit and every compaction took 30 % longer (measured A/B against the 40-register round-1 build)
__global__ void __launch_bounds__(kSelThreads, 3)
compact_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt, float* __restrict__ tau,
               uint32_t* __restrict__ overflow, const float* __restrict__ margin, float* __restrict__ spec,
               float* __restrict__ est, int cap, int k, int final_pass, int spec_rank, int est_rank, int verify,
               float* __restrict__ D, int64_t* __restrict__ I, int64_t id_base) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  __shared__ __align__(8) uint64_t load_bar;
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
  if (est && est_rank > 0 && threadIdx.x == 0) est[q] = CMX_NEG_PAD;  // no estimate unless set below
  if (!final_pass && n <= k) return;  // nothing to drop yet; tau stays
  uint64_t* buf = cand + q * (int64_t)cap;
  const uint32_t nbytes = (uint32_t)((n + 1) & ~1) * 8u;  // whole 16-byte units (cap is even, so n + 1 <= cap for odd n)
  if (threadIdx.x == 0 && nbytes) bulk_load_begin(keys, buf, nbytes, &load_bar);
  __syncthreads();  // the barrier is initialised
  if (nbytes) bulk_load_wait(&load_bar);
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
      thr = make_key_floor(new_tau);
    }
    if (spec_old > CMX_NEG_PAD && threadIdx.x == 0) {
      const bool cleared = kth != 0ull && (m > 0.f ? key_score(kth) >= spec_old + m : key_score(kth) > spec_old);
      if (!cleared) atomicOr(overflow, CMX_OVF_SPEC);
      spec[q] = CMX_NEG_PAD;
    }
    // not the final pass: survivors go back to the query's buffer AND to the front of keys (for the order statistics below)
    kk = final_pass ? block_partition<false>(keys, n, thr, top, sh) : block_partition<true>(keys, n, thr, buf, sh);
    float spec_tau = CMX_NEG_PAD, est_tau = CMX_NEG_PAD;
    const bool want_spec = spec_rank > 0 && spec_rank < k;
    const bool want_est = est != nullptr && est_rank > 0 && est_rank < k;
    if (!final_pass && kth != 0ull && (want_spec || want_est)) {
      // order statistics of the survivors only (kk << n after a large slab), packed at keys[0..kk)
      if (want_spec) {
        // the spec_rank-th best so far estimates (with a 3x safety factor, plan_slabs) where the k-th
        // best of the rows up to the end of the next slab will be
        const uint64_t rth = block_kth_largest(keys, kk, spec_rank, sh);
        if (rth != 0ull) spec_tau = key_score(rth) - m;
        __syncthreads();
      }
      if (want_est) {
        const uint64_t eth = block_kth_largest(keys, kk, est_rank, sh);
        if (eth != 0ull) est_tau = key_score(eth);
      }
    }
    if (threadIdx.x == 0) {
      cnt[q] = (uint32_t)kk;
      if (kth != 0ull) tau[q] = new_tau;
      if (spec_tau > new_tau && kth != 0ull) {  // only a threshold above the safe one is a speculation
        tau[q] = spec_tau;
        spec[q] = spec_tau;
      }
      if (want_est) est[q] = est_tau;
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

// ---- exact fp32 score of one (query, row) pair: THE arithmetic of the rescore precision -------
// One warp per row, lane l takes the float4 groups l, l+32, ...: an fp32 FMA chain over the lane's
// elements (x, y, z, w in order), then a xor-shuffle tree.  prescore_kernel and rescore_kernel
// both use these two functions, so a (query, row) score is bit-identical wherever and whenever it
// is computed (1 GPU == G GPUs, prescored == rescored).  qv: the query in shared memory.
__device__ __forceinline__ float warp_reduce_sum(float acc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// STREAM: row loads marked evict-first (ld.global.cs) -- prescore_kernel runs beside the scoring kernel, whose
// corpus tiles and query planes live in L2; 4 KB gathers that are used once must not push them out.  The
// arithmetic is the same either way.
template <bool STREAM>
__device__ __forceinline__ void exact_dot_pair(const float* __restrict__ x0, const float* __restrict__ x1,
                                               const float* qv, int d, int lane, float& out0, float& out1) {
  float acc0 = 0.f, acc1 = 0.f;
  if ((d & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(x0);
    const float4* b4 = reinterpret_cast<const float4*>(x1);
    const float4* q4 = reinterpret_cast<const float4*>(qv);
#pragma unroll 4
    for (int j = lane; j < (d >> 2); j += 32) {
      const float4 a = STREAM ? __ldcs(a4 + j) : a4[j], b = STREAM ? __ldcs(b4 + j) : b4[j], c = q4[j];
      acc0 = fmaf(a.x, c.x, acc0); acc0 = fmaf(a.y, c.y, acc0); acc0 = fmaf(a.z, c.z, acc0); acc0 = fmaf(a.w, c.w, acc0);
      acc1 = fmaf(b.x, c.x, acc1); acc1 = fmaf(b.y, c.y, acc1); acc1 = fmaf(b.z, c.z, acc1); acc1 = fmaf(b.w, c.w, acc1);
    }
  } else {
    // same element -> lane assignment as the vector loop: group j = elements 4j .. 4j+3
    for (int j = lane; 4 * j < d; j += 32) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * j + e;
        if (i < d) {
          const float c = qv[i];
          acc0 = fmaf(x0[i], c, acc0);
          acc1 = fmaf(x1[i], c, acc1);
        }
      }
    }
  }
  out0 = warp_reduce_sum(acc0);
  out1 = warp_reduce_sum(acc1);
}

// ---- exact scores ahead of time --------------------------------------------------------------
// While the LAST slab of a rescore-mode search is being scored (in several launches), this kernel
// runs on a second stream next to the scoring kernel (one small CTA per query: 4 warps, the query
// in shared memory -- it fits beside the persistent scoring CTA) and replaces, in place, the
// approximate score of every candidate in buffer positions [lo[q], hi[q]) whose score reaches
// est[q] by its exact fp32 score (key flag cleared).  Those are the candidates that will very
// likely survive to the end, so rescore_kernel finds most of its work done: the 4 KB row gathers
// (HBM-bound, 5-6 ms at C2 when done after the last tile) hide behind tensor-bound scoring that
// uses 3 % of the DRAM bandwidth.  Positions below hi[q] are final (the scoring launch that
// appended them has completed: hi is a snapshot taken between launches), positions being
// appended concurrently are >= hi[q].
constexpr int kPreThreads = 128;

__global__ void __launch_bounds__(kPreThreads)
prescore_kernel(const float* __restrict__ X, int d, const float* __restrict__ Q, uint64_t* __restrict__ cand,
                const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, const float* __restrict__ est, int cap) {
  extern __shared__ __align__(16) float qv_pre[];
  const int64_t q = blockIdx.x;
  const uint32_t a = lo ? min(lo[q], (uint32_t)cap) : 0u;
  const uint32_t b = min(hi[q], (uint32_t)cap);
  const float e = est[q];
  if (a >= b || !(e > CMX_NEG_PAD)) return;
  for (int i = threadIdx.x; i < d; i += blockDim.x) qv_pre[i] = Q[q * (int64_t)d + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint64_t* buf = cand + q * (int64_t)cap;
  for (uint32_t base = a + (uint32_t)warp * 32u; base < b; base += (uint32_t)nwarps * 32u) {
    const uint32_t i = base + (uint32_t)lane;
    const uint64_t key = i < b ? buf[i] : 0ull;
    const bool want = key != 0ull && !key_is_exact(key) && key_score(key) >= e;
    uint32_t mask = __ballot_sync(0xffffffffu, want);
    while (mask) {
      const int j0 = __ffs(mask) - 1;
      mask &= mask - 1u;
      int j1 = j0;
      if (mask) { j1 = __ffs(mask) - 1; mask &= mask - 1u; }
      const uint32_t row0 = key_row(__shfl_sync(0xffffffffu, key, j0));
      const uint32_t row1 = key_row(__shfl_sync(0xffffffffu, key, j1));
      float s0, s1;
      exact_dot_pair<true>(X + (int64_t)row0 * d, X + (int64_t)row1 * d, qv_pre, d, lane, s0, s1);
      if (lane == 0) {
        buf[base + j0] = s0 > CMX_NEG_PAD ? make_key_exact(s0, row0) : 0ull;
        if (j1 != j0) buf[base + j1] = s1 > CMX_NEG_PAD ? make_key_exact(s1, row1) : 0ull;
      }
    }
  }
}

__global__ void snapshot_counts_kernel(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ snap, int64_t nq) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) snap[i] = cnt[i];
}

int launch_snapshot_counts(const SearchWs& ws, int64_t nq, uint32_t* snap, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  snapshot_counts_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(ws.cnt, snap, nq);
  CMX_LAUNCHED();
  return CMX_OK;
}

int prescore_smem_bytes(int d) { return (d * (int)sizeof(float) + 15) / 16 * 16; }
static int g_prescore_pad = 0;  // extra dynamic smem per CTA: bounds how many prescore CTAs share an SM (experiments)
void set_prescore_pad(int bytes) { g_prescore_pad = bytes < 0 ? 0 : bytes; }

int launch_prescore(const float* X, int d, const float* Q, const SearchWs& ws, int64_t nq, const uint32_t* lo,
                    const uint32_t* hi, cudaStream_t st) {
  if (nq == 0) return CMX_OK;
  int smem = prescore_smem_bytes(d);
  CMX_CHECK(smem <= kPrescoreMaxSmem, "prescore: d=%d too large", d);
  if (smem < g_prescore_pad) smem = g_prescore_pad;
  // same shared-memory carveout as the persistent scoring CTA it has to run beside
  static bool configured = false;
  if (!configured) {
    CMX_CUDA(cudaFuncSetAttribute(prescore_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CMX_CUDA(cudaFuncSetAttribute(prescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = true;
  }
  prescore_kernel<<<(unsigned)nq, kPreThreads, smem, st>>>(X, d, Q, ws.cand, lo, hi, ws.est, ws.cap);
  CMX_LAUNCHED();
  return CMX_OK;
}

// Rescore mode, last step.  One CTA per query: every surviving candidate (a superset of the
// exact top-k, selected on approximate fp16 tensor-core scores) gets its EXACT fp32 score
// from the fp32 row store, then the exact top-k under (score desc, row asc) is selected,
// sorted and written to D / I.
//  1. the candidates that pass the cut are gathered into shared memory (sharded search: the
//     cut is the GLOBAL k-th best approximate score minus the margin, so a shard keeps ~1/G
//     of its local band): the ones still carrying an approximate score at the front, the ones
//     prescore_kernel already made exact at the back;
//  2. one warp per TWO approximate candidates at a time (16 independent 128-bit loads in flight
//     per lane), exact_dot_pair -- deterministic for a given (query, row);
//  3. radix select + bitonic sort of the k best exact keys.
// dynamic smem: cap/2 keys | next_pow2(k) keys | d floats (the query): 44 KB at k = 1000,
// d = 1024, so four 512-thread CTAs share an SM and the row gathers of one hide the select
// phase of another.
__global__ void __launch_bounds__(kSelThreads, RESCORE_MIN_CTAS)
rescore_kernel(const float* __restrict__ X, int d, const float* __restrict__ Q, const uint64_t* __restrict__ cand,
               const uint32_t* __restrict__ cnt, uint32_t* __restrict__ overflow, const float* __restrict__ margin,
               const RescoreCut cut, int cap, int k, int topn, float* __restrict__ D, int64_t* __restrict__ I,
               int64_t id_base) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  __shared__ uint32_t s_ma, s_me;
  const int half = cap >> 1;
  uint64_t* top = keys + half;
  float* qv = reinterpret_cast<float*>(top + topn);
  const int64_t q = blockIdx.x;
  const uint32_t n_raw = cnt[q];
  if (threadIdx.x == 0) {
    s_ma = 0;
    s_me = 0;
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
    const bool ex = valid && key_is_exact(key);
    const bool ap = valid && !ex;
    const uint32_t b_ap = __ballot_sync(0xffffffffu, ap), b_ex = __ballot_sync(0xffffffffu, ex);
    uint32_t base_ap = 0, base_ex = 0;
    if (lane == 0) {
      if (b_ap) base_ap = atomicAdd(&s_ma, (uint32_t)__popc(b_ap));
      if (b_ex) base_ex = atomicAdd(&s_me, (uint32_t)__popc(b_ex));
    }
    base_ap = __shfl_sync(0xffffffffu, base_ap, 0);
    base_ex = __shfl_sync(0xffffffffu, base_ex, 0);
    const uint32_t lt = (1u << lane) - 1u;
    if (ap) {
      const uint32_t pos = base_ap + __popc(b_ap & lt);
      if (pos < (uint32_t)half) keys[pos] = key;
    } else if (ex) {
      const uint32_t pos = base_ex + __popc(b_ex & lt);  // from the back: half-1, half-2, ...
      if (pos < (uint32_t)half) keys[half - 1 - (int)pos] = key;
    }
  }
  __syncthreads();
  // fronts and backs collide when more than cap/2 keys pass: compact_kernel flags that first
  const bool band_ovf = s_ma + s_me > (uint32_t)half;
  if (band_ovf && threadIdx.x == 0) atomicOr(overflow, CMX_OVF_BAND);
  const int ma = (int)min(s_ma, (uint32_t)half);
  const int me = band_ovf ? 0 : (int)s_me;
  // 2. exact scores of the approximate keys, two rows per warp and iteration
  for (int i = 2 * warp; i < ma; i += 2 * nwarps) {
    const bool two = i + 1 < ma;
    const uint32_t row0 = key_row(keys[i]);
    const uint32_t row1 = two ? key_row(keys[i + 1]) : row0;
    float acc0, acc1;
    exact_dot_pair<false>(X + (int64_t)row0 * d, X + (int64_t)row1 * d, qv, d, lane, acc0, acc1);
    __syncwarp();
    if (lane == 0) {
      keys[i] = acc0 > CMX_NEG_PAD ? make_key_exact(acc0, row0) : 0ull;
      if (two) keys[i + 1] = acc1 > CMX_NEG_PAD ? make_key_exact(acc1, row1) : 0ull;
    }
  }
  __syncthreads();
  // the prescored keys join the list: move the back segment down to [ma, ma + me)
  // (destination index < source index for every element, rounds in ascending order)
  for (int j0 = 0; j0 < me; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    uint64_t v = 0ull;
    if (j < me) v = keys[half - me + j];
    __syncthreads();
    if (j < me) keys[ma + j] = v;
    __syncthreads();
  }
  const int m = ma + me;
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
                   int64_t id_base, cudaStream_t st, int spec_rank, int verify, int est_rank) {
  if (nq == 0) return CMX_OK;
  CMX_CHECK(ws.cap % 2 == 0, "compact: candidate capacity %d is odd (lists are loaded in 16-byte units)", ws.cap);
  const size_t smem = ((size_t)ws.cap + pow2_at_least(k)) * sizeof(uint64_t);
  CMX_CUDA(cudaFuncSetAttribute(compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  compact_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(ws.cand, ws.cnt, ws.tau, ws.overflow, ws.margin, ws.spec, ws.est,
                                                          ws.cap, k, final_pass, ws.spec ? spec_rank : 0,
                                                          ws.est ? est_rank : 0, ws.spec ? verify : 0, D, I, id_base);
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
// Block 0 also ORs the shards' status words (buffer overflow / cannot run two-phase) into this
// shard's `flag_any`: every shard reads the same words after the same barrier, so all of them take
// the same decision at the end of the step -- no host round trip, no collective.
struct UnionArgs {
  const float* parts[CMX_MAX_PEERS];
  float* outs[CMX_MAX_PEERS];
  const uint32_t* flags[CMX_MAX_PEERS];
};

__global__ void __launch_bounds__(kSelThreads)
union_kth_kernel(const UnionArgs a, int nparts, int nouts, int nflags, uint32_t* __restrict__ flag_any, int64_t q0,
                 int64_t nq_slice, int k) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  if (blockIdx.x == 0 && threadIdx.x == 0 && flag_any != nullptr) {
    uint32_t f = 0;
    for (int g = 0; g < nflags; ++g) f |= a.flags[g][0];
    if (f) atomicOr(flag_any, f);
  }
  if ((int64_t)blockIdx.x >= nq_slice) return;
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
                     const uint32_t* const* flags, int nflags, uint32_t* flag_any, cudaStream_t st) {
  const bool have_flags = flags != nullptr && nflags > 0 && flag_any != nullptr;
  if (q1 <= q0 && !have_flags) return CMX_OK;
  CMX_CHECK(nparts >= 1 && nparts <= CMX_MAX_PEERS && nouts >= 1 && nouts <= CMX_MAX_PEERS && nflags <= CMX_MAX_PEERS,
            "union_kth: at most %d parts", CMX_MAX_PEERS);
  const size_t smem = (size_t)nparts * k * sizeof(uint64_t);
  CMX_CHECK(smem <= 200 * 1024, "union_kth: nparts*k = %d too large", nparts * k);
  UnionArgs a;
  for (int g = 0; g < nparts; ++g) a.parts[g] = parts[g];
  for (int o = 0; o < nouts; ++o) a.outs[o] = outs[o];
  for (int g = 0; g < (have_flags ? nflags : 0); ++g) a.flags[g] = flags[g];
  CMX_CUDA(cudaFuncSetAttribute(union_kth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nslice = q1 > q0 ? q1 - q0 : 0;
  const unsigned grid = (unsigned)(nslice > 0 ? nslice : 1);
  union_kth_kernel<<<grid, kSelThreads, smem, st>>>(a, nparts, nouts, have_flags ? nflags : 0, have_flags ? flag_any : nullptr,
                                                    q0, nslice, k);
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
// single-shard order exactly.  dynamic smem: nparts*k keys, then next_pow2(k) keys.
// Src tells where entry `pos` of part g of query q lives and where the merged row goes:
// GatheredSrc = parts stacked in one buffer [nparts, nq, k] (after an all_gather),
// PeerSrc = one pointer per part, each possibly resident on another GPU.
struct GatheredSrc {
  const float* Dp;
  const int64_t* Ip;
  float* D;
  int64_t* I;
  int64_t nq;
  __device__ __forceinline__ float score(int g, int64_t q, int k, int pos) const { return Dp[((int64_t)g * nq + q) * k + pos]; }
  __device__ __forceinline__ int64_t id(int g, int64_t q, int k, int pos) const { return Ip[((int64_t)g * nq + q) * k + pos]; }
  __device__ __forceinline__ void store(int64_t q, int k, int i, float s, int64_t v) const {
    D[q * k + i] = s;
    I[q * k + i] = v;
  }
};

struct PeerSrc {
  const float* D[CMX_MAX_PEERS];
  const int64_t* I[CMX_MAX_PEERS];
  float* Do[CMX_MAX_PEERS];
  int64_t* Io[CMX_MAX_PEERS];
  int nouts;
  __device__ __forceinline__ float score(int g, int64_t q, int k, int pos) const { return D[g][q * k + pos]; }
  __device__ __forceinline__ int64_t id(int g, int64_t q, int k, int pos) const { return I[g][q * k + pos]; }
  __device__ __forceinline__ void store(int64_t q, int k, int i, float s, int64_t v) const {
    for (int o = 0; o < nouts; ++o) {
      Do[o][q * k + i] = s;
      Io[o][q * k + i] = v;
    }
  }
};

template <typename Src>
__device__ __forceinline__ void merge_query(const Src& src, int nparts, int64_t q, int k, uint64_t* keys, SelectShared& sh,
                                            uint32_t& s_valid) {
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
        valid = src.id(g, q, k, pos) >= 0;
        if (valid) key = make_key(src.score(g, q, k, pos), (uint32_t)i);
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
      s = src.score(g, q, k, pos);
      id = src.id(g, q, k, pos);
    }
    src.store(q, k, i, s, id);
  }
}

__global__ void __launch_bounds__(kSelThreads)
merge_kernel(const GatheredSrc src, int nparts, int k) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  __shared__ uint32_t s_valid;
  merge_query(src, nparts, (int64_t)blockIdx.x, k, keys, sh, s_valid);
}

// Fused exchange + merge over peer memory: part g's list lives in GPU g's memory and is read
// here directly through NVLink peer pointers (no all_gather, no staging copy); the merged
// rows of queries [q0, q1) are stored into every output buffer (peer stores into the ranks'
// device buffers, or posted PCIe writes into a device-mapped pinned HOST buffer shared by the
// ranks), so after one cross-rank barrier the full result is in place.
__global__ void __launch_bounds__(kSelThreads)
merge_peers_kernel(const PeerSrc src, int nparts, int64_t q0, int k) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ SelectShared sh;
  __shared__ uint32_t s_valid;
  merge_query(src, nparts, q0 + blockIdx.x, k, keys, sh, s_valid);
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
  PeerSrc a;
  for (int g = 0; g < nparts; ++g) { a.D[g] = D_parts[g]; a.I[g] = I_parts[g]; }
  for (int o = 0; o < nouts; ++o) { a.Do[o] = D_outs[o]; a.Io[o] = I_outs[o]; }
  a.nouts = nouts;
  CMX_CUDA(cudaFuncSetAttribute(merge_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_peers_kernel<<<(unsigned)(q1 - q0), kSelThreads, smem, st>>>(a, nparts, q0, k);
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
  GatheredSrc src{D_parts, I_parts, D, I, nq};
  merge_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(src, nparts, k);
  CMX_LAUNCHED();
  return CMX_OK;
}

}  // namespace cmx

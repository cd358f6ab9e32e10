// Tensor-core scorer: Q x Corpus^T with a fused threshold-filter epilogue (sm_100a).
//
// Replaces the SGEMM + blockSelect pair behind index.search for batched queries
// (reference call sites onepass_dense_mix_run_custom_lang.py:878,
// onepass_bilingual_mix_hub_custom_lang.py:950).
//
// Arithmetic, PASSES = 1 (default "rescore" precision): ONE tcgen05 kind::f16 MMA pass over the
// fp16 hi planes, hi = f16(x*2^e) (prologue.cu), fp32 accumulation in TMEM.  The scores are
// approximate with a rigorous per-query error bound; the selection keeps a margin band and
// select.cu:rescore_kernel computes exact fp32 scores for the survivors.
// PASSES = 3 ("split" precision): every operand element is two fp16 numbers hi, lo =
// f16(x*2^e - hi), x*2^e = hi + lo up to 2^-22 relative, and a score is accumulated by three
// MMAs per k-step into one accumulator
//        D += Qlo*Bhi ;  D += Qhi*Blo ;  D += Qhi*Bhi
// (products of fp16 pairs are exact in fp32; only lo*lo ~ 2^-22 is dropped).  Either way the
// accumulator is rescaled by the exact power of two 2^-(eq+eb) in the epilogue.
//
// Structure (one persistent CTA per SM, 320 threads -- 192 in the CTA-pair and small-batch kernels --, warp-specialised):
//   warp 0    TMA producer: cp.async.bulk.tensor 2D tiles (128B swizzle) of Qhi(/Qlo)
//             [128 x 64] and Bhi(/Blo) [BN x 64] into a ring of smem stages; corpus tiles
//             are visited in a golden-ratio block order (TcParams::perm)
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16)
//   warps 2-5 epilogue: tcgen05.ld the 128 x BN fp32 accumulator (one query row per
//   (and 6-9) thread), compare with the row's threshold tau and append survivors
//             (score,row keys) to the query's candidate buffer -- the score matrix
//             never leaves the SM.  The single-CTA kernel has two such groups, one per
//             accumulator buffer.
// Accumulators are double-buffered in TMEM (2 x BN columns) so the epilogue of tile
// i overlaps the MMAs of tile i+1.  Tiles are ordered m-fastest so that CTAs running
// concurrently share the same corpus tile through L2: each corpus byte is read from
// HBM once per pass.
#include <cuda.h>

#include "common.cuh"

namespace cmx {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // fp16 elements = 128 bytes = one swizzle-128B row
constexpr int TC_THREADS = 192;
// the single-CTA scoring kernel runs TWO epilogue warpgroups (warps 2-5 and 6-9), one per TMEM accumulator buffer:
// a group then has two MMA tile times for its tile.  The epilogue of the early slabs (a few percent of all scores
// pass) is a lone warp per SM sub-partition issuing ~3500 dependent instructions per tile at IPC 0.25: 14 k cycles
// against 8.2 k of MMAs (ncu on the 84 k-row mid slab of a 1.1 M-row shard: tensor pipe 59 % active, the MMA warp
// waiting for a free accumulator).
constexpr int TC_THREADS1 = 64 + 2 * 128;

// -DCMX_TC_TIMERS: cycle counters of the single-CTA kernel's roles, summed over CTAs (scripts/exp_tc_timers.py).
// 0 MMA: wait for a free accumulator   1 MMA: wait for operands   2 MMA: tiles
// 4 epilogue (warp 2, lane 0): wait for the accumulator   5 accumulator held   8 after release   9 tiles
// 10 pool emptied in place (all epilogue warps)   11 chunks appended lane by lane
#ifdef CMX_TC_TIMERS
__device__ unsigned long long g_tc_timers[16];
#define TCT_NOW() clock64()
#define TCT_ADD(i, t0) atomicAdd(&g_tc_timers[i], (unsigned long long)(clock64() - (t0)))
#define TCT_COUNT(i) atomicAdd(&g_tc_timers[i], 1ull)
#else
#define TCT_COUNT(i) ((void)0)
#define TCT_NOW() 0ll
#define TCT_ADD(i, t0) ((void)(t0))
#endif
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KB per Q plane tile
constexpr int TC_TILE_SLOTS = 4;               // depth of the tile-id ring of the dynamic scheduler

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// L2 cache-policy descriptors (the fixed encodings CUTLASS's TMA::CacheHintSm90 uses)
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                                 int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void st_stream_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// `seen` is the producer's last observation of the counter.  It can only lag behind the truth, so a tile it admits
// is admitted; the fresh read that follows is not waited for -- its round trip to L2 (~700 cycles, once per tile,
// between the last operand request of one tile and the first of the next) overlaps the tile's TMA requests and the
// value is first looked at when the next tile starts.
__device__ __forceinline__ void throttle_wait(const unsigned long long* done, long long t, long long window,
                                              unsigned long long& seen) {
  if (window <= 0) return;
  while (!(t < (long long)seen + window)) {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(done) : "memory");
    if (t < (long long)seen + window) break;
    __nanosleep(200);
  }
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(done) : "memory");
}
__device__ __forceinline__ void throttle_tile_done(unsigned long long* done) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(done) : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (sm_100 "version 1"):
// start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B (8 rows x
// 128 B) | version = 1 | layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct TcParams {
  int64_t row0;        // first row POSITION of the slab in processing order
  int64_t nrows;       // row positions in the slab (whole 256-row blocks, or a piece of one block)
  // Processing order (perm_row below): the corpus is cut into blocks of 256 rows and block
  // position j is corpus block (j * perm) mod nblk, with perm ~ 0.618 * nblk coprime to nblk.
  // Every prefix of that order is spread evenly over the whole corpus, so the thresholds the
  // first slabs establish are estimates for ALL rows -- not just for the head of the file.
  // (A combined EN+ZH index stores one language after the other: in file order the thresholds
  // learnt on the first language would let the second one overflow the candidate buffers.)
  int64_t nblk;        // ceil(ntotal / 256)
  uint64_t perm;       // block multiplier, 1 = file order
  int64_t nvalid;      // ntotal: rows at or beyond it are padding
  int kblocks;         // d_pad / 64
  int mtiles;          // ceil(nq / 128)
  int64_t ntiles;      // mtiles * ceil(nrows / BN)
  int64_t nq;
  const float* q_inv_scale;  // device: 1 / query scale
  float b_inv_scale;         // 1 / corpus scale
  const float* tau;
  uint32_t* cnt;
  uint64_t* cand;
  int cap;
  int dense;
  int flags;  // bit0: Q tiles evict_last, bit1: corpus tiles evict_first, bit2: streaming (.cs) appends
  // Progress throttle: tiles are assigned round-robin (m fastest) so that CTAs running at the
  // same time share corpus tiles through L2 -- but in a long launch CTAs drift apart (they are
  // paced by memory latency, not by a clock) and the set of corpus tiles in flight outgrows
  // L2: ncu showed 8x the algorithmic DRAM reads on a 3 M-row slab.  `done` counts finished
  // tiles; a producer does not start tile t before t < done + window.
  unsigned long long* done;
  long long window;
  int two_pass;  // 1: count-then-store epilogue (dense early slabs), 0: staged epilogue
};

// corpus row of row position `pos_row` (see TcParams::perm)
__device__ __forceinline__ int64_t perm_row(const TcParams& p, int64_t pos_row) {
  const uint64_t pos = (uint64_t)pos_row >> 8;
  const uint64_t blk = (pos * p.perm) % (uint64_t)p.nblk;  // pos, perm < 2^24
  return (int64_t)(blk << 8) + (pos_row & 255);
}

// Dense epilogue (first slab): every column of the calling thread's query row becomes a
// candidate, written at its own position of the query's buffer (padding rows: null keys).
template <int BN>
__device__ __forceinline__ void epilogue_dense_tile(const TcParams& p, uint32_t taddr_row, int64_t q, bool qvalid,
                                                    float inv, int64_t tile_row0, int64_t cols_valid, int64_t slot0) {
  uint64_t* qcand = p.cand + q * (int64_t)p.cap + slot0;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t v[32];
    tmem_ld_x32(taddr_row + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    const int jmax = (int)min((int64_t)32, cols_valid - c * 32);
    if (qvalid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = __uint_as_float(v[j]) * inv;
        qcand[c * 32 + j] = (j < jmax && s > CMX_NEG_PAD) ? make_key(s, (uint32_t)(tile_row0 + c * 32 + j)) : 0ull;
      }
    }
  }
}

// ---- branch-free survivor extraction ------------------------------------------------
// Each epilogue warp is alone on its SM sub-partition, so nothing hides the latency of a
// dependent branch: 32 "if (score > tau) {...}" blocks per 32-column chunk cost ~60 cycles
// each whenever ANY lane of the warp has a survivor in the chunk (measured: the early slabs
// ran at half the MMA rate).  Instead: a 32-bit survivor mask per thread (independent
// compares), and for each set bit the value is fetched with a 31-select tree -- registers
// cannot be indexed dynamically.
__device__ __forceinline__ uint32_t survivor_mask(const uint32_t (&v)[32], float tau_raw, int jmax) {
  uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;  // four chains instead of one 32-deep dependency
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m0 |= (__uint_as_float(v[j]) > tau_raw) ? (1u << j) : 0u;
    m1 |= (__uint_as_float(v[8 + j]) > tau_raw) ? (1u << (8 + j)) : 0u;
    m2 |= (__uint_as_float(v[16 + j]) > tau_raw) ? (1u << (16 + j)) : 0u;
    m3 |= (__uint_as_float(v[24 + j]) > tau_raw) ? (1u << (24 + j)) : 0u;
  }
  uint32_t m = (m0 | m1) | (m2 | m3);
  if (jmax < 32) m &= (jmax <= 0) ? 0u : ((1u << jmax) - 1u);  // columns past the corpus end
  return m;
}

__device__ __forceinline__ uint32_t select32(const uint32_t (&v)[32], int j) {
  const bool b0 = j & 1, b1 = j & 2, b2 = j & 4, b3 = j & 8, b4 = j & 16;
  uint32_t a[16], b[8], c[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = b0 ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = b1 ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = b2 ? b[2 * i + 1] : b[2 * i];
  const uint32_t d0 = b3 ? c[1] : c[0], d1 = b3 ? c[3] : c[2];
  return b4 ? d1 : d0;
}

// Filter epilogue of one 128 x BN accumulator tile for the calling thread's query row.
// ONE pass over the row in TMEM: columns that beat the threshold are parked in a small
// per-thread staging area in shared memory (value + column).  The caller then releases the
// TMEM buffer to the MMA warp and only afterwards calls epilogue_flush, which reserves room
// in the query's candidate buffer with ONE atomicAdd and stores the keys contiguously -- the
// atomic's round trip is off the TMEM critical path and a warp pays it once per tile, not
// once per surviving column of any of its lanes.  When a thread's staging area fills up
// (dense early slabs) it is flushed in place.
template <int BN>
__device__ __forceinline__ void epilogue_filter_tile_two_pass(const TcParams& p, uint32_t taddr_row, int64_t q,
                                                              float tau_raw, float inv, int64_t tile_row0,
                                                              int64_t cols_valid) {
  // dense early slabs (tens of survivors per row and tile): count first, reserve with ONE
  // atomicAdd, then re-read the chunks that had survivors and store them contiguously
  constexpr int NC = BN / 32;
  uint32_t total = 0;
  uint32_t chunk_bits = 0;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(taddr_row + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    const int jmax = (int)min((int64_t)32, cols_valid - c * 32);
    const uint32_t n = (uint32_t)__popc(survivor_mask(v, tau_raw, jmax));
    total += n;
    if (n) chunk_bits |= 1u << c;
  }
  if (!__any_sync(0xffffffffu, total != 0)) return;
  uint32_t pos = 0;
  if (total) pos = atomicAdd(&p.cnt[q], total);
  uint64_t* qcand = p.cand + q * (int64_t)p.cap;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const bool mine = (chunk_bits >> c) & 1u;
    if (!__any_sync(0xffffffffu, mine)) continue;
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(taddr_row + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    if (mine) {
      const int jmax = (int)min((int64_t)32, cols_valid - c * 32);
      uint32_t mask = survivor_mask(v, tau_raw, jmax);
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1u;
        const float raw = __uint_as_float(select32(v, j));
        if (pos < (uint32_t)p.cap) qcand[pos] = make_key(raw * inv, (uint32_t)(tile_row0 + c * 32 + j));
        ++pos;
      }
    }
  }
}

constexpr int TC_STAGE_SLOTS = 8;
constexpr int TC_EPI_THREADS = 128;
constexpr int TC_EPI_GROUP_SMEM = TC_STAGE_SLOTS * TC_EPI_THREADS * 8;  // float value + int column per slot
constexpr int TC_POOL = 512;                                            // survivor pool entries per epilogue warp
constexpr int TC_EPI_SMEM = 8 * TC_POOL * 8;  // eight warp pools (single-CTA kernel, early slabs); the staging slots of the
                                              // two epilogue groups alias them (a launch runs one kind of epilogue)
static_assert(2 * TC_EPI_GROUP_SMEM <= TC_EPI_SMEM, "staging slots must fit the epilogue scratch");

struct EpiStage {
  float* val;  // [TC_STAGE_SLOTS][TC_EPI_THREADS]
  int* col;    // [TC_STAGE_SLOTS][TC_EPI_THREADS]
  int e;       // this thread's index among the epilogue threads
};

__device__ __forceinline__ void epilogue_flush(const TcParams& p, const EpiStage& st, int& nst, int64_t q, float inv,
                                               int64_t tile_row0) {
  if (nst == 0) return;
  const uint32_t pos = atomicAdd(&p.cnt[q], (uint32_t)nst);
  uint64_t* qcand = p.cand + q * (int64_t)p.cap;
  for (int i = 0; i < nst; ++i) {
    if (pos + i < (uint32_t)p.cap) {
      const uint64_t key = make_key(st.val[i * TC_EPI_THREADS + st.e] * inv,
                                    (uint32_t)(tile_row0 + st.col[i * TC_EPI_THREADS + st.e]));
      if (p.flags & 4) st_stream_u64(qcand + pos + i, key);
      else qcand[pos + i] = key;
    }
  }
  nst = 0;
}

// staging area full in the middle of a row (dense early slabs only): one out-of-line copy so
// that the 32-way unrolled compare loop stays small
__device__ __noinline__ void epilogue_flush_full(uint32_t* cnt_q, uint64_t* qcand, int cap, const float* val, const int* col,
                                                 float inv, int64_t tile_row0) {
  const uint32_t pos = atomicAdd(cnt_q, (uint32_t)TC_STAGE_SLOTS);
  for (int i = 0; i < TC_STAGE_SLOTS; ++i)
    if (pos + i < (uint32_t)cap)
      qcand[pos + i] = make_key(val[i * TC_EPI_THREADS] * inv, (uint32_t)(tile_row0 + col[i * TC_EPI_THREADS]));
}

template <int BN>
__device__ __forceinline__ void epilogue_filter_tile(const TcParams& p, uint32_t taddr_row, const EpiStage& st, int& nst,
                                                     int64_t q, float tau_raw, float inv, int64_t tile_row0,
                                                     int64_t cols_valid) {
  constexpr int NC = BN / 32;
  // rolled loop: the whole epilogue must fit the instruction cache (a fully unrolled version
  // was 180 KB of SASS and ran 2.4x slower than the MMAs it hides behind)
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(taddr_row + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    const int jmax = (int)min((int64_t)32, cols_valid - c * 32);
    uint32_t mask = survivor_mask(v, tau_raw, jmax);
    while (mask) {  // lanes without a survivor skip; the warp runs max-popcount iterations
      const int j = __ffs(mask) - 1;
      mask &= mask - 1u;
      const uint32_t raw = select32(v, j);
      if (nst == TC_STAGE_SLOTS) {
        epilogue_flush_full(&p.cnt[q], p.cand + q * (int64_t)p.cap, p.cap, st.val + st.e, st.col + st.e, inv, tile_row0);
        nst = 0;
      }
      st.val[nst * TC_EPI_THREADS + st.e] = __uint_as_float(raw);
      st.col[nst * TC_EPI_THREADS + st.e] = c * 32 + j;
      ++nst;
    }
  }
}

// ---- pool epilogue (early slabs of the single-CTA kernel) ------------------------------------
// Early slabs let a few percent of all scores pass (tens of survivors per query row and tile).  The TMEM accumulator
// can only be handed back to the MMA warp when the epilogue has read it for the last time, and with two accumulator
// buffers MMA(i) waits for epilogue(i - 2): whatever the epilogue does while it holds the buffer beyond one MMA tile
// time (8192 cycles) is exposed.  The count-then-store epilogue above holds it for ~12 k cycles on a 4 %-pass slab
// (cycle counters, scripts/exp_tc_timers.py: first pass 2.3 k, second pass 9.9 k -- a divergent per-lane loop over
// survivors, ~250 cycles per trip).  Here the warp walks the 32 columns of a chunk TOGETHER: one ballot per column
// (statically indexed register, no select tree, no divergence), survivors appended to a per-warp pool in shared
// memory as (raw value, column | owner lane | rank within the owner's row).  The buffer is released right after
// the single pass; the per-row atomicAdd and the global stores follow, all 32 lanes busy, outside the critical path.
struct EpiPool {
  float* val;      // [TC_POOL]
  uint32_t* meta;  // [TC_POOL]: column (8 bits) | owner lane << 8 | rank << 13
};

__device__ __forceinline__ void epilogue_pool_store(uint32_t* cnt, uint64_t* cand, int cap, const float* val, const uint32_t* meta_a,
                                                    int lane, int64_t q0, uint32_t mycnt, uint32_t fill, float inv, int64_t tile_row0) {
  uint32_t pos = 0;
  if (mycnt) pos = atomicAdd(&cnt[q0 + lane], mycnt);
  __syncwarp();  // the other lanes' entries are visible
  for (uint32_t i0 = 0; i0 < fill; i0 += 32) {
    const uint32_t i = i0 + (uint32_t)lane;
    const bool on = i < fill;
    const uint32_t meta = on ? meta_a[i] : 0u;
    const float raw = on ? val[i] : 0.f;
    const int owner = (int)((meta >> 8) & 31u);
    const uint32_t at = __shfl_sync(0xffffffffu, pos, owner) + (meta >> 13);
    if (on && at < (uint32_t)cap)
      cand[(q0 + owner) * (int64_t)cap + at] = make_key(raw * inv, (uint32_t)(tile_row0 + (meta & 255u)));
  }
  __syncwarp();  // before the pool is refilled
}
__device__ __forceinline__ void epilogue_pool_flush(const TcParams& p, const EpiPool& pl, int lane, int64_t q0,
                                                    uint32_t& mycnt, uint32_t& fill, float inv, int64_t tile_row0) {
  if (fill == 0) return;  // warp-uniform
  epilogue_pool_store(p.cnt, p.cand, p.cap, pl.val, pl.meta, lane, q0, mycnt, fill, inv, tile_row0);
  fill = 0;
  mycnt = 0;
}
// Out-of-line pieces of the overflow path (the unrolled column loop must not contain calls: the 32 accumulator
// registers would live on the stack).
__device__ __noinline__ void epilogue_pool_store_full(uint32_t* cnt, uint64_t* cand, int cap, const float* val, int lane, int64_t q0,
                                                      uint32_t mycnt, uint32_t fill, float inv, int64_t tile_row0) {
  epilogue_pool_store(cnt, cand, cap, val, reinterpret_cast<const uint32_t*>(val + TC_POOL), lane, q0, mycnt, fill, inv, tile_row0);
}
// one chunk whose survivors alone exceed the pool (more than half of its 1024 scores pass): every lane appends its own
__device__ __noinline__ void epilogue_chunk_direct(uint32_t* cnt_q, uint64_t* qcand, int cap, uint32_t taddr_chunk, float tau_raw,
                                                   float inv, int64_t row_c, int jmax) {
  uint32_t v[32];
  __syncwarp();
  tmem_ld_x32(taddr_chunk, v);
  tmem_ld_wait();
  uint32_t mask = survivor_mask(v, tau_raw, jmax);
  if (!mask) return;
  uint32_t pos = atomicAdd(cnt_q, (uint32_t)__popc(mask));
  while (mask) {
    const int j = __ffs(mask) - 1;
    mask &= mask - 1u;
    if (pos < (uint32_t)cap) qcand[pos] = make_key(__uint_as_float(select32(v, j)) * inv, (uint32_t)(row_c + j));
    ++pos;
  }
}

// Dense epilogue through the warp pool (single-CTA kernel).  epilogue_dense_tile has every lane store its own row's
// keys: one store instruction touches 32 rows 8 bytes each -- 32 sectors, a quarter used -- and the first slab was
// bound by that (53 k cycles per tile, ncu: stall_lg).  Here 8 columns at a time are transposed through shared
// memory (row stride 9 keys: conflict-free both ways), so that a store instruction writes 64 contiguous bytes for
// each of 4 rows: 8 sectors, all used.
template <int BN>
__device__ __forceinline__ void epilogue_dense_tile_pool(const TcParams& p, uint32_t taddr_row, uint64_t* tr, int lane, int64_t q0,
                                                         float inv, int64_t tile_row0, int64_t cols_valid, int64_t slot0) {
  static_assert(32 * 9 * 8 <= TC_POOL * 8, "transposition buffer must fit the warp pool");
  const int sub = lane >> 3, jj = lane & 7;  // read side: row sub (+4 per trip), column jj of the piece
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(taddr_row + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    const int jmax = (int)min((int64_t)32, cols_valid - c * 32);
#pragma unroll
    for (int piece = 0; piece < 4; ++piece) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = piece * 8 + u;
        const float s = __uint_as_float(v[j]) * inv;
        tr[lane * 9 + u] = (j < jmax && s > CMX_NEG_PAD) ? make_key(s, (uint32_t)(tile_row0 + c * 32 + j)) : 0ull;
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + sub;
        const uint64_t key = tr[row * 9 + jj];
        if (q0 + row < p.nq) p.cand[(q0 + row) * (int64_t)p.cap + slot0 + c * 32 + piece * 8 + jj] = key;
      }
      __syncwarp();
    }
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_filter_tile_pool(const TcParams& p, uint32_t taddr_row, const EpiPool& pl, int lane,
                                                          int64_t q0, float tau_raw, float inv, int64_t tile_row0,
                                                          int64_t cols_valid, uint32_t& mycnt, uint32_t& fill) {
  constexpr int NC = BN / 32;
  const uint32_t lt = (1u << lane) - 1u;
  int c = 0;
#pragma unroll 1
  while (c < NC) {
    const int jmax = (int)min((int64_t)32, cols_valid - c * 32);
    if (jmax <= 0) break;  // warp-uniform: columns past the corpus end
    const uint32_t fill0 = fill, cnt0 = mycnt;
    {
      uint32_t v[32];
      __syncwarp();
      tmem_ld_x32(taddr_row + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      const uint32_t meta_c = (uint32_t)(c * 32) | ((uint32_t)lane << 8);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool pass = __uint_as_float(v[j]) > tau_raw && j < jmax;
        const uint32_t b = __ballot_sync(0xffffffffu, pass);
        const uint32_t slot = fill + (uint32_t)__popc(b & lt);
        if (pass && slot < (uint32_t)TC_POOL) {
          pl.val[slot] = __uint_as_float(v[j]);
          pl.meta[slot] = (meta_c + (uint32_t)j) | (mycnt << 13);
        }
        mycnt += pass ? 1u : 0u;
        fill += (uint32_t)__popc(b);
      }
    }
    if (fill <= (uint32_t)TC_POOL) { ++c; continue; }
    // the chunk did not fit (warp-uniform, rare): forget it, empty the pool in place, and take the chunk again
    fill = fill0;
    mycnt = cnt0;
    if (fill0 == 0) {
      if (lane == 0) TCT_COUNT(11);
      epilogue_chunk_direct(&p.cnt[q0 + lane], p.cand + (q0 + lane) * (int64_t)p.cap, p.cap, taddr_row + (uint32_t)(c * 32), tau_raw, inv,
                            tile_row0 + c * 32, jmax);
      ++c;
      continue;
    }
    if (lane == 0) TCT_COUNT(10);
    epilogue_pool_store_full(p.cnt, p.cand, p.cap, pl.val, lane, q0, mycnt, fill, inv, tile_row0);
    fill = 0;
    mycnt = 0;
  }
}

// PASSES = 3: split precision (Qlo*Bhi + Qhi*Blo + Qhi*Bhi); PASSES = 1: Qhi*Bhi only
// (approximate filter scores of the rescore mode; the lo tiles are neither staged nor read)
template <int BN, int STAGES, int PASSES>
__global__ void __launch_bounds__(TC_THREADS1, 1)
tc_score_kernel(const __grid_constant__ CUtensorMap tmQhi, const __grid_constant__ CUtensorMap tmQlo,
                const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                const TcParams p) {
  constexpr bool SPLIT = PASSES == 3;
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int OFF_QLO = TC_A_BYTES;                       // split only
  constexpr int OFF_BHI = SPLIT ? 2 * TC_A_BYTES : TC_A_BYTES;
  constexpr int OFF_BLO = OFF_BHI + B_BYTES;                // split only
  constexpr int STAGE_BYTES = SPLIT ? 2 * TC_A_BYTES + 2 * B_BYTES : TC_A_BYTES + B_BYTES;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  static_assert(2 * BN <= 512, "two accumulator buffers must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SW128 needs 1024B alignment
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  // tile-id ring (dynamic scheduler): full[TC_TILE_SLOTS], empty[TC_TILE_SLOTS], then the ids
  auto tid_full_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 5 + i); };
  auto tid_empty_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 5 + TC_TILE_SLOTS + i); };
  volatile int* tile_ring = reinterpret_cast<volatile int*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 8u * (2 * STAGES + 5 + 2 * TC_TILE_SLOTS));
  static_assert(8 * (2 * STAGES + 5 + 2 * TC_TILE_SLOTS) + 4 * TC_TILE_SLOTS <= 256, "barrier block overflows its 256 bytes");
  uint8_t* epi_smem = smem_raw + (bar_base - smem_u32(smem_raw)) + 256;  // after the barrier block

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    for (int i = 0; i < TC_TILE_SLOTS; ++i) { mbar_init(tid_full_bar(i), 1); mbar_init(tid_empty_bar(i), 5); }  // consumers: MMA thread + 4 epilogue warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmQhi); tma_prefetch_desc(&tmQlo);
    tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int mtiles = p.mtiles;

  // Two tile schedulers (p.window selects).  Static (default): tile t = blockIdx + i * gridDim, m fastest, with the
  // progress throttle of TcParams.  Dynamic (p.window == 0): tiles are handed out in order from ONE global counter;
  // the producer publishes each tile id to the MMA thread and the epilogue warps through a small ring, -1 ends the
  // launch.  A CTA that is slowed down (a co-resident prescore CTA) then takes fewer tiles instead of holding
  // everybody back -- with prescoring on that cut its cost from +10 ms to +1..5 ms per C2 step -- but when nothing
  // shares the SMs the static order is 3 % faster (all CTAs of the dynamic order hit the same 2-3 corpus tiles in
  // L2 at the same moment), so static stays the default (profiles/r02_scheduler_ab.md).
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int slot = 0;
      uint32_t sphase = 0;
      const uint64_t pol_q = (p.flags & 1) ? kL2EvictLast : kL2EvictNormal;
      const uint64_t pol_b = (p.flags & 2) ? kL2EvictFirst : kL2EvictNormal;
      const bool dyn = p.window == 0;  // window > 0: static round-robin tiles + progress throttle (round-1 scheduler, kept for A/B)
      unsigned long long done_seen = 0;
      unsigned long long t_next = dyn ? atomicAdd(p.done, 1ull) : (unsigned long long)blockIdx.x;
      for (;;) {
        const int64_t t = (int64_t)t_next;
        const bool live = t < p.ntiles;
        if (dyn) {
          // the end marker goes into two consecutive slots: one for each epilogue group (slots alternate between them)
          for (int rep = 0; rep < (live ? 1 : 2); ++rep) {
            mbar_wait(tid_empty_bar(slot), sphase ^ 1u);
            tile_ring[slot] = live ? (int)t : -1;
            mbar_arrive(tid_full_bar(slot));  // release: the id is visible to whoever sees this phase complete
            if (++slot == TC_TILE_SLOTS) { slot = 0; sphase ^= 1u; }
          }
        }
        if (!live) break;
        if (dyn) t_next = atomicAdd(p.done, 1ull);  // the next id travels while this tile's loads are issued
        else { t_next = (unsigned long long)(t + gridDim.x); throttle_wait(p.done, t, p.window, done_seen); }
        const int m = (int)(t % mtiles);
        const int64_t n = t / mtiles;
        const int32_t qrow = m * TC_BM;
        const int32_t brow = (int32_t)perm_row(p, p.row0 + n * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d_hint(sbase, &tmQhi, full_bar(stage), kb * TC_BK, qrow, pol_q);
          tma_load_2d_hint(sbase + OFF_BHI, &tmBhi, full_bar(stage), kb * TC_BK, brow, pol_b);
          if (SPLIT) {
            tma_load_2d_hint(sbase + OFF_QLO, &tmQlo, full_bar(stage), kb * TC_BK, qrow, pol_q);
            tma_load_2d_hint(sbase + OFF_BLO, &tmBlo, full_bar(stage), kb * TC_BK, brow, pol_b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int slot = 0;
      uint32_t sphase = 0;
      const bool dyn = p.window == 0;
      int64_t t_static = blockIdx.x;
      for (;;) {
        if (dyn) {
          mbar_wait(tid_full_bar(slot), sphase);
          const int t = tile_ring[slot];
          mbar_arrive(tid_empty_bar(slot));
          if (++slot == TC_TILE_SLOTS) { slot = 0; sphase ^= 1u; }
          if (t < 0) break;
        } else {
          if (t_static >= p.ntiles) break;
          t_static += gridDim.x;
        }
        [[maybe_unused]] long long tw = TCT_NOW();
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        TCT_ADD(0, tw);
        TCT_COUNT(2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tw = TCT_NOW();
          mbar_wait(full_bar(stage), phase);
          TCT_ADD(1, tw);
          tc_fence_after();
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          const uint64_t qhi = make_smem_desc(sbase);
          const uint64_t qlo = make_smem_desc(sbase + OFF_QLO);
          const uint64_t bhi = make_smem_desc(sbase + OFF_BHI);
          const uint64_t blo = make_smem_desc(sbase + OFF_BLO);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);  // 16 fp16 = 32 bytes per k-step
            const uint32_t first = (kb | k) != 0 ? 1u : 0u;
            if (SPLIT) {
              tc_mma_f16(d_tmem, qlo + koff, bhi + koff, IDESC, first);
              tc_mma_f16(d_tmem, qhi + koff, blo + koff, IDESC, 1u);
              tc_mma_f16(d_tmem, qhi + koff, bhi + koff, IDESC, 1u);
            } else {
              tc_mma_f16(d_tmem, qhi + koff, bhi + koff, IDESC, first);
            }
          }
          tc_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: threshold filter + append =====================
    const int lane_base = (warp & 3) * 32;  // TMEM lanes this warp may touch
    const float inv = p.q_inv_scale[0] * p.b_inv_scale;
    const float fwd = 1.0f / inv;  // power of two
    // group g (warps 2-5: 0, warps 6-9: 1) takes this CTA's tiles g, g + 2, ...: always accumulator buffer g
    const int grp = (warp - 2) >> 2;
    EpiStage stg;
    stg.val = reinterpret_cast<float*>(epi_smem + grp * TC_EPI_GROUP_SMEM);
    stg.col = reinterpret_cast<int*>(epi_smem + grp * TC_EPI_GROUP_SMEM + TC_STAGE_SLOTS * TC_EPI_THREADS * 4);
    stg.e = ((warp - 2) & 3) * 32 + lane;
    int nst = 0;
    EpiPool pool;
    pool.val = reinterpret_cast<float*>(epi_smem + (warp - 2) * (TC_POOL * 8));
    pool.meta = reinterpret_cast<uint32_t*>(pool.val + TC_POOL);
    uint32_t pool_cnt = 0, pool_fill = 0;
    const int acc = grp;
    uint32_t acc_phase = 0;
    int slot = grp;
    uint32_t sphase = 0;
    const bool dyn = p.window == 0;
    int64_t t_static = (int64_t)blockIdx.x + (int64_t)grp * gridDim.x;
    static_assert(TC_TILE_SLOTS % 2 == 0, "ring slots alternate between the two epilogue groups");
    for (;;) {
      int64_t t;
      if (dyn) {
        mbar_wait(tid_full_bar(slot), sphase);
        t = tile_ring[slot];
        __syncwarp();  // every lane has its copy before the slot goes back to the producer
        if (lane == 0) mbar_arrive(tid_empty_bar(slot));
        slot += 2;
        if (slot >= TC_TILE_SLOTS) { slot -= TC_TILE_SLOTS; sphase ^= 1u; }
        if (t < 0) break;
      } else {
        t = t_static;
        if (t >= p.ntiles) break;
        t_static += 2 * (int64_t)gridDim.x;
      }
      const int m = (int)(t % mtiles);
      const int64_t n = t / mtiles;
      const int64_t q = (int64_t)m * TC_BM + lane_base + lane;
      const int64_t tile_row0 = perm_row(p, p.row0 + n * BN);
      int64_t cols_valid = p.nvalid - tile_row0;  // <= 0 for a padding tile
      if (cols_valid > p.nrows - n * BN) cols_valid = p.nrows - n * BN;  // slab ends inside the tile (safe slabs)
      if (cols_valid > BN) cols_valid = BN;
      const bool qvalid = q < p.nq;
      // compare raw accumulators against tau expressed in accumulator units
      const float tau_raw = qvalid ? p.tau[q] * fwd : __int_as_float(0x7f800000);
      [[maybe_unused]] long long tw = TCT_NOW();
      mbar_wait(tfull_bar(acc), acc_phase);
      if (threadIdx.x == 64) { TCT_ADD(4, tw); TCT_COUNT(9); }
      tw = TCT_NOW();
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(acc * BN);
      if (p.dense) epilogue_dense_tile_pool<BN>(p, taddr_row, reinterpret_cast<uint64_t*>(pool.val), lane, q - lane, inv, tile_row0, cols_valid, n * BN);
      else if (p.two_pass) epilogue_filter_tile_pool<BN>(p, taddr_row, pool, lane, q - lane, tau_raw, inv, tile_row0, cols_valid, pool_cnt, pool_fill);
      else epilogue_filter_tile<BN>(p, taddr_row, stg, nst, q, tau_raw, inv, tile_row0, cols_valid);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));  // TMEM buffer back to the MMA warp ...
      if (threadIdx.x == 64) TCT_ADD(5, tw);
      tw = TCT_NOW();
      if (!dyn && ((warp - 2) & 3) == 0 && lane == 0) throttle_tile_done(p.done);
      epilogue_flush(p, stg, nst, q, inv, tile_row0);  // ... before the atomic round trip
      epilogue_pool_flush(p, pool, lane, q - lane, pool_cnt, pool_fill, inv, tile_row0);
      if (threadIdx.x == 64) TCT_ADD(8, tw);
      acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---- CTA-pair variant (cta_group::2) ------------------------------------------------
// Two CTAs of a cluster (an SM pair) cooperate on a 256-query x 256-row tile: each CTA
// stages its own 128 query rows (hi+lo) and HALF of the corpus tile (128 rows, hi+lo),
// the leader CTA issues tcgen05.mma.cta_group::2 (M=256) which reads A and B halves
// from both CTAs' shared memory, and each CTA ends up with its 128 x 256 accumulator in
// its own TMEM.  Per CTA this halves the corpus-tile bytes moved L2->smem and read by
// the tensor core (64 KB instead of 96 KB per k-block), which leaves room for a third
// pipeline stage.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_cta(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                                 int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int STAGES, int PASSES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_score_pair_kernel(const __grid_constant__ CUtensorMap tmQhi, const __grid_constant__ CUtensorMap tmQlo,
                     const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                     const TcParams p) {
  constexpr int BN = 256;                        // corpus rows per pair tile (TMEM columns)
  constexpr bool SPLIT = PASSES == 3;
  constexpr int HALF_BYTES = 128 * TC_BK * 2;    // 16 KB: 128 rows x 64 fp16
  constexpr int OFF_QLO = HALF_BYTES, OFF_BHI = SPLIT ? 2 * HALF_BYTES : HALF_BYTES, OFF_BLO = 3 * HALF_BYTES;
  constexpr int STAGE_BYTES = (SPLIT ? 4 : 2) * HALF_BYTES;    // Qhi, [Qlo,] Bhi-half[, Blo-half]
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // used in the leader
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };           // one per CTA
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };       // one per CTA
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };  // used in the leader
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* epi_smem = smem_raw + (bar_base - smem_u32(smem_raw)) + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmQhi); tma_prefetch_desc(&tmQlo);
    tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers initialised and TMEM allocated before any cross-CTA signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int mtiles = p.mtiles;  // pair tiles of 256 queries
  const int64_t cid = blockIdx.x >> 1;
  const int64_t ncl = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_q = (p.flags & 1) ? kL2EvictLast : kL2EvictNormal;
      const uint64_t pol_b = (p.flags & 2) ? kL2EvictFirst : kL2EvictNormal;
      unsigned long long done_seen = 0;
      for (int64_t t = cid; t < p.ntiles; t += ncl) {
        const int m = (int)(t % mtiles);
        const int64_t n = t / mtiles;
        const int32_t qrow = m * 256 + (int32_t)rank * 128;
        const int32_t brow = (int32_t)perm_row(p, p.row0 + n * BN) + (int32_t)rank * 128;
        throttle_wait(p.done, t, p.window, done_seen);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          // the leader's barrier collects the bytes of both CTAs
          if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          tma_load_2d_pair(sbase, &tmQhi, full_bar(stage), kb * TC_BK, qrow, pol_q);
          tma_load_2d_pair(sbase + OFF_BHI, &tmBhi, full_bar(stage), kb * TC_BK, brow, pol_b);
          if (SPLIT) {
            tma_load_2d_pair(sbase + OFF_QLO, &tmQlo, full_bar(stage), kb * TC_BK, qrow, pol_q);
            tma_load_2d_pair(sbase + OFF_BLO, &tmBlo, full_bar(stage), kb * TC_BK, brow, pol_b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t t = cid; t < p.ntiles; t += ncl) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          const uint64_t qhi = make_smem_desc(sbase);
          const uint64_t qlo = make_smem_desc(sbase + OFF_QLO);
          const uint64_t bhi = make_smem_desc(sbase + OFF_BHI);
          const uint64_t blo = make_smem_desc(sbase + OFF_BLO);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);
            const uint32_t first = (kb | k) != 0 ? 1u : 0u;
            if (SPLIT) {
              tc_mma_f16_pair(d_tmem, qlo + koff, bhi + koff, IDESC, first);
              tc_mma_f16_pair(d_tmem, qhi + koff, blo + koff, IDESC, 1u);
              tc_mma_f16_pair(d_tmem, qhi + koff, bhi + koff, IDESC, 1u);
            } else {
              tc_mma_f16_pair(d_tmem, qhi + koff, bhi + koff, IDESC, first);
            }
          }
          tc_commit_pair(empty_bar(stage));  // frees this stage in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_pair(tfull_bar(acc));  // accumulators complete in BOTH CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own TMEM) =====================
    const int lane_base = (warp & 3) * 32;
    const float inv = p.q_inv_scale[0] * p.b_inv_scale;
    const float fwd = 1.0f / inv;
    EpiStage stg;
    stg.val = reinterpret_cast<float*>(epi_smem);
    stg.col = reinterpret_cast<int*>(epi_smem + TC_STAGE_SLOTS * TC_EPI_THREADS * 4);
    stg.e = (warp - 2) * 32 + lane;
    int nst = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = cid; t < p.ntiles; t += ncl) {
      const int m = (int)(t % mtiles);
      const int64_t n = t / mtiles;
      const int64_t q = (int64_t)m * 256 + (int64_t)rank * 128 + lane_base + lane;
      const int64_t tile_row0 = perm_row(p, p.row0 + n * BN);
      int64_t cols_valid = p.nvalid - tile_row0;  // <= 0 for a padding tile
      if (cols_valid > p.nrows - n * BN) cols_valid = p.nrows - n * BN;  // slab ends inside the tile (safe slabs)
      if (cols_valid > BN) cols_valid = BN;
      const bool qvalid = q < p.nq;
      const float tau_raw = qvalid ? p.tau[q] * fwd : __int_as_float(0x7f800000);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(acc * BN);
      if (p.dense) epilogue_dense_tile<BN>(p, taddr_row, q, qvalid, inv, tile_row0, cols_valid, n * BN);
      else if (p.two_pass) epilogue_filter_tile_two_pass<BN>(p, taddr_row, q, tau_raw, inv, tile_row0, cols_valid);
      else epilogue_filter_tile<BN>(p, taddr_row, stg, nst, q, tau_raw, inv, tile_row0, cols_valid);
      tc_fence_before();
      __syncwarp();
      // the leader's MMA thread waits for the epilogues of both CTAs (8 warps)
      if (lane == 0) mbar_arrive_cluster(mapa_cta(tempty_bar(acc), 0));
      if (threadIdx.x == 64 && leader && p.window > 0) throttle_tile_done(p.done);
      epilogue_flush(p, stg, nst, q, inv, tile_row0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA frees TMEM / exits while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---- small-batch variant: corpus rows on the M side, queries on the N side ----------
// For nq <= 64 the 128 x BN tile above wastes tensor work on padded query rows (three
// MMA passes over 128 rows for, say, 8 real queries) and under the power cap that waste
// keeps the pass above the HBM time.  Here the roles are swapped: A = a 128-row corpus
// tile (M = 128), B = the NQ (16/32/64) query rows (N = NQ), D = [128 corpus rows x NQ
// queries] in TMEM.  MMA work scales with the real batch, the stage is almost pure corpus
// bytes (32 KB + NQ*256 B) so 3-5 stages fit, and the kernel is bound by the HBM stream
// of the operand planes.  Epilogue: one corpus row per thread; survivors of one query
// (= one TMEM column) are found with a warp ballot and appended with ONE atomicAdd per
// warp and column.
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

template <int NQ, int STAGES, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_score_small_kernel(const __grid_constant__ CUtensorMap tmQhi, const __grid_constant__ CUtensorMap tmQlo,
                      const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                      const TcParams p) {
  constexpr bool SPLIT = PASSES == 3;
  constexpr int A_BYTES = 128 * TC_BK * 2;  // corpus tile, one plane
  constexpr int Q_BYTES = NQ * TC_BK * 2;   // query tile, one plane
  constexpr int OFF_BLO = A_BYTES, OFF_QHI = SPLIT ? 2 * A_BYTES : A_BYTES, OFF_QLO = OFF_QHI + Q_BYTES;
  constexpr int STAGE_BYTES = SPLIT ? 2 * A_BYTES + 2 * Q_BYTES : A_BYTES + Q_BYTES;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NQ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t TMEM_COLS = (2 * NQ < 32) ? 32u : (uint32_t)(2 * NQ);
  constexpr int CW = (NQ < 32) ? NQ : 32;  // columns per TMEM load
  static_assert(NQ == 16 || NQ == 32 || NQ == 64, "NQ must be 16/32/64");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  float* tau_s = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 8u * (2 * STAGES + 6));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmQhi); tma_prefetch_desc(&tmQlo);
    tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    // thresholds in accumulator units, one per query column (padding columns never pass)
    const float inv0 = p.q_inv_scale[0] * p.b_inv_scale;
    for (int j = threadIdx.x; j < NQ; j += blockDim.x)
      tau_s[j] = (j < p.nq) ? p.tau[j] * (1.0f / inv0) : __int_as_float(0x7f800000);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int32_t brow = (int32_t)perm_row(p, p.row0 + t * 128);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d_hint(sbase, &tmBhi, full_bar(stage), kb * TC_BK, brow, kL2EvictFirst);
          tma_load_2d_hint(sbase + OFF_QHI, &tmQhi, full_bar(stage), kb * TC_BK, 0, kL2EvictLast);
          if (SPLIT) {
            tma_load_2d_hint(sbase + OFF_BLO, &tmBlo, full_bar(stage), kb * TC_BK, brow, kL2EvictFirst);
            tma_load_2d_hint(sbase + OFF_QLO, &tmQlo, full_bar(stage), kb * TC_BK, 0, kL2EvictLast);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NQ);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          const uint64_t bhi = make_smem_desc(sbase);
          const uint64_t blo = make_smem_desc(sbase + OFF_BLO);
          const uint64_t qhi = make_smem_desc(sbase + OFF_QHI);
          const uint64_t qlo = make_smem_desc(sbase + OFF_QLO);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);
            const uint32_t first = (kb | k) != 0 ? 1u : 0u;
            if (SPLIT) {
              tc_mma_f16(d_tmem, bhi + koff, qlo + koff, IDESC, first);
              tc_mma_f16(d_tmem, blo + koff, qhi + koff, IDESC, 1u);
              tc_mma_f16(d_tmem, bhi + koff, qhi + koff, IDESC, 1u);
            } else {
              tc_mma_f16(d_tmem, bhi + koff, qhi + koff, IDESC, first);
            }
          }
          tc_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: one corpus row per thread =====================
    // Survivors of one query (= one TMEM column) are counted with warp ballots; the four epilogue warps pool
    // their counts in shared memory, ONE thread per query reserves room in the query's buffer with one global
    // atomicAdd for the whole 128-row tile, and the keys are then stored at deterministic offsets.  (The first
    // version issued a dependent atomicAdd per warp and column with a survivor: fine at the ~1e-4 pass rates of
    // a last slab, but at the 1 % of a k = 1000 mid slab the serialised round trips took three times the tile's
    // HBM time.)  The TMEM buffer goes back to the MMA warp as soon as the columns are in registers.
    const int lane_base = (warp & 3) * 32;
    const int ew = warp & 3;  // epilogue warp number = TMEM lane quarter
    const float inv = p.q_inv_scale[0] * p.b_inv_scale;
    // two sets of {wcnt[4][32], gbase[32], any}, alternating per chunk: a warp that is already counting the next
    // chunk (it cannot be further ahead: the next barrier needs everybody) must not disturb one still storing this one
    uint32_t* pool = reinterpret_cast<uint32_t*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 512);
    uint32_t set = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      const int64_t grow = perm_row(p, p.row0 + t * 128) + lane_base + lane;
      const bool rvalid = grow < p.nvalid && t * 128 + lane_base + lane < p.nrows;
      const int64_t slot = t * 128 + lane_base + lane;  // dense slab: position in the query's buffer
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(acc * NQ);
#pragma unroll 1
      for (int c = 0; c < NQ / CW; ++c) {
        uint32_t v[32];
        __syncwarp();
        if (CW == 16) tmem_ld_x16(taddr_row + (uint32_t)(c * CW), v);
        else tmem_ld_x32(taddr_row + (uint32_t)(c * CW), v);
        tmem_ld_wait();
        if (c == NQ / CW - 1) {  // last chunk in registers: the accumulator buffer is free
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (p.dense) {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int qj = c * CW + j;
            if (qj < p.nq) {  // warp-uniform
              const float sc = __uint_as_float(v[j]) * inv;
              p.cand[(int64_t)qj * p.cap + slot] = (rvalid && sc > CMX_NEG_PAD) ? make_key(sc, (uint32_t)grow) : 0ull;
            }
          }
          continue;
        }
        uint32_t* wcnt = pool + set * 192;           // [4][32]
        uint32_t* gbase = wcnt + 128;                // [32]
        volatile uint32_t* tile_any = gbase + 32;
        set ^= 1u;
        // A. per-column survivor ballots of this warp; lane j keeps the count of column j
        uint32_t mine = 0, pass_bits = 0;
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          const int qj = c * CW + j;
          const bool pass = qj < p.nq && rvalid && __uint_as_float(v[j]) > tau_s[qj];
          const uint32_t b = __ballot_sync(0xffffffffu, pass);
          if (lane == j) mine = (uint32_t)__popc(b);
          pass_bits |= pass ? (1u << j) : 0u;
        }
        if (lane < CW) wcnt[ew * 32 + lane] = mine;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // B. one thread per query reserves room for the whole tile
        if (ew == 0) {
          uint32_t total = 0;
          if (lane < CW) total = wcnt[lane] + wcnt[32 + lane] + wcnt[64 + lane] + wcnt[96 + lane];
          uint32_t base = 0;
          if (total) base = atomicAdd(&p.cnt[c * CW + lane], total);
          if (lane < CW) gbase[lane] = base;
          const uint32_t any = __ballot_sync(0xffffffffu, total != 0);
          if (lane == 0) *tile_any = any;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // C. store the keys: offset = tile base + the lower warps' counts + the lower lanes of this warp
        const uint32_t any = *tile_any;
        if (any) {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            if ((any >> j) & 1u) {  // CTA-uniform
              const bool pass = (pass_bits >> j) & 1u;
              const uint32_t b = __ballot_sync(0xffffffffu, pass);
              if (pass) {
                uint32_t pos = gbase[j] + __popc(b & ((1u << lane) - 1u));
                for (int w = 0; w < ew; ++w) pos += wcnt[w * 32 + j];
                const int qj = c * CW + j;
                if (pos < (uint32_t)p.cap) p.cand[(int64_t)qj * p.cap + pos] = make_key(__uint_as_float(v[j]) * inv, (uint32_t)grow);
              }
            }
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

static int make_plane_map(CUtensorMap* tm, const __half* base, int64_t rows, int d_pad, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CMX_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)d_pad * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return CMX_ERR_CUDA; }
  return CMX_OK;
}

int tensor_path_available() { return get_encode() != nullptr; }

static int g_tc_bn = 256;  // tile width in corpus rows: 256 (2 stages) or 128 (3 stages)
static int g_tc_flags = 0;
static int g_tc_window = 3;  // progress throttle: a CTA may run this many round-robin iterations ahead (0 = off)
void set_tensor_window(int w) { g_tc_window = w < 0 ? 3 : w; }
// CTA-pair kernel (cta_group::2) for nq > 128: each CTA stages half of the corpus tile, so a third less L2 -> shared
// memory traffic per FLOP.  Measured alternating in one process, randomised order (profiles/r02_variants_ab.md):
//   8.8 M rows  single 106.3 ms per step, pair 104.8      4.4 M rows  56.9 / 56.2
//   2.2 M rows  31.6 / 31.4                                1.1 M rows  18.6 / 19.2 (pair SLOWER)
// Long launches run at the 1000 W power cap, where fewer bytes moved per FLOP buy clock; short launches are not
// power-limited and the single-CTA kernel's finer tile granularity wins.  -1 = automatic: pair for slabs of at least
// kPairMinRows rows, 0 = never, 1 = always.
constexpr int64_t kPairMinRows = 3 << 20;
static int g_tc_pair = -1;
void set_tensor_pair(int on) { g_tc_pair = (on < 0) ? -1 : (on ? 1 : 0); }
void set_tensor_tile(int bn) { g_tc_bn = (bn == 128) ? 128 : 256; }
void set_tensor_flags(int f) { g_tc_flags = f; }
#ifdef CMX_TC_TIMERS
extern "C" __attribute__((visibility("default"))) int cmx_debug_tc_timers(unsigned long long* out16, int reset) {
  if (out16 && cudaMemcpyFromSymbol(out16, g_tc_timers, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
  static const unsigned long long zeros[16] = {};
  if (reset && cudaMemcpyToSymbol(g_tc_timers, zeros, sizeof(zeros)) != cudaSuccess) return 1;
  return 0;
}
#endif

template <int BN, int STAGES, int PASSES>
static int launch_tc(const CUtensorMap& tq_hi, const CUtensorMap& tq_lo, const CUtensorMap& tb_hi,
                     const CUtensorMap& tb_lo, const TcParams& p, cudaStream_t st, int sm_count) {
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TC_A_BYTES + BN * TC_BK * 2);
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + TC_EPI_SMEM;
  static_assert((size_t)STAGES * STAGE_BYTES + 1024 + 256 + TC_EPI_SMEM <= 227 * 1024, "stage ring exceeds shared memory");
  CMX_CUDA(cudaFuncSetAttribute(tc_score_kernel<BN, STAGES, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = p.ntiles < sm_count ? p.ntiles : sm_count;
  if (grid < 1) return CMX_OK;
  tc_score_kernel<BN, STAGES, PASSES><<<(unsigned)grid, TC_THREADS1, smem, st>>>(tq_hi, tq_lo, tb_hi, tb_lo, p);
  CMX_LAUNCHED();
  return CMX_OK;
}

template <int STAGES, int PASSES>
static int launch_tc_pair(const CUtensorMap& tq_hi, const CUtensorMap& tq_lo, const CUtensorMap& tb_hi,
                          const CUtensorMap& tb_lo, const TcParams& p, cudaStream_t st, int sm_count) {
  constexpr int STAGE_BYTES = (PASSES == 3 ? 4 : 2) * (128 * TC_BK * 2);
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + TC_EPI_SMEM;
  static_assert((size_t)STAGES * STAGE_BYTES + 1024 + 256 + TC_EPI_SMEM <= 227 * 1024, "stage ring exceeds shared memory");
  CMX_CUDA(cudaFuncSetAttribute(tc_score_pair_kernel<STAGES, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t clusters = sm_count / 2;
  if (p.ntiles < clusters) clusters = p.ntiles;
  if (clusters < 1) return CMX_OK;
  tc_score_pair_kernel<STAGES, PASSES><<<(unsigned)(2 * clusters), TC_THREADS, smem, st>>>(tq_hi, tq_lo, tb_hi, tb_lo, p);
  CMX_LAUNCHED();
  return CMX_OK;
}

template <int NQ, int STAGES, int PASSES>
static int launch_tc_small(const __half* Bhi, const __half* Blo, int64_t plane_rows, const __half* Qhi, const __half* Qlo,
                           int64_t nq_pad, int d_pad, TcParams p, cudaStream_t st, int sm_count) {
  CUtensorMap tq_hi, tq_lo, tb_hi, tb_lo;
  CMX_TRY(make_plane_map(&tq_hi, Qhi, nq_pad, d_pad, NQ));
  CMX_TRY(make_plane_map(&tb_hi, Bhi, plane_rows, d_pad, 128));
  if (PASSES == 3) {
    CMX_TRY(make_plane_map(&tq_lo, Qlo, nq_pad, d_pad, NQ));
    CMX_TRY(make_plane_map(&tb_lo, Blo, plane_rows, d_pad, 128));
  } else {
    tq_lo = tq_hi;
    tb_lo = tb_hi;
  }
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * ((128 * TC_BK * 2) + (NQ * TC_BK * 2));
  // 1024 alignment slack | 512 barriers + thresholds | 2 x 776 B survivor counts of the epilogue
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 512 + 2048;
  static_assert((size_t)STAGES * STAGE_BYTES + 1024 + 512 + 2048 <= 227 * 1024, "stage ring exceeds shared memory");
  CMX_CUDA(cudaFuncSetAttribute(tc_score_small_kernel<NQ, STAGES, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.mtiles = 1;
  p.ntiles = (p.nrows + 127) / 128;
  int64_t grid = p.ntiles < sm_count ? p.ntiles : sm_count;
  if (grid < 1) return CMX_OK;
  tc_score_small_kernel<NQ, STAGES, PASSES><<<(unsigned)grid, TC_THREADS, smem, st>>>(tq_hi, tq_lo, tb_hi, tb_lo, p);
  CMX_LAUNCHED();
  return CMX_OK;
}

static int g_tc_small = 1;  // nq <= 64: corpus-as-M kernel (measured: 5.0-5.8 ms per 36 GB sweep vs 6.2-6.4 ms padded)
void set_tensor_small(int on) { g_tc_small = on ? 1 : 0; }

int launch_tensor_score(const __half* Bhi, const __half* Blo, int64_t plane_rows, int64_t row0,
                        int64_t nrows, int d_pad, const __half* Qhi, const __half* Qlo,
                        int64_t nq, int64_t nq_pad, const float* q_inv_scale_dev, float b_inv_scale,
                        const SearchWs& ws, int dense, uint64_t perm, int passes, double expected_pass_rate,
                        unsigned long long* progress, cudaStream_t st, int sm_count) {
  if (nrows <= 0 || nq <= 0) return CMX_OK;
  CMX_CHECK(passes == 1 || passes == 3, "tensor path: passes must be 1 or 3");
  CMX_CHECK(d_pad % TC_BK == 0, "tensor path: padded dim must be a multiple of %d", TC_BK);
  CMX_CHECK(plane_rows < (int64_t)0x7fffffff, "tensor path: more than 2^31 rows per shard");
  const bool split = passes == 3;
  TcParams p;
  // a slab is a run of whole 256-row blocks, or (worst-case-safe slabs of a tiny candidate buffer) a piece of one block
  CMX_CHECK(((row0 & 255) == 0 && (nrows & 255) == 0) || ((row0 & 255) + nrows <= 256),
            "tensor path: slab [%lld, +%lld) straddles 256-row blocks", (long long)row0, (long long)nrows);
  p.row0 = row0;
  p.nrows = nrows;
  p.nblk = (plane_rows + 255) / 256;
  p.perm = perm % (uint64_t)p.nblk;
  if (p.perm == 0) p.perm = 1;
  p.nvalid = plane_rows;
  p.kblocks = d_pad / TC_BK;
  p.nq = nq;
  p.q_inv_scale = q_inv_scale_dev;
  p.b_inv_scale = b_inv_scale;
  p.tau = ws.tau;
  p.cnt = ws.cnt;
  p.cand = ws.cand;
  p.cap = ws.cap;
  p.dense = dense;
  p.flags = g_tc_flags;
  p.mtiles = 1;
  p.ntiles = 0;
  p.done = nullptr;
  p.window = 0;
  // more than ~6 expected survivors per row and tile (slab right after the dense one): the staging
  // area would be flushed several times per tile with the TMEM buffer held
  p.two_pass = (!dense && expected_pass_rate * 256.0 > 6.0) ? 1 : 0;
  if (g_tc_flags & 16) p.two_pass = 1;
  if (g_tc_flags & 32) p.two_pass = 0;
  // corpus-as-M kernel: up to 64 queries with three passes, up to 32 with one pass (measured: at
  // 64 queries and one pass the padded 128-row kernel is faster, 2.7 vs 3.9 ms per sweep)
  if (g_tc_small && nq <= (split ? 64 : 32)) {
    if (split) {
      if (nq <= 16) return launch_tc_small<16, 5, 3>(Bhi, Blo, plane_rows, Qhi, Qlo, nq_pad, d_pad, p, st, sm_count);
      if (nq <= 32) return launch_tc_small<32, 4, 3>(Bhi, Blo, plane_rows, Qhi, Qlo, nq_pad, d_pad, p, st, sm_count);
      return launch_tc_small<64, 4, 3>(Bhi, Blo, plane_rows, Qhi, Qlo, nq_pad, d_pad, p, st, sm_count);
    }
    if (nq <= 16) return launch_tc_small<16, 8, 1>(Bhi, Blo, plane_rows, Qhi, Qlo, nq_pad, d_pad, p, st, sm_count);
    return launch_tc_small<32, 8, 1>(Bhi, Blo, plane_rows, Qhi, Qlo, nq_pad, d_pad, p, st, sm_count);
  }
  const bool pair = nq > 128 && (g_tc_pair > 0 || (g_tc_pair < 0 && nrows >= kPairMinRows));
  const int bn = pair ? 128 : g_tc_bn;  // pair: each CTA loads a 128-row half of the 256-row tile
  CUtensorMap tq_hi, tq_lo, tb_hi, tb_lo;
  CMX_TRY(make_plane_map(&tq_hi, Qhi, nq_pad, d_pad, TC_BM));
  CMX_TRY(make_plane_map(&tb_hi, Bhi, plane_rows, d_pad, bn));
  if (split) {
    CMX_TRY(make_plane_map(&tq_lo, Qlo, nq_pad, d_pad, TC_BM));
    CMX_TRY(make_plane_map(&tb_lo, Blo, plane_rows, d_pad, bn));
  } else {
    tq_lo = tq_hi;
    tb_lo = tb_hi;
  }
  p.mtiles = pair ? (int)((nq + 255) / 256) : (int)((nq + TC_BM - 1) / TC_BM);
  p.ntiles = pair ? (int64_t)p.mtiles * ((nrows + 255) / 256) : (int64_t)p.mtiles * ((nrows + bn - 1) / bn);
  // single-CTA kernel: `progress` is the tile counter of its dynamic scheduler; pair kernel: the progress
  // counter of its throttle (static round-robin tiles)
  CMX_CHECK(progress != nullptr, "tensor path: no tile counter");
  CMX_CUDA(cudaMemsetAsync(progress, 0, sizeof(unsigned long long), st));
  p.done = progress;
  if (pair && g_tc_window > 0) p.window = (long long)g_tc_window * (sm_count / 2);  // iterations of slack
  if (pair && g_tc_window <= 0) p.done = nullptr;
  // single-CTA kernel: static round-robin tiles + progress throttle by default; flag 512 selects the dynamic
  // scheduler (window = 0).  Measured alternating in one process (profiles/r02_scheduler_ab.md): static is 3 %
  // faster when nothing shares the SMs; dynamic only pays off beside prescoring, which is off by default.
  if (!pair && !(g_tc_flags & 512)) p.window = (long long)(g_tc_window > 0 ? g_tc_window : 3) * sm_count;
  if (pair) {
    if (split) return launch_tc_pair<3, 3>(tq_hi, tq_lo, tb_hi, tb_lo, p, st, sm_count);
    return launch_tc_pair<6, 1>(tq_hi, tq_lo, tb_hi, tb_lo, p, st, sm_count);
  }
  if (split) {
    if (bn == 256) return launch_tc<256, 2, 3>(tq_hi, tq_lo, tb_hi, tb_lo, p, st, sm_count);
    return launch_tc<128, 3, 3>(tq_hi, tq_lo, tb_hi, tb_lo, p, st, sm_count);
  }
  if (bn == 256) return launch_tc<256, 4, 1>(tq_hi, tq_lo, tb_hi, tb_lo, p, st, sm_count);
  return launch_tc<128, 6, 1>(tq_hi, tq_lo, tb_hi, tb_lo, p, st, sm_count);
}

}  // namespace cmx

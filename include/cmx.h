/*
 * cmx.h -- C ABI of libcmx.so: B200 (sm_100a) vector-mix + flat inner-product
 * top-k search.  This is the drop-in boundary for the search path of
 * cmHuang777/codemix-dense-retrieval: every entry point below replaces one call
 * the reference makes into faiss / torch (reference file:line cited per entry).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns 0 on success and a non-zero code on error, with a
 *     thread-local message available from cmx_last_error(); nothing aborts the
 *     process (the reference lets faiss raise RuntimeError: the Python shim maps
 *     non-zero -> RuntimeError(cmx_last_error()));
 *   - `*_on_device` = 0: the pointer is host memory (pageable or pinned),
 *                     1: the pointer is device memory on the index's device;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default
 *     stream).  Calls are synchronous: results are valid on return;
 *   - an index owns its device storage; add() copies the caller's rows; no
 *     caller pointer is retained after a call returns;
 *   - one index lives on one device ("shard").  Multi-GPU = one shard per
 *     device/rank plus cmx_merge_topk.
 *
 * There is NO CPU fallback anywhere behind this header: if no CUDA device is
 * usable every compute entry point fails with an error.
 */
#ifndef CMX_H_
#define CMX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CMX_API __attribute__((visibility("default")))
#else
#define CMX_API
#endif

typedef struct cmx_index cmx_index;

/* error codes */
#define CMX_OK 0
#define CMX_ERR_INVALID 1   /* bad argument (shape, k, null pointer ...)   */
#define CMX_ERR_CUDA 2      /* CUDA runtime / driver error                  */
#define CMX_ERR_NOMEM 3     /* device allocation failed                     */
#define CMX_ERR_INTERNAL 4

/* scoring path selector for search */
#define CMX_PATH_AUTO 0     /* tensor; stream for nq <= 4 in split precision */
#define CMX_PATH_STREAM 1   /* CUDA-core fp32 streaming scorer (HBM-bound)  */
#define CMX_PATH_TENSOR 2   /* tcgen05 fp16-split (hi/lo) tensor-core scorer */

/* arithmetic of the tensor path (the stream path is always exact fp32) */
#define CMX_PRECISION_SPLIT 0   /* three fp16 MMA passes (hi*hi + hi*lo + lo*hi), fp32 accumulate:
                                   fp32-faithful scores straight out of the tensor core            */
#define CMX_PRECISION_RESCORE 1 /* default: ONE fp16 MMA pass selects a provably sufficient superset
                                   (everything within 2*eps(q) of the k-th best approximate score,
                                   eps from the fp16 rounding bound), then the survivors are scored
                                   EXACTLY in fp32 from the row store and the exact top-k is taken  */

/* FAISS-GPU's own cap on k is 2048; kept here. */
#define CMX_MAX_K 2048

CMX_API const char* cmx_last_error(void);
CMX_API int cmx_version(void);
/* number of visible CUDA devices (0 and an error string when none). */
CMX_API int cmx_device_count(int* out);
/* cumulative number of kernels launched by this library in this process. */
CMX_API uint64_t cmx_launch_count(void);

/* ---- index lifecycle -------------------------------------------------------
 * replaces: faiss.IndexFlatIP(dim) + faiss.index_cpu_to_gpu(res, gpu_id, index)
 *   onepass_dense_mix_run_custom_lang.py:604,658-664,726-730
 *   onepass_bilingual_mix_hub_custom_lang.py:558,931-936
 *   encode_multilingual_corpus.py:367-373
 */
CMX_API int cmx_index_create(int d, int device, cmx_index** out);
CMX_API int cmx_index_free(cmx_index* ix);
/* pre-size storage for n rows (avoids re-allocation while adding). */
CMX_API int cmx_index_reserve(cmx_index* ix, int64_t n);
/* replaces: index.add(x) / index.add_with_ids(x, ids) (ids stay host-side in the
 * Python IndexIDMap)  onepass_dense_mix_run_custom_lang.py:719,
 * onepass_bilingual_mix_hub_custom_lang.py:646,672, encode_multilingual_corpus.py:440 */
CMX_API int cmx_index_add(cmx_index* ix, const float* x, int64_t n, int x_on_device);
/* replaces: faiss.read_index(path) + index_cpu_to_gpu (onepass_dense_mix_run_custom_lang.py:250,658-664) for the
 * vector block of an index.faiss: appends the n rows stored row-major as little-endian float32 at byte `offset`
 * of `path`.  Reader threads (nthreads <= 0: automatic) fill page-locked staging buffers while the previous chunk
 * crosses PCIe; nothing but the staging buffers lives in host memory.  seconds_out (may be NULL): [0] = time spent
 * inside pread, [1] = wall time of the call. */
CMX_API int cmx_index_add_from_file(cmx_index* ix, const char* path, int64_t offset, int64_t n, int nthreads,
                                    double* seconds_out);
/* host-side helper of faiss.read_index (onepass_dense_mix_run_custom_lang.py:250): `bytes` of `path` at `offset` into
 * host memory `dst` with nthreads concurrent preads (<= 0: automatic) -- 36 GB of vectors in seconds, not in the
 * ~9 s of one sequential read.  No GPU involved. */
CMX_API int cmx_read_file(const char* path, int64_t offset, int64_t bytes, void* dst, int nthreads);
/* replaces: np.vstack([base_index.reconstruct(i) for i in batch]) + add_with_ids of the bilingual combined-index
 * build (onepass_bilingual_mix_hub_custom_lang.py:644-646): appends rows[0..n) (host array of row numbers) of
 * `src` -- another index on the same device -- without the rows leaving the GPU. */
CMX_API int cmx_index_add_gather(cmx_index* ix, const cmx_index* src, const int64_t* rows, int64_t n);
CMX_API int cmx_index_reset(cmx_index* ix);
CMX_API int cmx_index_ntotal(const cmx_index* ix, int64_t* out);
/* device bytes held by the index: out4 = {fp32 row store, fp16 operand plane(s), search workspace, ntotal*d*4 = what
 * a FAISS flat index of the same rows holds}.  Default precision: store + one plane = 1.5x a FAISS flat index. */
CMX_API int cmx_index_memory(const cmx_index* ix, int64_t* out4);
CMX_API int cmx_index_dim(const cmx_index* ix, int* out);
CMX_API int cmx_index_device(const cmx_index* ix, int* out);
/* replaces: base_index.reconstruct(i[, out])  onepass_dense_mix_run_custom_lang.py:268-269,
 * onepass_bilingual_mix_hub_custom_lang.py:644.  Copies rows [i0, i0+n). */
CMX_API int cmx_index_reconstruct(const cmx_index* ix, int64_t i0, int64_t n, float* out, int out_on_device);
/* device pointer of the fp32 row store (rows [0, ntotal), leading dim = d);
 * valid until the next add/reserve/reset/free.  For zero-copy device concat
 * (bilingual combined-index build, onepass_bilingual_mix_hub_custom_lang.py:606-702). */
CMX_API int cmx_index_data(const cmx_index* ix, const float** out);

/* ---- search ----------------------------------------------------------------
 * replaces: D, I = index.search(q_chunk, k)
 *   onepass_dense_mix_run_custom_lang.py:878 (k=100)
 *   onepass_bilingual_mix_hub_custom_lang.py:950 (k=--topk)
 *   onepass_dense_run.py:427,460 (nq=1)
 * q [nq,d] fp32 row-major; D [nq,k] fp32 sorted descending; I [nq,k] int64 =
 * id_base + row number, ties by ascending row; tail padded with I=-1,
 * D=-FLT_MAX when k > ntotal (FAISS semantics).  1 <= k <= CMX_MAX_K. */
CMX_API int cmx_index_search(cmx_index* ix, const float* q, int64_t nq, int k, float* D, int64_t* I,
                     int io_on_device, int64_t id_base, int path, void* stream);

/* ---- vector-mix prologue ---------------------------------------------------
 * replaces: safe_mix(...) called per query per alpha
 *   onepass_dense_mix_run_custom_lang.py:342-377,853-867
 *   onepass_bilingual_mix_hub_custom_lang.py:390-424,901-919
 * out [nA,nq,d]: |alpha|<=1e-8 -> P row; |alpha-1|<=1e-8 -> S row; else
 * normalize(fl32(1-alpha)*P + fl32(alpha)*S) with each mul/add rounded (no FMA),
 * x / max(||x||,1e-12); non-finite result -> S row if |alpha|>0.5 else P row.
 * flags [nA,nq] (may be NULL): 0 ok, 1 fell back to P, 2 fell back to S. */
CMX_API int cmx_mix_normalize(const float* P, const float* S, int64_t nq, int d, const double* alphas,
                      int nA, float* out, uint8_t* flags, int io_on_device, int device,
                      void* stream);

/* ---- fused prologue + search over a batch of alphas -------------------------
 * replaces the body of the alpha loop: onepass_dense_mix_run_custom_lang.py:846-886,
 * onepass_bilingual_mix_hub_custom_lang.py:901-919 + 942-952.
 * D [nA,nq,k], I [nA,nq,k], flags [nA,nq] (may be NULL). */
CMX_API int cmx_search_mixed(cmx_index* ix, const float* P, const float* S, int64_t nq,
                     const double* alphas, int nA, int k, float* D, int64_t* I, uint8_t* flags,
                     int io_on_device, int64_t id_base, int path, void* stream);

/* ---- two-phase search for row-sharded indexes (rescore precision, device buffers) -------------
 * New capability (BASELINE north_star (4)); the reference never shards (it calls
 * index_cpu_to_gpu(res, ONE gpu, index): onepass_dense_mix_run_custom_lang.py:661-663).
 * A shard that rescored its own k best would do G times the necessary exact work, and a step that
 * stopped for host decisions would leave G GPUs idle.  These entry points are building blocks that
 * only ENQUEUE work on `stream` (no host synchronisation; unlike the rest of this header results are
 * valid once the stream has been synchronised).  The caller provides the cross-shard barriers
 * (symmetric-memory barriers between processes, events between the streams of one process).
 *  0. cmx_search_prepare: the fused mix + normalise prologue for nA alphas into the index's own query
 *     buffer; *q_out [nA*nq, d] (device) is valid until the next prepare / search_mixed on the index.
 *     cmx_index_export_bounds: this shard's {max row norm, max fp16-residual norm} -> out2 (device,
 *     e.g. this shard's slot of a symmetric buffer).          [barrier: bounds of all shards visible]
 *  1. cmx_search_begin: approximate pass over the shard for queries q [nq <= 8192, d] (chunk larger
 *     batches); the filter margin uses the maxima over bounds_parts[0..nparts) (peer memory allowed);
 *     writes the shard's k best APPROXIMATE scores per query to scores_out [nq, k] (order arbitrary,
 *     short lists padded with lowest-float) and its status word to flag_out (0 = ok; buffer overflow,
 *     failed speculation or "this shard cannot run the one-pass arithmetic" otherwise).  est_scale =
 *     1 / number of shards: how deep exact scores are computed ahead of time beside the scoring
 *     kernel.  An empty shard is fine.               [barrier: scores + status of all shards visible]
 *  2. cmx_union_kth: the GLOBAL k-th best approximate score of queries [q0, q1) = k-th largest of the
 *     union of all shards' lists, read in place (score_parts[g] may be peer memory), written to every
 *     kth_outs[o][q] (local or peer); ORs flag_parts[0..nflags) into *flag_any (device).
 *                                                          [barrier: kth of all queries visible]
 *  3. cmx_search_end: exact fp32 rescoring of the rows with approx >= max_p kth_parts[p][q] -
 *     2*eps(q) only; D, I [nq,k] then hold the shard's exact hits that can still belong to the global
 *     top-k (fewer than k: padded with -1), ready for cmx_merge_topk(_peers).
 *  After the step the caller reads *flag_any: non-zero on ANY shard is non-zero on ALL of them (every
 *  shard ORs the same words), so all shards redo the step with cmx_search_mixed + merge together. */
CMX_API int cmx_search_prepare(cmx_index* ix, const float* P, const float* S, int64_t nq,
                               const double* alphas, int nA, const float** q_out, void* stream);
CMX_API int cmx_index_export_bounds(cmx_index* ix, float* out2_dev, void* stream);
CMX_API int cmx_search_begin(cmx_index* ix, const float* q, int64_t nq, int k, int64_t id_base,
                             const float* const* bounds_parts, int nparts, float est_scale,
                             float* scores_out, uint32_t* flag_out, void* stream);
CMX_API int cmx_union_kth(const float* const* score_parts, int nparts, int64_t nq, int k, int64_t q0,
                          int64_t q1, float* const* kth_outs, int nouts,
                          const uint32_t* const* flag_parts, int nflags, uint32_t* flag_any,
                          int device, void* stream);
CMX_API int cmx_search_end(cmx_index* ix, const float* const* kth_parts, int nparts, float* D,
                           int64_t* I, void* stream);
/* src [bytes] (device) -> every dsts[i] (device, usually peer memory): one read, ndst 128-bit stores
 * per element over NVLink.  Used to replicate each rank's slice of the uploaded query vectors, so the
 * 57 MB of P,S cross PCIe once per NODE instead of once per GPU.  Asynchronous. */
CMX_API int cmx_peer_broadcast(const void* src, void* const* dsts, int ndst, int64_t bytes, int device,
                               void* stream);
/* Page-lock and device-map an existing host range (e.g. a POSIX shared-memory segment every rank
 * maps): kernels may then write results straight into it (cmx_merge_topk_peers outputs,
 * cmx_index_search / cmx_search_mixed host outputs).  *dev_ptr = the device-side alias. */
CMX_API int cmx_host_register(void* p, int64_t bytes, void** dev_ptr);
CMX_API int cmx_host_unregister(void* p);
/* one process driving several GPUs: let `device` map the memory of `peer`. */
CMX_API int cmx_enable_peer_access(int device, int peer);

/* ---- k-way merge of per-shard results (multi-GPU) ---------------------------
 * new capability (BASELINE north_star (4)); the reference never shards.
 * D_parts [nparts,nq,k], I_parts [nparts,nq,k] -> D [nq,k], I [nq,k]; order:
 * score desc, then part number, then position inside the part. */
CMX_API int cmx_merge_topk(const float* D_parts, const int64_t* I_parts, int nparts, int64_t nq, int k,
                   float* D, int64_t* I, int io_on_device, int device, void* stream);

/* fused exchange + merge over peer memory (NVLink P2P / symmetric memory): D_parts[g] /
 * I_parts[g] are DEVICE pointers to part g's [nq,k] lists, each possibly resident on another
 * GPU and mapped into this process; the kernel reads them in place, merges queries
 * [q0, q1) and stores the merged rows into each of the nouts output buffers D_outs[o] /
 * I_outs[o] ([nq,k]: local, peer, or the device alias of a registered host buffer).  The caller provides the cross-rank barriers before
 * (parts complete) and after (outputs complete).  Asynchronous on `stream` unlike the
 * other entry points: it returns once the kernel is enqueued.  nparts, nouts <= 16. */
CMX_API int cmx_merge_topk_peers(const float* const* D_parts, const int64_t* const* I_parts, int nparts,
                                 int64_t nq, int k, int64_t q0, int64_t q1, float* const* D_outs,
                                 int64_t* const* I_outs, int nouts, int device, void* stream);

/* ---- run-file text (host side, multi-threaded) -------------------------------------
 * replaces the per-line f-string loops: onepass_dense_mix_run_custom_lang.py:879-888 (mono),
 * onepass_bilingual_mix_hub_custom_lang.py:950-958 (raw) and :165-181 (collapse_run_max).
 * Strings travel as one UTF-8 buffer plus n+1 offsets.  D, I are HOST arrays [nq,k].
 * The returned buffers are malloc'ed; release them with cmx_free_text.
 * mono: lines "qid\tQ0\tdoc\trank\tscore(.4f)\ttag" joined by '\n' (no trailing newline);
 *   doc = docs[pos] where pos = position of the id in the sorted doc_keys (or the id itself
 *   when doc_keys is NULL), else the decimal id (id_lookup.get(int(doc), str(doc))).
 * bilingual: raw lines "qid Q0 did rank score(.6f) tag\n" for ids inside [0, ndocs) (others
 *   skipped, rank kept) and the collapsed run: per query group by base_code[id], max of the
 *   6-decimal rounded scores, stable sort descending, "qid Q0 base rank score bilingual-mix\n". */
CMX_API int cmx_trec_mono(const float* D, const int64_t* I, int64_t nq, int k, const char* qids,
                          const int64_t* qid_off, const char* docs, const int64_t* doc_off,
                          const int64_t* doc_keys, int64_t ndocs, const char* tag, int nthreads,
                          char** out, int64_t* out_len);
CMX_API int cmx_trec_bilingual(const float* D, const int64_t* I, int64_t nq, int k, const char* qids,
                               const int64_t* qid_off, const char* docs, const int64_t* doc_off,
                               int64_t ndocs, const int32_t* base_code, const char* bases,
                               const int64_t* base_off, int64_t nbases, const char* tag, int nthreads,
                               char** raw_out, int64_t* raw_len, char** col_out, int64_t* col_len);
CMX_API void cmx_free_text(char* p);
/* The same text written straight to files (created / truncated): every formatting thread
 * pwrites its part at its final offset, so the run file (300 MB at 6980 x 1000 hits) is never
 * assembled in memory.  *_len receive the file sizes.  Replaces the reference's
 * write_text("\n".join(lines)) (onepass_dense_mix_run_custom_lang.py:889) and the raw / collapsed
 * writers (onepass_bilingual_mix_hub_custom_lang.py:942-962); callers that want the reference's
 * all-or-nothing files write to a temporary name and rename. */
CMX_API int cmx_trec_mono_file(const float* D, const int64_t* I, int64_t nq, int k, const char* qids,
                               const int64_t* qid_off, const char* docs, const int64_t* doc_off,
                               const int64_t* doc_keys, int64_t ndocs, const char* tag, int nthreads,
                               const char* path, int64_t* out_len);
CMX_API int cmx_trec_bilingual_file(const float* D, const int64_t* I, int64_t nq, int k, const char* qids,
                                    const int64_t* qid_off, const char* docs, const int64_t* doc_off,
                                    int64_t ndocs, const int32_t* base_code, const char* bases,
                                    const int64_t* base_off, int64_t nbases, const char* tag, int nthreads,
                                    const char* raw_path, const char* col_path, int64_t* raw_len,
                                    int64_t* col_len);

/* ---- collapse-by-base-id on the device ------------------------------------------------
 * replaces the grouping of collapse_run_max (onepass_bilingual_mix_hub_custom_lang.py:165-181), which re-parses
 * the raw run text: D, I [nq,k] and base_code [ndocs] (the base-id number of every corpus row; derived id =
 * "<base>#<lang>") are DEVICE pointers (or device aliases of pinned host memory) on `device`; per query the
 * groups come out in final order: col_code / col_val6 [nq,k] (base code, sign * rint(|score| * 1e6) = the value
 * the raw file prints), col_count [nq] (device).  Max over the 6-decimal rounded scores, stable descending order
 * (equal values keep the order in which their bases first appeared), hits with ids outside [0, ndocs) skipped.
 * *needs_host = 1 when some score cannot be carried exactly (non-finite, |score| >= 2^20, a negative zero): use
 * the host grouping of cmx_trec_bilingual_file for that batch.  Synchronous.
 * cmx_trec_bilingual_file_pre = cmx_trec_bilingual_file with the groups given (HOST arrays). */
CMX_API int cmx_collapse_max(const float* D, const int64_t* I, int64_t nq, int k, const int32_t* base_code,
                             int64_t ndocs, int32_t* col_code, int64_t* col_val6, int32_t* col_count,
                             int* needs_host, int device, void* stream);
CMX_API int cmx_trec_bilingual_file_pre(const float* D, const int64_t* I, int64_t nq, int k, const char* qids,
                                        const int64_t* qid_off, const char* docs, const int64_t* doc_off,
                                        int64_t ndocs, const int32_t* base_code, const char* bases,
                                        const int64_t* base_off, int64_t nbases, const int32_t* col_code,
                                        const int64_t* col_val6, const int32_t* col_count, const char* tag,
                                        int nthreads, const char* raw_path, const char* col_path,
                                        int64_t* raw_len, int64_t* col_len);

/* ---- instrumentation --------------------------------------------------------*/
typedef struct cmx_search_stats {
  int32_t path;            /* CMX_PATH_STREAM / CMX_PATH_TENSOR actually used       */
  int32_t slabs;           /* corpus slabs (score launches) of the last search      */
  int32_t reruns;          /* attempts that overflowed a candidate buffer and were
                              repeated: rescore -> split precision with the planned
                              slabs -> split precision with worst-case-safe slabs    */
  int32_t launches;        /* kernels launched by the last search                   */
  int64_t nq, ntotal;      /* shape of the last search                              */
  float score_ms;          /* CUDA-event time inside the scoring kernels (0 unless
                              cmx_set_profiling(1))                                 */
  float select_ms;         /* ... inside the compaction / finalize kernels          */
  float total_ms;          /* ... whole device part of the search                   */
  int32_t score_launches;
  int32_t select_launches;
} cmx_search_stats;
CMX_API int cmx_index_last_stats(const cmx_index* ix, cmx_search_stats* out);
/* 1: bracket kernel groups with CUDA events on the launching stream. */
CMX_API int cmx_set_profiling(int on);
/* arithmetic of the tensor path: CMX_PRECISION_RESCORE (default) or CMX_PRECISION_SPLIT. */
CMX_API int cmx_index_set_precision(cmx_index* ix, int mode);
/* Rescore precision, sharded corpus: the filter margin of every shard must be derived from the
 * maxima over ALL shards of the corpus row norm (out2[0]) and of the norm of what the fp16 plane
 * loses of a row (out2[1]).  cmx_index_error_bounds builds the plane if needed and reports this
 * shard's values; after a max-reduction over the shards cmx_index_raise_error_bounds installs
 * the global ones (values only ever grow; cmx_index_reset clears them). */
CMX_API int cmx_index_error_bounds(cmx_index* ix, float* out2);
CMX_API int cmx_index_raise_error_bounds(cmx_index* ix, const float* in2);
/* precision given to indexes created afterwards (process-wide default: CMX_PRECISION_RESCORE). */
CMX_API int cmx_set_default_precision(int mode);
/* tuning knob (tests): candidate-buffer capacity per query (0 = automatic). */
CMX_API int cmx_index_set_cand_capacity(cmx_index* ix, int cap);

#ifdef __cplusplus
}
#endif
#endif /* CMX_H_ */

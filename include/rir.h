/*
 * rir.h — C ABI of librir.so: the B200-native (sm_100a) retrieval hot path for
 * Mak-GIBA/research_image_retrieval.
 *
 * The reference has NO native/FFI interface: its boundary is a set of Python
 * call sites.  Each entry point below names the reference call
 * (file:line under the reference's src/benchmark/) that it replaces.  A
 * reference maintainer binds these with ctypes.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types in signatures;
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the
 *     parameter is documented "host";
 *   - the caller owns every buffer; the library never allocates device memory —
 *     scratch space is sized by the *_workspace() query and passed in;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream);
 *   - return value: RIR_OK (0) or a negative RIR_E_* code; rir_last_error()
 *     returns a thread-local human readable message for the last failure;
 *   - there is NO CPU fallback: on a device that is not compute capability 10.x
 *     every launching entry point returns RIR_E_ARCH.
 */
#ifndef RIR_H_
#define RIR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIR_VERSION 100 /* 0.1.0 */

/* error codes */
#define RIR_OK 0
#define RIR_E_ARG (-1)       /* bad argument (shape / alignment / dtype)            */
#define RIR_E_ARCH (-2)      /* current device is not sm_100 (B200)                 */
#define RIR_E_CUDA (-3)      /* a CUDA runtime / driver call failed                 */
#define RIR_E_WORKSPACE (-4) /* workspace too small for the requested problem       */

/* element types of descriptor / feature-map storage */
#define RIR_F32 0
#define RIR_BF16 1
#define RIR_FP8E4M3 2

/* pooling modes */
#define RIR_POOL_GEM 0 /* (mean_hw max(x,eps)^p)^(1/p)   networks/RetrievalNet.py:318-325 */
#define RIR_POOL_MAX 1 /* max_hw x (MAC)                 models/spoc.py:12-49 level 1     */
#define RIR_POOL_AVG 2 /* mean_hw max(x,eps) (SPoC)      networks/RetrievalNet.py:359-365 */

/* rir_sim_topk path selection */
#define RIR_PATH_AUTO 0
#define RIR_PATH_STREAM 1 /* TMA-bulk ring + CUDA-core dot products (tiny query batches) */
#define RIR_PATH_MMA 2    /* tcgen05 tensor-core contraction, TMEM accumulators          */
#define RIR_PATH_EXACT 3  /* one-CTA-per-query robust scan (overflow fallback; slow)     */
/* flag, OR-ed into `path`: the workspace header was initialised with rir_sim_topk_workspace_init and has since been
 * used only by successful rir_sim_topk / rir_sim_topk_sharded / rir_search_host calls (any shapes) — the search then
 * issues no memset launches (its select kernel leaves the header initialised for the next call). */
#define RIR_WS_CLEAN 0x100
/* flag, OR-ed into `path` of rir_sim_topk_sharded: ASYNCHRONOUS exchange.  Scan and select (which publishes this
 * rank's lists) run on `stream`; the merge that waits for the peers' lists runs on a library-owned side stream, so the
 * caller's stream goes straight on to the next search and the exchange latency (and the skew between ranks) hides
 * under that search's scan.  out_score / out_idx are valid once rir_exchange_join (stream order) or rir_exchange_sync
 * (host) has been called for the call's epoch; at most two searches may be outstanding per inbox (distinct outputs). */
#define RIR_EXCHANGE_ASYNC 0x200

/* per-(protocol,query) status written by rir_revisited_map / rir_compute_map */
#define RIR_MAP_OK 0
#define RIR_MAP_EMPTY_OK 1         /* no positives for this query: ap=+inf, excluded (utils/evaluate.py:65-68) */
#define RIR_MAP_NO_POS_RETRIEVED 2 /* positives exist, none in the ranked list: ap=0; with kappas the reference
                                      raises ValueError at utils/evaluate.py:101 — the host shim re-raises       */

int rir_version(void);
const char* rir_last_error(void);
/* RIR_OK iff the current device is compute capability 10.x. */
int rir_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Descriptor build: pooling -> L2 -> whitening -> L2 -> multi-scale aggregate
 * ------------------------------------------------------------------------------------------ */

/* Global spatial pooling of feature maps x[B,C,HW] (NCHW contiguous, HW = H*W) to out[B,C] fp32.
 * Replaces gem.forward (networks/RetrievalNet.py:318-325), GeMPooling.forward (models/gem_pooling.py:12-23),
 * G2Pooling.forward (models/senet_g2.py:132-153: out = alpha*gem + beta), AttentionBasedGlobalPooling.gem_pooling
 * (models/ultron_modules/ultron.py:193-205), spoc.forward (networks/RetrievalNet.py:359-365) and the level-1 max
 * pool of SpatialPyramidPooling (models/spoc.py:33-35).  dtype in {RIR_F32, RIR_BF16}.  alpha=1,beta=0 for plain GeM. */
int rir_pool(const void* x, int dtype, int B, int C, int HW, int mode, float p, float eps, float alpha, float beta,
             float* out, void* stream);

/* Row-wise L2 normalisation out[i,:] = x[i,:] / max(||x[i,:]||_2, eps); in-place allowed.
 * Replaces F.normalize(x, p=2, dim=-1) at networks/RetrievalNet.py:343,587,589, networks/spca.py:65,
 * models/gem_pooling.py:91, iris_evaluate.py:379-380 (torch default eps = 1e-12). */
int rir_l2_normalize(const float* x, int64_t n_rows, int d, float eps, float* out, void* stream);

/* Whitening / dimensionality reduction out[B,d_out] = x[B,C] * W[d_out,C]^T + bias, fp32 accumulate.
 * Replaces the 1x1 Conv2d / Linear `self.whiten` (networks/RetrievalNet.py:332,342,577,588; networks/spca.py:31-46,61-64)
 * with weights produced by ConvDimReduction.initialize_pca_whitening (networks/spca.py:215-227). bias may be NULL.
 * If l2_after != 0 each output row is L2-normalised (eps 1e-12) in the same launch sequence. */
int rir_whiten(const float* x, const float* W, const float* bias, int B, int C, int d_out, int l2_after, float* out,
               void* stream);

/* The descriptor head as ONE call (north star "GeM / L2-norm / whitening"): feature maps x[B,C,HW] ->
 *   pool (rir_pool semantics) -> [L2 if l2_before] -> W v + bias -> [L2 if l2_after] -> out[B, d_out] fp32.
 * Replaces the tail of GeM.forward_test (networks/RetrievalNet.py:337-344: l2_before = 0, l2_after = 1) and of
 * SOLAR.forward_test (networks/RetrievalNet.py:583-590: l2_before = 1, l2_after = 1).  Internally: the pooling kernel
 * (HBM-bound, the feature maps are read once) writes the pooled descriptors as exact bf16 pairs, a tcgen05 tensor-core
 * contraction with fp32 accumulation in TMEM applies the whitening (three split-bf16 passes: fp32-accurate to ~1e-6
 * relative), a finishing kernel adds bias and normalises — three launches chained with programmatic dependent launch.
 *   W12: whitening weights prepared ONCE by rir_whiten_prepare (rir_whiten_prepared_bytes(d_out, C) bytes, 16-byte
 *   aligned); NULL = no whitening layer (pool [+ L2] only, models/gem_pooling.py:86-92; out is [B, C], d_out ignored).
 *   workspace: rir_gem_l2_whiten_workspace(B, C, d_out) bytes (d_out = 0 without whitening), 256-byte aligned. */
size_t rir_whiten_prepared_bytes(int d_out, int C);
int rir_whiten_prepare(const float* W /* [d_out, C] fp32 */, int d_out, int C, void* W12, void* stream);
size_t rir_gem_l2_whiten_workspace(int B, int C, int d_out);
int rir_gem_l2_whiten(const void* x, int dtype, int B, int C, int HW, int mode, float p, float eps, float alpha, float beta,
                      const void* W12, const float* bias, int d_out, int l2_before, int l2_after, float* out,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The same whitening applied to already pooled descriptors x[B,C] fp32 (networks/RetrievalNet.py:342,588;
 * networks/spca.py:61-64) on the tensor cores; workspace as for rir_gem_l2_whiten.  rir_whiten (above) is the plain
 * fp32 CUDA-core version that takes raw weights. */
int rir_whiten_prepared(const float* x, const void* W12, const float* bias, int B, int C, int d_out, int l2_before,
                        int l2_after, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* PCA-whitening LEARN, dense part: column mean and covariance of descriptors X[N,D] (fp32, row-major):
 *   mean[D] = X.mean(0);  cov[D,D] = (X - mean)^T (X - mean) / N   (exactly symmetric)
 * Replaces networks/backbone.py:47-50 of pcawhitenlearn_shrinkage; the eigen-decomposition (:51-56) is done by the
 * host (research_image_retrieval_b200/whitening.py) and its W, b feed rir_whiten (networks/spca.py:215-227).
 * The covariance runs on the tensor cores (split-bf16 tcgen05 contraction over the upper triangle, fp32 accumulate);
 * the workspace holds the transposed bf16 copy of X (4 * D * N bytes) — 256-byte aligned. */
size_t rir_pca_covariance_workspace(int64_t N, int D);
int rir_pca_covariance(const float* X, int64_t N, int D, float* mean, float* cov, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Multi-scale aggregation of extract_vectors (utils/helpfunc.py:31-44): v[N,S,D] per-scale descriptors,
 * keep[N,S] (1 = scale used, 0 = dropped because the resized image was < 36 px; NULL = all kept);
 * out[n,:] = L2( sum_s keep*v[n,s,:] / #kept ). */
int rir_scale_mean_l2(const float* v, const uint8_t* keep, int64_t N, int S, int D, float* out, void* stream);

/* Cast fp32 descriptors v[n,d] into the database layout sim_topk reads: row-major bf16, or fp8 e4m3 with a
 * per-row fp32 scale (row = scale * fp8).  (The reference keeps descriptors as fp32 torch tensors:
 * utils/helpfunc.py:21,27 — this is the B200 storage format for them.)  scale_out ignored for bf16. */
int rir_pack_descriptors(const float* v, int64_t n, int d, int dtype_out, void* out, float* scale_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Search: similarity + top-k (never materialises the [nq, n] score matrix)
 * ------------------------------------------------------------------------------------------ */

/* Bytes of scratch rir_sim_topk needs for this problem (0 on invalid arguments). */
size_t rir_sim_topk_workspace(int nq, int64_t n_local, int d, int k, int dtype);

/* Exact top-k of S = Q X^T per query, descending score, ties -> ascending index.
 * Replaces torch.mm(q, g.t()) + np.argsort(-sim, axis=1) (iris_evaluate.py:383-386) and
 * compute_similarity + torch.topk (reference/manus/7_AdaptiveHybridModel/modified/
 * adaptive_hybrid_retrieval_complete.py:11-16,428).
 *   Q[nq,d], X[n_local,d] row-major, dtype in {RIR_BF16, RIR_FP8E4M3, RIR_F32}; d*elemsize % 16 == 0, 16-byte aligned
 *   bases.  RIR_F32 keeps the reference's fp32 arithmetic (CUDA-core stream path only, for ROxford/RParis-size sets);
 *   q_scale[nq], x_scale[n_local]: per-row dequantisation scales (NULL = 1);
 *   1 <= k <= n_local; k <= 8192, or any k <= n_local when n_local <= 16384 (k == n_local yields the full ranking);
 *   idx_offset is added to every returned row index (global index of shard row 0);
 *   out_score[nq,k] fp32, out_idx[nq,k] int32. */
int rir_sim_topk(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                 int64_t n_local, int d, int k, int64_t idx_offset, float* out_score, int32_t* out_idx,
                 void* workspace, size_t workspace_bytes, int path, void* stream);

/* One-time initialisation of a search workspace (the first bytes of a rir_sim_topk / rir_search_host workspace are a
 * fixed-layout header: per-query thresholds, candidate counters, the grid-barrier counter).  Enqueues two memsets on
 * `stream`.  Calls that pass RIR_WS_CLEAN in `path` rely on it; calls without the flag initialise the header
 * themselves (two extra launches per search).  Re-initialise after a call that returned an error. */
int rir_sim_topk_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/* Measurement hook (bench.py's roofline).  rir_profile_scan_begin arms THIS host thread: every full-scan kernel
 * launch of the following rir_sim_topk / _sharded / rir_search_host calls is bracketed by a pair of CUDA events from a
 * library-owned pool, on the call's stream (a search of more than 4096 queries launches one scan per query group —
 * each is recorded).  rir_profile_scan_end disarms, synchronises on the recorded events and returns the number of
 * scan launches since begin in *n_launches and the duration of launch i (milliseconds) in ms_out[i], i < cap. */
int rir_profile_scan_begin(void);
int rir_profile_scan_end(float* ms_out, int cap, int* n_launches);
/* Sampled profiling: while paused (non-zero) an armed thread records nothing.  An event record between two kernels
 * costs a few microseconds of stream time and breaks their programmatic-dependent-launch overlap — at 8-way sharding
 * that is several percent of a step — so bench.py brackets only every 4th step's scan. */
int rir_profile_scan_pause(int paused);

/* Development hook: event timeline of the tcgen05 scan.  dev_buf = device buffer of (1 + 2 * cap_events) uint64, zeroed
 * by the caller; the following rir_sim_topk calls of THIS host thread append (meta, %globaltimer ns) pairs —
 * meta = cta << 32 | event << 16 | round; events: 0 kernel start, 1 first database TMA of the round, 2 MMA round start,
 * 3 MMA round issued, 4 accumulator ready (epilogue), 5 epilogue of the round done, 6 threshold phase begin, 7 grid
 * barrier passed, 8 all thresholds published, 9 CTA done.  NULL disarms.  See tools/timeline.py. */
int rir_profile_timeline(void* dev_buf, int cap_events);

/* Re-score a candidate list with higher-precision rows and keep the best k (same order rule).  Used after an fp8
 * scan that returned k_in > k candidates, so the final list meets the fp8 bar "5e-3 relative against an fp32 rescore".
 *   Q[nq,d], X[n_local,d] in dtype RIR_BF16 or RIR_F32 (the rescoring copy of the shard); ix_in[nq,k_in] GLOBAL row
 *   indices (-1 or rows outside [idx_offset, idx_offset+n_local) are skipped); out[nq,k], k <= k_in <= 8192.
 * No reference counterpart (the reference only has fp32 rows). */
int rir_rescore_topk(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                     int64_t n_local, int64_t idx_offset, int d, const int32_t* ix_in, int k_in, int k,
                     float* out_score, int32_t* out_idx, void* stream);

/* k-way merge of G per-shard top-k lists sc/ix[G,nq,k] (as produced by an allgather of rir_sim_topk outputs)
 * into the global top-k (same order rule).  No reference counterpart. */
size_t rir_merge_topk_workspace(int G, int nq, int k);
int rir_merge_topk(const float* sc, const int32_t* ix, int G, int nq, int k, float* out_sc, int32_t* out_ix,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sharded search over NVLink peer memory (one process per GPU; no reference counterpart)
 *
 * Every rank owns an inbox of rir_exchange_bytes(G, nq_max, k_max) bytes that all peers map with CUDA IPC.
 * rir_sim_topk_sharded = rir_sim_topk on the local shard whose select kernel STORES the local top-k into slot `rank`
 * of every rank's inbox (P2P stores + a release flag per query), followed by a merge kernel that waits for the G
 * flags of each query and merges the G lists: out_score / out_idx [nq,k] hold the GLOBAL top-k on every rank.  There
 * is no collective-library call on the data path.  Semantics of a collective: every rank calls it with the same nq,
 * k and `epoch` (1, 2, 3, ... per call).  inbox[g] = rank g's inbox as mapped in this process (own allocation for
 * g == rank).  A shard may hold fewer than k rows (it publishes what it has).
 * These are the only entry points that allocate device memory (an IPC-exportable cudaMalloc region).
 * ------------------------------------------------------------------------------------------ */
size_t rir_exchange_bytes(int G, int nq_max, int k_max);
int rir_peer_alloc(size_t bytes, void** ptr);             /* zero-filled */
int rir_peer_free(void* ptr);
int rir_peer_export(void* ptr, void* handle64);           /* 64-byte cudaIpcMemHandle_t, host buffer */
int rir_peer_open(const void* handle64, void** ptr);      /* in another process */
int rir_peer_close(void* ptr);
int rir_sim_topk_sharded(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                         int64_t n_local, int d, int k, int64_t idx_offset, float* out_score, int32_t* out_idx,
                         void* workspace, size_t workspace_bytes, int path, void* stream, int G, int rank,
                         uint32_t epoch, int nq_max, int k_max, void* const* inbox /* host array [G] */);
/* RIR_EXCHANGE_ASYNC: make `stream` (join) or the host (sync) wait for the merge of the search issued with `epoch` on
 * the inbox `own_inbox` (= inbox[rank]).  No-ops when no asynchronous search was issued on that inbox. */
int rir_exchange_join(const void* own_inbox, uint32_t epoch, void* stream);
int rir_exchange_sync(const void* own_inbox, uint32_t epoch);

/* The whole query path as one call on HOST buffers (what the reference call site holds: CPU fp32 query features,
 * iris_evaluate.py:378-386): H2D of q_host[nq,d] fp32 (pinned memory recommended) -> pack to the shard's dtype ->
 * rir_sim_topk (or rir_sim_topk_sharded when G > 1; pass G = 1, inbox = NULL otherwise) -> D2H of the top-k into
 * out_score_host / out_idx_host [nq,k].  Everything is enqueued on `stream`; synchronise the stream before reading
 * the outputs.  d must already be a multiple of 16 bytes in the shard's dtype.  Without exchange k <= n_local. */
size_t rir_search_host_workspace(int nq, int64_t n_local, int d, int k, int dtype);
int rir_search_host(const float* q_host, const void* X, int dtype, const float* x_scale, int nq, int64_t n_local, int d,
                    int k, int64_t idx_offset, float* out_score_host, int32_t* out_idx_host, void* workspace,
                    size_t workspace_bytes, int path, void* stream, int G, int rank, uint32_t epoch, int nq_max,
                    int k_max, void* const* inbox);

/* alpha query expansion (skeleton reference/manus/1_SPARSE/sparse_model.py:374-405):
 *   acc[q,:] (+)= sum_{j<kq, idx_off <= ix[q,j] < idx_off+n_local} max(sc[q,j],0)^alpha * dequant(X[ix[q,j]-idx_off,:])
 * accumulates the contribution of the rows this shard owns (acc fp32 [nq,d], zero it first; all-reduce across shards). */
int rir_aqe_accumulate(const void* X, int dtype, const float* x_scale, int64_t n_local, int64_t idx_offset, int d,
                       const float* sc, const int32_t* ix, int nq, int ld_topk, int kq, float alpha, float* acc,
                       void* stream);
/* q'[q,:] = L2(dequant(Q[q,:]) + acc[q,:]); written as fp32 (out_f32, may be NULL) and in the search dtype
 * (out_q + out_scale, may be NULL). */
int rir_aqe_finalize(const void* Q, int dtype, const float* q_scale, const float* acc, int nq, int d, float* out_f32,
                     void* out_q, float* out_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Revisited-protocol evaluation
 * ------------------------------------------------------------------------------------------ */

/* Junk-aware mAP and P@k for P protocols in one launch.
 * Replaces compute_ap / compute_map (utils/evaluate.py:4-34, 37-150; duplicate iris_evaluate.py:11-187).
 *   ranks[nq, L] int32, row q = ranked db ids of query q (best first), -1 = padding (truncated / ragged lists);
 *   ld = row stride in elements (>= L).  (The reference layout is ranks[L, nq]; the host shim transposes.)
 *   Three sorted id lists per query in CSR form: lists A, B, C (x_ids = concatenated ids, x_off[nq+1] = offsets;
 *   a NULL list is empty).
 *   proto[P] selects membership per protocol p: bits 0-2 = lists whose union is `ok`, bits 4-6 = lists whose
 *   union is `junk` (bit0/4 = A, bit1/5 = B, bit2/6 = C).  Revisited protocols (utils/evaluate.py:163-185) with
 *   A=easy, B=hard, C=junk:  Easy 0x61 (ok=A, junk=C|B), Medium 0x43 (ok=A|B, junk=C), Hard 0x52 (ok=B, junk=C|A).
 *   kappas[nk] int32 (nk may be 0).  proto and kappas are small HOST arrays (copied into the launch parameters).
 * Outputs (fp64, identical operation order to the reference's Python floats):
 *   map[P], aps[P,nq] (+inf for empty ok), mpr[P,nk], prs[P,nq,nk], status[P,nq] (RIR_MAP_*). */
int rir_compute_map(const int32_t* ranks, int nq, int64_t L, int64_t ld, const int32_t* a_ids, const int32_t* a_off,
                    const int32_t* b_ids, const int32_t* b_off, const int32_t* c_ids, const int32_t* c_off,
                    const int32_t* proto, int P, const int32_t* kappas, int nk, double* map, double* aps, double* mpr,
                    double* prs, int32_t* status, void* stream);

/* The same evaluation for a COMPACT list: row q holds only the query's ground-truth ids, ordered by their position
 * in the full ranking, and positions[nq, ld] gives that position (0-based, ascending along the row; ignored where
 * ranks is -1).  Everything else as rir_compute_map.  Used with rir_rank_count below when the full ranked list is
 * never materialised (R1M: 1M distractors per query). */
int rir_compute_map_at(const int32_t* ranks, const int32_t* positions, int nq, int64_t L, int64_t ld, const int32_t* a_ids,
                       const int32_t* a_off, const int32_t* b_ids, const int32_t* b_off, const int32_t* c_ids,
                       const int32_t* c_off, const int32_t* proto, int P, const int32_t* kappas, int nk, double* map,
                       double* aps, double* mpr, double* prs, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * Positions of ground-truth ids in the full ranking WITHOUT sorting the database: what
 * `np.arange(N)[np.in1d(ranks[:, i], ok)]` (utils/evaluate.py:76-80) reads off the reference's full
 * np.argsort (iris_evaluate.py:386).  position(p) = #{rows x : (score(x), -index(x)) > (score(p), -index(p))} —
 * a count, so it shards: each rank counts over its rows, one all-reduce adds the counts.
 *   ids / off: CSR of the GLOBAL row ids to locate per query (off[nq+1], total_ids = off[nq]).
 *   1. rir_gnd_scores: scores[i] = <q, x_id> for ids whose row lives in this shard, 0 otherwise (all-reduce SUM over
 *      the shards gives every rank every score);
 *   2. rir_rank_thresholds: keys_sorted[nq, m_pad] (uint64 ranking keys: ordered score << 32 | ~id), each row sorted
 *      descending and 0-padded; m_pad = power of two in [32, 4096] >= the longest id list;
 *   3. rir_rank_count: counts[nq, m_pad] = shard rows that outrank threshold j of query q (all-reduce SUM over the
 *      shards gives the position); workspace >= nq * m_pad * 4 bytes.  Queries and rows as in rir_sim_topk.
 * ------------------------------------------------------------------------------------------ */
int rir_gnd_scores(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                   int64_t n_local, int d, int64_t idx_offset, const int32_t* ids, const int32_t* off, int64_t total_ids,
                   float* scores, void* stream);
int rir_rank_thresholds(const float* scores, const int32_t* ids, const int32_t* off, int nq, int m_pad,
                        uint64_t* keys_sorted, void* stream);
int rir_rank_count(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                   int64_t n_local, int d, int64_t idx_offset, const uint64_t* keys_sorted, int m_pad, int32_t* counts,
                   void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIR_H_ */

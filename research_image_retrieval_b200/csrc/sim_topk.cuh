// sim_topk.cuh — parameters shared by the similarity + top-k kernels.
//
// Replaces torch.mm(q, g.t()) + np.argsort / torch.topk (iris_evaluate.py:383-386;
// reference/manus/7_AdaptiveHybridModel/modified/adaptive_hybrid_retrieval_complete.py:11-16,428)
// without ever writing the [nq, n] score matrix to HBM.
//
// Scheme (exact):
//   1. SAMPLE       : score a spread-out subset of 256-row blocks; keep, per query, the best few keys of every block.
//   2. threshold    : tau[q] = (a lower bound of) the k-th best key among the kept ones.  They are scores of distinct
//                     rows of the shard, so tau[q] <= the true k-th best key: filtering can never drop a top-k row.
//   3. SCAN         : score every row once; rows whose key >= tau[q] are appended to cand[q, cap] (rare: ~k*n/S).
//   4. final select : exact top-k of cand[q] (radix select + bitonic sort) -> out (or -> the peers' inboxes).
//   5. overflow     : a query whose candidate list overflowed `cap` (adversarial row order) is redone by a
//                     one-CTA exact scan.  Small shards (n <= 16384) skip 1-2 and keep every row.
// tcgen05 path (kModeFused): 1-3 are ONE launch — the first round of the scan is the sample, tau is derived inside
// the kernel behind a grid barrier (sim_topk_mma.cu).  Stream path / large k: three launches (sample, threshold, scan).
#pragma once
#include "rir_common.cuh"

namespace rir {

constexpr int kSampleBlockRows = 256;  // rows per sample block == MMA tile N

enum SimMode : int { kModeScanFilter = 0, kModeSample = 1, kModeScanAll = 2, kModeFused = 3 };
constexpr uint32_t kTauUnset = 0xFFFFFFFFu;  // tau_score[q] before the fused scan has published it (a NaN pattern)
constexpr int kFusedTopT = 8;  // keys kept per (query, first-phase tile) in the fused scan

// Sharded search without a collective library call: every rank owns an INBOX in its HBM that all peers can write
// over NVLink (CUDA IPC mappings).  The select kernel of rank r stores its sorted local top-k (global-index keys) into
// slot r of every rank's inbox and then raises a per-(slot, query) flag to the call's epoch (release, system scope);
// the merge kernel of each rank waits for the G flags of its query (acquire) and merges.  Two parities alternate so
// a rank that is one call ahead never overwrites data a slower peer is still merging.
constexpr int kMaxPeers = 16;
struct Exchange {
  int G;            // ranks (0 = no exchange: results go to out_score / out_idx)
  int rank;
  uint32_t epoch;   // call counter, identical on every rank, starts at 1
  int nq_max, k_max;
  int q_base;       // global query index of this launch's query 0
  int k_push;       // list length every rank publishes (the global k; a shard shorter than k pads with 0 keys)
  int fold;         // 1: the select kernel also waits for the peers' lists and merges (no separate merge launch)
  unsigned long long* inbox[kMaxPeers];  // inbox of rank g as mapped in THIS process
};
__host__ __device__ __forceinline__ size_t exchange_key_slots(int G, int nq_max, int k_max) {
  return (size_t)2 * G * nq_max * k_max;
}
__host__ __device__ __forceinline__ size_t exchange_bytes(int G, int nq_max, int k_max) {
  return exchange_key_slots(G, nq_max, k_max) * 8 + (size_t)2 * G * nq_max * 4;
}

// keys of (parity b, source rank g, query q) inside an inbox
__host__ __device__ __forceinline__ unsigned long long* exchange_keys(unsigned long long* inbox, const Exchange& ex,
                                                                      int b, int g, int q) {
  return inbox + (((size_t)b * ex.G + g) * ex.nq_max + q) * ex.k_max;
}
__host__ __device__ __forceinline__ uint32_t* exchange_flag(unsigned long long* inbox, const Exchange& ex, int b, int g,
                                                            int q) {
  return reinterpret_cast<uint32_t*>(inbox + exchange_key_slots(ex.G, ex.nq_max, ex.k_max)) +
         ((size_t)b * ex.G + g) * ex.nq_max + q;
}

struct SimParams {
  const void* Q;         // [nq, d]
  const void* X;         // [n, d]
  const float* q_scale;  // [nq] or nullptr
  const float* x_scale;  // [n]  or nullptr
  int nq;                // queries handled by this launch (<= 128 per MMA query block; any for select kernels)
  int q0;                // first query of this launch (offset into Q / tau / cnt / cand / sample_scores rows)
  long long n;           // rows in this shard
  int d;
  int row_bytes;         // d * sizeof(element)
  int mode;              // SimMode
  int nblk;              // ceil(n / 256)
  int sblk;              // sample blocks (<= nblk)
  float* sample_scores;  // [nq_total, sblk*256]   dense sample (topt == 0)
  int topt;              // > 0: the sample pass keeps only the best `topt` keys per query per CTA (MMA path) or the best
                         //      key per query per consumer warp (stream path) in sample_keys — enough to bound the k-th
  int sample_m;          // slots per query in sample_keys (unwritten slots are 0 == "nothing")
  unsigned long long* sample_keys;  // [nq_total, sample_m], aliases the sample_scores region
  float* tau_score;      // [nq_total]
  uint32_t* tau_idx;     // [nq_total]
  uint32_t* cnt;         // [nq_total] candidates appended per query (may exceed cap -> overflow)
  unsigned long long* cand;  // [nq_total, cap]
  int cap;
  // fused scan (kModeFused, tcgen05 path): the first round(s) of the scan itself are the sample — each CTA keeps the
  // best kFusedTopT keys per query of its first tile(s) in sample_keys, a grid barrier follows, tau is computed by the
  // scan kernel and the remaining rounds filter with it.  Tiles are visited in a multiplicative permutation so the
  // first `fused_tiles` of them are spread over the whole shard.
  uint32_t* gbar;        // grid-barrier counter (zeroed by the host before the launch); tau_score[q] is preset to
                         // kTauUnset and carries its own "published" state
  int fused_tiles;       // first-phase tiles == slots per query in sample_keys / kFusedTopT
  long long perm_mul;    // physical tile = (virtual tile * perm_mul) % perm_n
  long long perm_n;      // number of database tiles
  int tile_rows;         // rows per database tile of the fused scan (256 or 128)
  int range_rows;        // range mode (sim_topk_mma.cu): first-phase slot s covers rows [s * range_rows, + first_rows);
  int first_rows;        //   0 = classic mode (slot s is tile (s * perm_mul) % perm_n)
  int k;                 // top-k requested (the fused scan computes tau itself)
  Exchange ex;           // sharded search: push the local top-k to the peers instead of writing out_score / out_idx
  // development: per-CTA event timeline of the tcgen05 scan (rir_profile_timeline); null in production
  unsigned long long* timeline;  // [0] = event counter, then (meta, globaltimer ns) pairs
  int timeline_cap;              // events that fit
};

// first row of sample block j (strided over the whole shard so clustered / sorted databases are sampled fairly)
__host__ __device__ __forceinline__ long long sample_block_row0(int j, int nblk, int sblk) {
  return ((long long)j * nblk / sblk) * (long long)kSampleBlockRows;
}

#ifdef __CUDACC__
// append one candidate; cnt is allowed to run past cap (detected by the final select)
__device__ __forceinline__ void push_candidate(const SimParams& p, int q, float score, uint32_t idx) {
  const uint32_t slot = atomicAdd(&p.cnt[q], 1u);
  if (slot < (uint32_t)p.cap) p.cand[(size_t)q * p.cap + slot] = make_key(score, idx);
}
__device__ __forceinline__ bool passes(float score, uint32_t idx, float ts, uint32_t ti) {
  return score > ts || (score == ts && idx <= ti);
}
#endif

constexpr int RIR_E_NOFUSE = -100;  // internal: the fused one-launch scan is unavailable, take the three-launch route

// kernels' host launchers (defined in the .cu files)
int launch_sim_stream(const SimParams& p, int dtype, cudaStream_t st);
// fills p.topt / sample_m / fused_tiles / perm_* when p.mode == kModeFused
int launch_sim_mma(SimParams& p, int dtype, cudaStream_t st);
// can the tcgen05 path run this problem as ONE fused launch (first-phase sample + in-kernel threshold)?
bool mma_can_fuse(int nq, long long n, int k);
int launch_sample_threshold(const SimParams& p, int nq_total, int k, cudaStream_t st);
int launch_final_select(const SimParams& p, int dtype, int nq_total, int k, long long idx_offset, float* out_score,
                        int32_t* out_idx, uint32_t* ovf, cudaStream_t st);
bool select_handles_overflow(int k);  // the select kernel redoes overflowed queries itself (no fallback launch needed)
bool select_can_fold_merge(int nq_total, int G, int k_push);
int launch_merge_exchange(const Exchange& ex, int nq, int k, float* out_score, int32_t* out_idx, cudaStream_t st,
                          bool side_stream = false);
int launch_exact_scan(const SimParams& p, int dtype, int nq_total, int k, long long idx_offset, float* out_score,
                      int32_t* out_idx, const uint32_t* ovf /*nullptr = all queries*/, cudaStream_t st);

}  // namespace rir

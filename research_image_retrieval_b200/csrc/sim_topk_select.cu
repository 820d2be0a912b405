// sim_topk_select.cu — the ranking side of sim_topk: sample threshold, final exact select, robust fallback scan,
// and the cross-shard k-way merge.
//
// Replaces np.argsort(-similarity, axis=1) (iris_evaluate.py:386) / torch.topk(similarity, k)
// (reference/manus/7_AdaptiveHybridModel/modified/adaptive_hybrid_retrieval_complete.py:428): the [nq, n] matrix is
// never sorted; only the candidates that survive the per-query threshold are.
#include <stdlib.h>
#include "sim_topk.cuh"
#include "topk_select.cuh"
#include "dot_rows.cuh"

namespace rir {

constexpr int kSelectThreads = 512;  // upper bound (launch bounds, shared-memory sizing)

// threads per query CTA of the select / merge kernels (development override RIR_SELECT_THREADS; measured per-step
// overhead beyond the scan, 70 queries: 512 threads 28.4 us, 256 threads 31.7 us, 128 threads 38.3 us)
static int select_threads() {
  static int n = 0;
  if (n == 0) {
    n = 512;
    if (const char* e = getenv("RIR_SELECT_THREADS")) {
      const int v = atoi(e);
      if (v == 128 || v == 256 || v == 512) n = v;
    }
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// tau[q] = k-th best key of the dense sample scores
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelectThreads) sample_threshold_kernel(const SimParams p, int k, int kpad) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* dst = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ SelectScratch sc;
  const int q = blockIdx.x;
  if (p.topt > 0) {
    // kept keys already carry their row index
    const unsigned long long* kk = p.sample_keys + (size_t)q * p.sample_m;
    auto key_at = [=](int i) -> unsigned long long { return kk[i]; };
    const int got = block_select_topk(key_at, p.sample_m, k, dst, kpad, &sc);
    if (threadIdx.x == 0) {
      if (got >= k && dst[k - 1] != 0ull && key_score(dst[k - 1]) > -INFINITY) {
        p.tau_score[q] = key_score(dst[k - 1]);
        p.tau_idx[q] = key_index(dst[k - 1]);
      } else {  // fewer than k kept keys: accept everything (the candidate list then overflows into the exact path)
        p.tau_score[q] = -INFINITY;
        p.tau_idx[q] = 0xFFFFFFFFu;
      }
    }
    return;
  }
  const int m = p.sblk * kSampleBlockRows;
  const float* s = p.sample_scores + (size_t)q * m;
  const int nblk = p.nblk, sblk = p.sblk;
  // sample position i is monotonic in the database row index, so it can stand in for the row inside the key
  // (no 64-bit division per access); only the selected k-th key is mapped back to its row.
  auto key_at = [=](int i) -> unsigned long long { return make_key(s[i], (uint32_t)i); };
  const int got = block_select_topk(key_at, m, k, dst, kpad, &sc);
  if (threadIdx.x == 0) {
    // got == k whenever the host sized the sample (S >= k valid rows); otherwise fall back to "accept everything"
    if (got >= k && key_score(dst[k - 1]) > -INFINITY) {
      const int pos = (int)key_index(dst[k - 1]);
      p.tau_score[q] = key_score(dst[k - 1]);
      p.tau_idx[q] = (uint32_t)(sample_block_row0(pos / kSampleBlockRows, nblk, sblk) + (pos % kSampleBlockRows));
    } else {
      p.tau_score[q] = -INFINITY;
      p.tau_idx[q] = 0xFFFFFFFFu;
    }
  }
}

int launch_sample_threshold(const SimParams& p, int nq_total, int k, cudaStream_t st) {
  const int kpad = pow2_ceil_int(k < 32 ? 32 : k);
  const size_t smem = (size_t)kpad * sizeof(uint64_t);
  RIR_CUDA_OK(ensure_dyn_smem(sample_threshold_kernel, smem));
  sample_threshold_kernel<<<nq_total, kSelectThreads, smem, st>>>(p, k, kpad);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// ---------------------------------------------------------------------------------------------
// final: exact top-k of each query's candidate list
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_sorted(const uint64_t* dst, int got, int k, long long idx_offset,
                                             float* out_score, int32_t* out_idx) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = dst[i];
    if (i < got && key != 0ull) {
      out_score[i] = key_score(key);
      out_idx[i] = (int32_t)((long long)key_index(key) + idx_offset);
    } else {
      out_score[i] = -INFINITY;
      out_idx[i] = -1;
    }
  }
}

// Sharded search: store this rank's sorted top-k of query q (keys re-based to GLOBAL row indices, 0 = padding) into
// slot `rank` of every rank's inbox, then publish it.  All threads of the block must call.
__device__ __forceinline__ void push_sorted_to_peers(const Exchange& ex, int q, const uint64_t* dst, int got, int k,
                                                     long long idx_offset) {
  const int b = (int)(ex.epoch & 1u);
  const int qg = ex.q_base + q;
  const int kp = ex.k_push;
  for (int i = threadIdx.x; i < ex.G * kp; i += blockDim.x) {
    const int g = i / kp, j = i - g * kp;
    const uint64_t key = (j < got && j < k) ? dst[j] : 0ull;
    uint64_t out = 0ull;
    if (key != 0ull) out = make_key(key_score(key), (uint32_t)((long long)key_index(key) + idx_offset));
    exchange_keys(ex.inbox[g], ex, b, ex.rank, qg)[j] = out;
  }
  // The key stores of all threads happen-before the flag store through the block barrier, and a release at system
  // scope is cumulative over that order (the pattern of a cooperative-groups grid sync): ONE fence + release store by
  // the flag-writing threads instead of a system-scope fence in all 512 (each of which waits out an NVLink round trip).
  __syncthreads();
  if ((int)threadIdx.x < ex.G) {
    uint32_t* f = exchange_flag(ex.inbox[threadIdx.x], ex, b, ex.rank, qg);
    asm volatile("fence.acq_rel.sys;\n\tst.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(ex.epoch) : "memory");
  }
}

// Sharded search with the merge FOLDED into the select kernel (ex.fold: every query's CTA is co-resident, i.e. at most
// one CTA per SM): push this rank's list, wait until all G ranks have published query q for this epoch, merge the G
// sorted lists out of this rank's own inbox and write the global top-k — no separate merge launch.  Nobody waits
// before having pushed, and every waiting CTA is resident, so the wait cannot dead-lock.  `lists` holds >= G * k_push
// keys, `mdst` >= k_push keys (shared memory).  All threads of the block must call.
__device__ __forceinline__ void exchange_and_merge(const Exchange& ex, int q, const uint64_t* sorted, int got, int k,
                                                   long long idx_offset, uint64_t* lists, uint64_t* mdst, float* out_score,
                                                   int32_t* out_idx) {
  push_sorted_to_peers(ex, q, sorted, got, k, idx_offset);
  if (!ex.fold) return;
  const int b = (int)(ex.epoch & 1u);
  const int qg = ex.q_base + q;
  unsigned long long* mine = ex.inbox[ex.rank];
  if ((int)threadIdx.x < ex.G) {
    const uint32_t* f = exchange_flag(mine, ex, b, threadIdx.x, qg);
    const long long t0 = clock64();
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == ex.epoch) break;
      __nanosleep(100);
      if (clock64() - t0 > 20000000000ll) {
        printf("librir: rank %d never received query %d of rank %d (epoch %u, saw %u)\n", ex.rank, qg, (int)threadIdx.x,
               ex.epoch, v);
        __trap();
      }
    }
  }
  __syncthreads();
  const int kp = ex.k_push;
  for (int i = threadIdx.x; i < ex.G * kp; i += blockDim.x) {
    const int g = i / kp, j = i - g * kp;
    lists[i] = __ldcg(exchange_keys(mine, ex, b, g, qg) + j);
  }
  __syncthreads();
  block_merge_sorted_lists(lists, ex.G, kp, mdst);
  write_sorted(mdst, kp, kp, 0, out_score + (size_t)qg * kp, out_idx + (size_t)qg * kp);
}

// ---------------------------------------------------------------------------------------------
// robust exact scan of ONE query by one CTA: running top-k in shared memory.  Used (a) for queries whose candidate
// list overflowed — called from the select kernel itself, so the common no-overflow case costs no extra launch —
// and (b) as RIR_PATH_EXACT, an independent second implementation for the parity tests.
// ---------------------------------------------------------------------------------------------
// running-buffer size of the exact redo inside the select kernel: a power of two with room for k survivors plus two
// iterations of pushes (kSelectThreads / 32 warps x 4 rows each)
__host__ __device__ __forceinline__ int inline_exact_bufcap(int k) {
  const int need = k + 2 * (kSelectThreads / 32) * 4;
  int c = 64;
  while (c < need) c <<= 1;
  return c;
}

template <int DT>
__device__ void exact_scan_body(const SimParams& p, int q, int k, int bufcap, uint64_t* buf, float* qs,
                                long long idx_offset, float* out_score, int32_t* out_idx, uint64_t* merge_lists = nullptr,
                                uint64_t* merge_dst = nullptr) {
  __shared__ int count;
  __shared__ unsigned long long tau;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_iter = (int)(blockDim.x >> 5) * 4;  // 4 rows per warp and iteration
  load_query_f32<DT>(p, q, qs);
  for (int i = threadIdx.x; i < bufcap; i += blockDim.x) buf[i] = 0ull;
  if (threadIdx.x == 0) { count = 0; tau = 0ull; }
  __syncthreads();
  const int chunks = p.row_bytes >> 4;
  for (long long base = 0; base < p.n; base += rows_per_iter) {
    const unsigned long long t = tau;
    {
      const long long b4 = base + warp * 4;
      const int nvalid = b4 >= p.n ? 0 : (int)((p.n - b4) < 4 ? (p.n - b4) : 4);
      if (nvalid > 0) {
        float sc4[4];
        dot_rows4<DT>(reinterpret_cast<const uint8_t*>(p.X) + (size_t)b4 * p.row_bytes, (size_t)p.row_bytes, nvalid, qs,
                      chunks, lane, sc4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r >= nvalid) continue;
          float s = sc4[r];
          if (p.x_scale) s *= p.x_scale[b4 + r];
          const unsigned long long key = make_key(s, (uint32_t)(b4 + r));
          if (lane == 0 && key > t) buf[atomicAdd(&count, 1)] = key;
        }
      }
    }
    __syncthreads();
    if (count > bufcap - rows_per_iter) {  // uniform branch: compact to the best k
      block_bitonic_sort_desc(buf, bufcap);
      if (threadIdx.x == 0 && count >= k) {
        count = k;
        tau = buf[k - 1];
      }
      for (int i = k + threadIdx.x; i < bufcap; i += blockDim.x) buf[i] = 0ull;
    }
    __syncthreads();
  }
  block_bitonic_sort_desc(buf, bufcap);
  const int got = count < k ? count : k;
  if (p.ex.G > 0 && merge_lists != nullptr) {
    // the sorted list moves to merge_dst first: merge_lists may alias buf
    for (int i = threadIdx.x; i < k; i += blockDim.x) merge_dst[i] = i < got ? buf[i] : 0ull;
    __syncthreads();
    exchange_and_merge(p.ex, q, merge_dst, got, k, idx_offset, merge_lists, merge_dst, out_score, out_idx);
  } else if (p.ex.G > 0) {
    push_sorted_to_peers(p.ex, q, buf, got, k, idx_offset);
  } else {
    write_sorted(buf, got, k, idx_offset, out_score + (size_t)q * k, out_idx + (size_t)q * k);
  }
}

constexpr int kMergeStage = 8192;  // G*k keys the exchange merge stages in shared memory (64 KB)
constexpr int kStageKeys = 8192;  // candidates staged in shared memory (64 KB) so the select passes do not re-read L2
constexpr int kMaxRedo = 64;      // first-phase tiles per query that may need a re-score before the exact path takes over

// Fused scan: the first-phase tiles left only their best kFusedTopT keys per query (sample_keys).  Keys >= tau join
// the candidate list here (`list`, holding `*m_io` keys, room for `limit`; shared or global memory).  A tile whose
// LAST kept key still passes tau may hold more passing rows than were kept: it is re-scored with CUDA-core dot
// products (rare: P[>= 8 of 256 rows above the ~k/148-per-tile rate]).  Returns false when the query must be redone
// exactly (too many such tiles, or the list ran out of room).
template <int DT>
__device__ __forceinline__ bool merge_first_phase(const SimParams& p, int q, float ts, unsigned long long* list,
                                                  uint32_t limit, uint32_t* m_io, float* qs, uint32_t* s_extra,
                                                  int* s_nredo, int* s_redo) {
  const unsigned long long* keys = p.sample_keys + (size_t)q * p.sample_m;
  const uint32_t cnt = *m_io;
  const int T = p.topt, slots = p.fused_tiles;
  if (threadIdx.x == 0) {
    *s_extra = 0u;
    *s_nredo = 0;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < slots; s += blockDim.x) {
    unsigned long long kk[kFusedTopT];  // the slot's keys in one round of independent loads (T == kFusedTopT)
#pragma unroll
    for (int i = 0; i < kFusedTopT; ++i) kk[i] = i < T ? __ldcg(keys + (size_t)s * T + i) : 0ull;
    unsigned long long last = kk[0];
#pragma unroll
    for (int i = 1; i < kFusedTopT; ++i)
      if (i < T) last = kk[i];
    if (last != 0ull && key_score(last) >= ts) {
      const int i = atomicAdd(s_nredo, 1);
      if (i < kMaxRedo) s_redo[i] = s;
    } else {
#pragma unroll
      for (int i = 0; i < kFusedTopT; ++i) {
        const unsigned long long key = kk[i];
        if (i < T && key != 0ull && key_score(key) >= ts) {
          const uint32_t pos = cnt + atomicAdd(s_extra, 1u);
          if (pos < limit) list[pos] = key;
        }
      }
    }
  }
  __syncthreads();
  const int nredo = *s_nredo;
  if (nredo > kMaxRedo) return false;  // hand the query to the exact path
  if (nredo > 0) {
    load_query_f32<DT>(p, q, qs);
    __syncthreads();
    // the re-score uses fp32 FMAs, the scan used the tensor core: accept with a margin, the select sorts it out
    const float ts_lo = ts - 1e-4f * (fabsf(ts) + 1e-2f);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int chunks = p.row_bytes >> 4;
    for (int ri = 0; ri < nredo; ++ri) {
      long long row0, end;
      if (p.range_rows > 0) {  // range mode: slot s = first tile of CTA s
        row0 = (long long)s_redo[ri] * p.range_rows;
        end = row0 + p.first_rows;
      } else {
        const long long tile = ((long long)s_redo[ri] * p.perm_mul) % p.perm_n;
        row0 = tile * p.tile_rows;
        end = row0 + p.tile_rows;
      }
      if (end > p.n) end = p.n;
      for (long long base = row0 + warp * 4; base < end; base += (long long)nwarps * 4) {
        const int nvalid = (int)((end - base) < 4 ? (end - base) : 4);
        float sc4[4];
        dot_rows4<DT>(reinterpret_cast<const uint8_t*>(p.X) + (size_t)base * p.row_bytes, (size_t)p.row_bytes, nvalid, qs,
                      chunks, lane, sc4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r >= nvalid) continue;
          float s = sc4[r];
          if (p.x_scale) s *= p.x_scale[base + r];
          if (lane == 0 && s >= ts_lo) {
            const uint32_t pos = cnt + atomicAdd(s_extra, 1u);
            if (pos < limit) list[pos] = make_key(s, (uint32_t)(base + r));
          }
        }
      }
    }
    __syncthreads();
  }
  *m_io = cnt + *s_extra;
  return *m_io <= limit;
}

// One CTA per query.  Launched with programmatic stream serialization right behind the scan: the CTAs are scheduled
// while the scan drains and pass pdl_wait() once it has completed.  The kernel also leaves the workspace header CLEAN
// for the next search (cnt[q] = 0, tau_score[q] = sentinel, grid-barrier counter = 0): no memset launches per step.
template <int DT>
__global__ void __launch_bounds__(kSelectThreads)
    final_select_kernel(const SimParams p, int k, int kpad, long long idx_offset, float* out_score, int32_t* out_idx,
                        uint32_t* ovf) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* dst = reinterpret_cast<uint64_t*>(smem_raw);  // [kpad]
  uint64_t* stage = dst + kpad;                            // [kStageKeys]
  float* qs = reinterpret_cast<float*>(stage + kStageKeys);  // [d] (first-phase redo / exact redo)
  __shared__ union {
    SelectScratch sel;
    BucketScratch bkt;
  } scr;
  __shared__ uint32_t s_extra, s_cnt;
  __shared__ float s_ts;
  __shared__ int s_nredo;
  __shared__ int s_redo[kMaxRedo];
  pdl_wait();
  pdl_launch_dependents();
  const int q = blockIdx.x;
  if (threadIdx.x == 0) {
    s_cnt = (p.mode == kModeScanAll) ? (uint32_t)p.n : __ldcg(&p.cnt[q]);  // scan-all: slot == row
    s_ts = __ldcg(&p.tau_score[q]);
    // self-cleaning workspace header (rir_sim_topk_workspace_init's invariant)
    p.cnt[q] = 0u;
    reinterpret_cast<uint32_t*>(p.tau_score)[q] = kTauUnset;
    if (q == 0 && p.gbar != nullptr) p.gbar[0] = 0u;
    ovf[q] = 0u;
  }
  __syncthreads();
  uint32_t m = s_cnt;
  const float ts = s_ts;
  const unsigned long long* c = p.cand + (size_t)q * p.cap;
  bool ok = m <= (uint32_t)p.cap;
  // everything fits in shared memory (the common case): candidates are read from L2 exactly once
  const uint32_t fp_room = (p.mode == kModeFused) ? (uint32_t)p.sample_m + 512u : 0u;
  const bool staged = ok && p.mode != kModeScanAll && m + fp_room <= (uint32_t)kStageKeys;
  if (staged) {
    for (int i = threadIdx.x; i < (int)m; i += blockDim.x) stage[i] = __ldcg(c + i);
    __syncthreads();
  }
  if (ok && p.mode == kModeFused)
    ok = merge_first_phase<DT>(p, q, ts, staged ? reinterpret_cast<unsigned long long*>(stage) : p.cand + (size_t)q * p.cap,
                               staged ? (uint32_t)kStageKeys : (uint32_t)p.cap, &m, qs, &s_extra, &s_nredo, s_redo);
  if (!ok) {  // candidate list overflowed (adversarial row order): redo this query exactly
    const int bufcap = inline_exact_bufcap(k);
    if (bufcap <= kStageKeys) {  // right here, no extra launch
      __syncthreads();
      exact_scan_body<DT>(p, q, k, bufcap, stage, qs, idx_offset, out_score, out_idx, p.ex.fold ? stage : nullptr, dst);
    } else if (threadIdx.x == 0) {
      ovf[q] = 1u;  // very large k: the separate exact kernel owns this query
    }
    return;
  }
  int got;
  if (staged) {
    got = -1;
    if ((int)m > k && kpad <= kFastSelectMaxK) got = block_select_bucket(stage, (int)m, k, dst, kpad, &scr.bkt);
    if (got < 0) {
      const uint64_t* st = stage;
      auto key_at = [=](int i) -> unsigned long long { return st[i]; };
      got = block_select_topk(key_at, (int)m, k, dst, kpad, &scr.sel);
    }
  } else if (m <= (uint32_t)kStageKeys && (int)m > k) {
    for (int i = threadIdx.x; i < (int)m; i += blockDim.x) stage[i] = c[i];
    __syncthreads();
    const uint64_t* st = stage;
    auto key_at = [=](int i) -> unsigned long long { return st[i]; };
    got = block_select_topk(key_at, (int)m, k, dst, kpad, &scr.sel);
  } else {
    auto key_at = [=](int i) -> unsigned long long { return c[i]; };
    got = block_select_topk(key_at, (int)m, k, dst, kpad, &scr.sel);
  }
  if (p.ex.G > 0) exchange_and_merge(p.ex, q, dst, got, k, idx_offset, stage, dst, out_score, out_idx);
  else write_sorted(dst, got, k, idx_offset, out_score + (size_t)q * k, out_idx + (size_t)q * k);
}

// can the exchange merge run inside the select kernel?  (every select CTA resident, lists fit the staging area)
bool select_can_fold_merge(int nq_total, int G, int k_push) {
  static const int on = getenv("RIR_FOLD_MERGE") ? atoi(getenv("RIR_FOLD_MERGE")) : 1;
  const int kpad = pow2_ceil_int(k_push < 32 ? 32 : k_push);
  return on != 0 && nq_total <= sm_count() && (long long)G * k_push <= kStageKeys && k_push <= kpad &&
         select_handles_overflow(k_push);
}

bool select_handles_overflow(int k) { return inline_exact_bufcap(k) <= kStageKeys; }

int launch_final_select(const SimParams& p, int dtype, int nq_total, int k, long long idx_offset, float* out_score,
                        int32_t* out_idx, uint32_t* ovf, cudaStream_t st) {
  int kk = k < 32 ? 32 : k;
  if (p.ex.G > 0 && p.ex.k_push > kk) kk = p.ex.k_push;  // a shard shorter than k still merges lists of k_push keys
  const int kpad = pow2_ceil_int(kk);
  const size_t smem = (size_t)(kpad + kStageKeys) * sizeof(uint64_t) + (size_t)p.d * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("sim_topk(select): k=%d d=%d needs %zu B of shared memory", k, p.d, smem);
    return RIR_E_ARG;
  }
  const dim3 grid((unsigned)nq_total), block((unsigned)select_threads());
  if (dtype == RIR_BF16) {
    RIR_CUDA_OK(ensure_dyn_smem(final_select_kernel<RIR_BF16>, smem));
    RIR_CUDA_OK(launch_pdl(final_select_kernel<RIR_BF16>, grid, block, smem, st, p, k, kpad, idx_offset, out_score, out_idx, ovf));
  } else if (dtype == RIR_FP8E4M3) {
    RIR_CUDA_OK(ensure_dyn_smem(final_select_kernel<RIR_FP8E4M3>, smem));
    RIR_CUDA_OK(launch_pdl(final_select_kernel<RIR_FP8E4M3>, grid, block, smem, st, p, k, kpad, idx_offset, out_score, out_idx, ovf));
  } else {
    RIR_CUDA_OK(ensure_dyn_smem(final_select_kernel<RIR_F32>, smem));
    RIR_CUDA_OK(launch_pdl(final_select_kernel<RIR_F32>, grid, block, smem, st, p, k, kpad, idx_offset, out_score, out_idx, ovf));
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// ---------------------------------------------------------------------------------------------
// robust exact scan: one CTA per query, running top-k in shared memory.  Used (a) for queries whose candidate
// list overflowed, (b) as RIR_PATH_EXACT — an independent second implementation for the parity tests.
// ---------------------------------------------------------------------------------------------
constexpr int kExactThreads = 256;
constexpr int kExactRowsPerIter = 32;  // 8 warps x 4 rows

template <int DT>
__global__ void __launch_bounds__(kExactThreads)
    exact_scan_kernel(const SimParams p, int k, int bufcap, long long idx_offset, float* out_score, int32_t* out_idx,
                      const uint32_t* ovf) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int q = blockIdx.x;
  if (ovf != nullptr && ovf[q] == 0u) return;
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);           // [bufcap]
  float* qs = reinterpret_cast<float*>(buf + bufcap);              // [d]
  exact_scan_body<DT>(p, q, k, bufcap, buf, qs, idx_offset, out_score, out_idx);
}

int launch_exact_scan(const SimParams& p, int dtype, int nq_total, int k, long long idx_offset, float* out_score,
                      int32_t* out_idx, const uint32_t* ovf, cudaStream_t st) {
  const int kpad = pow2_ceil_int(k < 32 ? 32 : k);
  // room for k survivors + one iteration of pushes; all rows fit when n <= kpad (full ranking)
  int bufcap = (p.n <= (long long)kpad) ? kpad : 2 * kpad;
  if (bufcap < 64) bufcap = 64;
  const size_t smem = (size_t)bufcap * sizeof(uint64_t) + (size_t)p.d * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("sim_topk(exact): k=%d with n=%lld needs %zu B of shared memory", k, p.n, smem);
    return RIR_E_ARG;
  }
  if (dtype == RIR_BF16) {
    RIR_CUDA_OK(ensure_dyn_smem(exact_scan_kernel<RIR_BF16>, smem));
    exact_scan_kernel<RIR_BF16><<<nq_total, kExactThreads, smem, st>>>(p, k, bufcap, idx_offset, out_score, out_idx, ovf);
  } else if (dtype == RIR_FP8E4M3) {
    RIR_CUDA_OK(ensure_dyn_smem(exact_scan_kernel<RIR_FP8E4M3>, smem));
    exact_scan_kernel<RIR_FP8E4M3><<<nq_total, kExactThreads, smem, st>>>(p, k, bufcap, idx_offset, out_score, out_idx, ovf);
  } else if (dtype == RIR_F32) {
    RIR_CUDA_OK(ensure_dyn_smem(exact_scan_kernel<RIR_F32>, smem));
    exact_scan_kernel<RIR_F32><<<nq_total, kExactThreads, smem, st>>>(p, k, bufcap, idx_offset, out_score, out_idx, ovf);
  } else {
    set_error("sim_topk(exact): unsupported dtype %d", dtype);
    return RIR_E_ARG;
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// ---------------------------------------------------------------------------------------------
// rescore: recompute the scores of a candidate list with higher-precision rows (bf16 / fp32) and keep the best k.
// Used after an fp8 scan (k_in > k candidates with margin) so the final list meets the 5e-3 bar against an fp32
// rescore (BASELINE.json north_star; SURVEY.md §7.3 "fp8 quantisation of unit-norm vectors").
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kExactThreads)
    rescore_kernel(const void* Q, const void* X, const float* q_scale, const float* x_scale, long long n_local,
                   long long idx_offset, int d, const int32_t* ix_in, int k_in, int k, int kpad, float* out_sc,
                   int32_t* out_ix) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);       // [kin_pad]
  const int kin_pad = pow2_ceil_int(k_in < 32 ? 32 : k_in);
  uint64_t* dst = keys + kin_pad;                                // [kpad]
  float* qs = reinterpret_cast<float*>(dst + kpad);              // [d]
  __shared__ SelectScratch scr;
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ESZ = DT == RIR_F32 ? 4 : (DT == RIR_BF16 ? 2 : 1);
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    float v;
    if (DT == RIR_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(Q)[(size_t)q * d + i]);
    else v = reinterpret_cast<const float*>(Q)[(size_t)q * d + i];
    if (q_scale) v *= q_scale[q];
    qs[i] = v;
  }
  __syncthreads();
  const int chunks = (d * ESZ) >> 4;
  for (int j = warp; j < k_in; j += kExactThreads / 32) {
    const long long gid = ix_in[(size_t)q * k_in + j];
    const long long loc = gid - idx_offset;
    unsigned long long key = 0ull;
    if (gid >= 0 && loc >= 0 && loc < n_local) {
      float s = dot_row<DT>(reinterpret_cast<const uint8_t*>(X) + (size_t)loc * d * ESZ, qs, chunks, lane);
      if (x_scale) s *= x_scale[loc];
      key = make_key(s, (uint32_t)loc);
    }
    if (lane == 0) keys[j] = key;
  }
  __syncthreads();
  const uint64_t* kk = keys;
  auto key_at = [=](int i) -> unsigned long long { return kk[i]; };
  const int got = block_select_topk(key_at, k_in, k, dst, kpad, &scr);
  write_sorted(dst, got, k, idx_offset, out_sc + (size_t)q * k, out_ix + (size_t)q * k);
}

// ---------------------------------------------------------------------------------------------
// cross-shard merge: [G, nq, k] lists -> global top-k
// ---------------------------------------------------------------------------------------------
constexpr int kMergeRankMax = 1024;  // G*k up to which the lists are ranked by counting in shared memory

// Any order inside the lists (the re-ranking hook feeds unsorted scores): rank by counting, ties (only possible
// between padding-free duplicates a caller passes in) broken by position so the result is always a permutation.
__global__ void __launch_bounds__(kSelectThreads)
    merge_topk_kernel(const float* sc, const int32_t* ix, int G, int nq, int k, int kpad, float* out_sc,
                      int32_t* out_ix) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* dst = reinterpret_cast<uint64_t*>(smem_raw);  // [kpad]
  uint64_t* all = dst + kpad;                              // [G*k] when G*k <= kMergeRankMax
  __shared__ SelectScratch scr;
  const int q = blockIdx.x;
  const int m = G * k;
  auto key_at = [=](int i) -> unsigned long long {
    const int g = i / k, j = i - g * k;
    const size_t o = ((size_t)g * nq + q) * k + j;
    const int32_t id = ix[o];
    return id < 0 ? 0ull : make_key(sc[o], (uint32_t)id);
  };
  if (m <= kMergeRankMax) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) all[i] = key_at(i);
    for (int i = threadIdx.x; i < kpad; i += blockDim.x) dst[i] = 0ull;
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const uint64_t key = all[i];
      if (key == 0ull) continue;
      int r = 0;
      for (int j = 0; j < m; ++j) {
        const uint64_t o = all[j];
        r += (o > key || (o == key && j < i)) ? 1 : 0;
      }
      if (r < k) dst[r] = key;
    }
    __syncthreads();
    write_sorted(dst, k, k, 0, out_sc + (size_t)q * k, out_ix + (size_t)q * k);
    return;
  }
  const int got = block_select_topk(key_at, m, k, dst, kpad, &scr);
  write_sorted(dst, got, k, 0, out_sc + (size_t)q * k, out_ix + (size_t)q * k);
}

// Sharded search, receiving side: wait until every rank has published query q for this epoch, then merge the G
// SORTED lists of k keys out of this rank's own inbox: rank = position in the own list + number of greater keys in
// every other list (binary searches in shared memory) — no sort.  PDL: scheduled while the select kernel drains.
__global__ void __launch_bounds__(kSelectThreads)
    merge_exchange_kernel(const Exchange ex, int k, int kpad, float* out_sc, int32_t* out_ix) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* dst = reinterpret_cast<uint64_t*>(smem_raw);  // [kpad]
  uint64_t* lists = dst + kpad;                            // [G*k] when G*k <= kMergeStage
  __shared__ SelectScratch scr;
  pdl_wait();
  pdl_launch_dependents();
  const int q = blockIdx.x;
  const int b = (int)(ex.epoch & 1u);
  unsigned long long* mine = ex.inbox[ex.rank];
  if ((int)threadIdx.x < ex.G) {
    const uint32_t* f = exchange_flag(mine, ex, b, threadIdx.x, q);
    const long long t0 = clock64();
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == ex.epoch) break;
      __nanosleep(100);
      if (clock64() - t0 > 20000000000ll) {
        printf("librir: rank %d never received query %d of rank %d (epoch %u, saw %u)\n", ex.rank, q, (int)threadIdx.x,
               ex.epoch, v);
        __trap();
      }
    }
  }
  __syncthreads();
  const unsigned long long* base = mine + (size_t)b * ex.G * ex.nq_max * ex.k_max;
  const int nq_max = ex.nq_max, k_max = ex.k_max;
  auto key_at = [=](int i) -> unsigned long long {
    const int g = i / k, j = i - g * k;
    return __ldcg(base + ((size_t)g * nq_max + q) * k_max + j);
  };
  const int m = ex.G * k;
  if (m <= kMergeStage) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) lists[i] = key_at(i);
    __syncthreads();
    block_merge_sorted_lists(lists, ex.G, k, dst);
    write_sorted(dst, k, k, 0, out_sc + (size_t)q * k, out_ix + (size_t)q * k);
    return;
  }
  const int got = block_select_topk(key_at, m, k, dst, kpad, &scr);
  write_sorted(dst, got, k, 0, out_sc + (size_t)q * k, out_ix + (size_t)q * k);
}

int launch_merge_exchange(const Exchange& ex, int nq, int k, float* out_score, int32_t* out_idx, cudaStream_t st,
                          bool side_stream) {
  const int kpad = pow2_ceil_int(k < 32 ? 32 : k);
  const int m = ex.G * k;
  const size_t smem = (size_t)(kpad + (m <= kMergeStage ? m : 0)) * sizeof(uint64_t);
  RIR_CUDA_OK(ensure_dyn_smem(merge_exchange_kernel, smem));
  if (side_stream) {
    // runs next to the following search's scan (one scan CTA per SM leaves room): a small CTA, no PDL edge
    merge_exchange_kernel<<<nq, 256, smem, st>>>(ex, k, kpad, out_score, out_idx);
  } else {
    RIR_CUDA_OK(launch_pdl(merge_exchange_kernel, dim3((unsigned)nq), dim3((unsigned)select_threads()), smem, st, ex, k,
                           kpad, out_score, out_idx));
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

}  // namespace rir

extern "C" int rir_rescore_topk(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale,
                                int nq, int64_t n_local, int64_t idx_offset, int d, const int32_t* ix_in, int k_in,
                                int k, float* out_score, int32_t* out_idx, void* stream) {
  using namespace rir;
  if (int e = check_arch()) return e;
  RIR_REQUIRE(nq >= 0 && d >= 1 && k >= 1 && k_in >= k && k_in <= 8192, "rescore_topk: bad shape nq=%d k_in=%d k=%d", nq,
              k_in, k);
  RIR_REQUIRE(dtype == RIR_BF16 || dtype == RIR_F32, "rescore_topk: rescoring rows must be bf16 or fp32");
  const int esz = dtype == RIR_BF16 ? 2 : 4;
  RIR_REQUIRE(((size_t)d * esz) % 16 == 0, "rescore_topk: row size must be a multiple of 16 bytes");
  if (nq == 0) return RIR_OK;
  RIR_REQUIRE(Q && X && ix_in && out_score && out_idx, "rescore_topk: null pointer");
  const int kpad = pow2_ceil_int(k < 32 ? 32 : k);
  const int kin_pad = pow2_ceil_int(k_in < 32 ? 32 : k_in);
  const size_t smem = (size_t)(kpad + kin_pad) * sizeof(uint64_t) + (size_t)d * sizeof(float);
  RIR_REQUIRE(smem <= 220 * 1024, "rescore_topk: k_in/k/d too large for shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RIR_BF16) {
    RIR_CUDA_OK(ensure_dyn_smem(rescore_kernel<RIR_BF16>, smem));
    rescore_kernel<RIR_BF16><<<nq, kExactThreads, smem, st>>>(Q, X, q_scale, x_scale, n_local, idx_offset, d, ix_in, k_in, k,
                                                             kpad, out_score, out_idx);
  } else {
    RIR_CUDA_OK(ensure_dyn_smem(rescore_kernel<RIR_F32>, smem));
    rescore_kernel<RIR_F32><<<nq, kExactThreads, smem, st>>>(Q, X, q_scale, x_scale, n_local, idx_offset, d, ix_in, k_in, k,
                                                            kpad, out_score, out_idx);
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

extern "C" size_t rir_merge_topk_workspace(int G, int nq, int k) {
  (void)G; (void)nq; (void)k;
  return 0;  // the merge runs entirely in shared memory
}

extern "C" int rir_merge_topk(const float* sc, const int32_t* ix, int G, int nq, int k, float* out_sc, int32_t* out_ix,
                              void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  using namespace rir;
  if (int e = check_arch()) return e;
  RIR_REQUIRE(sc && ix && out_sc && out_ix, "merge_topk: null pointer");
  RIR_REQUIRE(G >= 1 && nq >= 0 && k >= 1 && k <= 16384, "merge_topk: bad shape G=%d nq=%d k=%d", G, nq, k);
  RIR_REQUIRE((long long)G * k < (1ll << 30), "merge_topk: G*k too large");
  if (nq == 0) return RIR_OK;
  const int kpad = pow2_ceil_int(k < 32 ? 32 : k);
  const size_t smem = (size_t)(kpad + ((long long)G * k <= kMergeRankMax ? G * k : 0)) * sizeof(uint64_t);
  RIR_CUDA_OK(ensure_dyn_smem(merge_topk_kernel, smem));
  merge_topk_kernel<<<nq, kSelectThreads, smem, (cudaStream_t)stream>>>(sc, ix, G, nq, k, kpad, out_sc, out_ix);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// rank_positions.cu — positions of the ground-truth ids in the FULL ranking, without sorting the database.
//
// compute_map needs, per query, the position of every positive / junk id in the complete ranked list
// (utils/evaluate.py:76-94: `np.arange(N)[np.in1d(ranks[:, i], ok)]`), which the reference obtains from a full
// np.argsort of the [nq, N] score matrix (iris_evaluate.py:386).  With 1M distractors (R1M) that list is neither
// sorted nor materialised here.  Position of id p in the order "score descending, ties -> lower index":
//     pos(p) = #{ rows x : key(x) > key(p) },      key = (ordered score, ~index)
// — a count, so it shards: every rank counts over its rows and ONE all-reduce adds the counts.
//
//   gnd_score_kernel     score of every ground-truth id on the shard that owns its row (0 elsewhere; all-reduce SUM)
//   rank_threshold_kernel per query: keys of its ids sorted descending (bitonic) -> the thresholds, and the compact
//                         ranked list of the ids themselves
//   rank_count_kernel    one pass over the shard for 8 queries at a time: fp32 dot products (the SAME per-lane FMA
//                         chain and butterfly order as dot_rows.cuh, so a row scores bit-identically to its threshold
//                         and never counts itself), binary search of the row key in the query's thresholds,
//                         shared-memory histogram -> global histogram
//   rank_prefix_kernel   inclusive prefix of the histogram = rows outranking each threshold (this shard)
// The compact lists + positions feed rir_compute_map_at (evaluate_map.cu).
#include "dot_rows.cuh"
#include "topk_select.cuh"

namespace rir {

constexpr int kRQ = 8;        // queries per CTA of the counting pass
constexpr int kRT = 1024;     // thresholds per query and pass
constexpr int kRThreads = 256;

template <int DT>
__global__ void __launch_bounds__(kRThreads)
    gnd_score_kernel(const SimParams p, long long idx_offset, const int32_t* __restrict__ ids,
                     const int32_t* __restrict__ off, float* __restrict__ scores) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);
  const int q = blockIdx.x;
  load_query_f32<DT>(p, q, qs);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = p.row_bytes >> 4;
  for (int i = off[q] + warp; i < off[q + 1]; i += kRThreads / 32) {
    const long long loc = (long long)ids[i] - idx_offset;
    float s = 0.f;
    if (loc >= 0 && loc < p.n) {
      s = dot_row<DT>(reinterpret_cast<const uint8_t*>(p.X) + (size_t)loc * p.row_bytes, qs, chunks, lane);
      if (p.x_scale) s *= p.x_scale[loc];
    }
    if (lane == 0) scores[i] = s;
  }
}

// keys of query q's ids, sorted descending, 0-padded to m_pad (a power of two <= 4096)
__global__ void __launch_bounds__(kRThreads)
    rank_threshold_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids, const int32_t* __restrict__ off,
                          int m_pad, unsigned long long* __restrict__ keys_sorted) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int q = blockIdx.x;
  const int m = off[q + 1] - off[q];
  for (int i = threadIdx.x; i < m_pad; i += blockDim.x)
    keys[i] = i < m ? make_key(scores[off[q] + i], (uint32_t)ids[off[q] + i]) : 0ull;
  block_bitonic_sort_desc(keys, m_pad);
  for (int i = threadIdx.x; i < m_pad; i += blockDim.x) keys_sorted[(size_t)q * m_pad + i] = keys[i];
}

template <int DT>
__global__ void __launch_bounds__(kRThreads, 1)
    rank_count_kernel(const SimParams p, long long idx_offset, const unsigned long long* __restrict__ keys_sorted,
                      int m_pad, int t0, int mt, long long rows_per_cta, uint32_t* __restrict__ hist) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);                                   // [kRQ][d]
  uint64_t* thr = reinterpret_cast<uint64_t*>(qs + (size_t)kRQ * p.d);              // [kRQ][kRT]
  uint32_t* hs = reinterpret_cast<uint32_t*>(thr + (size_t)kRQ * kRT);              // [kRQ][kRT]
  __shared__ int s_m[kRQ];
  const int q0 = blockIdx.y * kRQ;
  const int nqg = (p.nq - q0) < kRQ ? (p.nq - q0) : kRQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // queries of this group -> fp32 shared memory (zero rows for the missing ones)
  for (int qi = 0; qi < kRQ; ++qi) {
    if (qi < nqg) load_query_f32<DT>(p, q0 + qi, qs + (size_t)qi * p.d);
    else for (int i = threadIdx.x; i < p.d; i += blockDim.x) qs[(size_t)qi * p.d + i] = 0.f;
  }
  for (int i = threadIdx.x; i < kRQ * kRT; i += blockDim.x) {
    const int qi = i / kRT, j = i - qi * kRT;
    thr[i] = (qi < nqg && j < mt) ? keys_sorted[(size_t)(q0 + qi) * m_pad + t0 + j] : 0ull;
    hs[i] = 0u;
  }
  __syncthreads();
  if (threadIdx.x < kRQ) {  // real thresholds of this chunk (keys are sorted: zeros only at the end)
    const uint64_t* t = thr + (size_t)threadIdx.x * kRT;
    int lo = 0, hi = mt;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (t[mid] != 0ull) lo = mid + 1; else hi = mid;
    }
    s_m[threadIdx.x] = lo;
  }
  __syncthreads();
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  long long r_end = r_begin + rows_per_cta;
  if (r_end > p.n) r_end = p.n;
  const int chunks = p.row_bytes >> 4;
  const int my_r = lane >> 3, my_q = lane & 7;        // after the reduction lane l holds (row l / 8, query l % 8)
  const int my_m = s_m[my_q];
  const uint64_t* my_thr = thr + (size_t)my_q * kRT;
  const uint64_t my_min = my_m > 0 ? my_thr[my_m - 1] : ~0ull;
  for (long long base = r_begin + warp * 4; base < r_end; base += (kRThreads / 32) * 4) {
    const int nvalid = (int)((r_end - base) < 4 ? (r_end - base) : 4);
    const uint8_t* row0 = reinterpret_cast<const uint8_t*>(p.X) + (size_t)base * p.row_bytes;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    for (int c = lane; c < chunks; c += 32) {
      uint4 x[4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        x[r] = r < nvalid ? ldg_stream_16B(row0 + (size_t)r * p.row_bytes + (size_t)c * 16) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int qi = 0; qi < kRQ; ++qi)
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r * 8 + qi] = chunk_fma<DT>(x[r], qs + (size_t)qi * p.d, c, v[r * 8 + qi]);
    }
    // transposing butterfly: the same lane pairs are added at the same levels (16, 8, 4, 2, 1) as warp_sum does, so
    // every total is bit-identical to dot_row's — but 31 shuffles instead of 160, and lane l ends with total l
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const float send = (lane & h) ? v[i] : v[i + h];
        const float keep = (lane & h) ? v[i + h] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
      }
    }
    const long long row = base + my_r;
    if (my_r < nvalid && my_q < nqg && my_m > 0) {
      float s = v[0];
      if (p.x_scale) s *= p.x_scale[row];
      const uint64_t key = make_key(s, (uint32_t)(row + idx_offset));
      if (key > my_min) {  // outranks at least the last threshold
        int lo = 0, hi = my_m;  // first threshold strictly below the row's key
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (my_thr[mid] >= key) lo = mid + 1; else hi = mid;
        }
        atomicAdd(&hs[my_q * kRT + lo], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRQ * kRT; i += blockDim.x) {
    const int qi = i / kRT, j = i - qi * kRT;
    if (qi < nqg && j < mt && hs[i] != 0u) atomicAdd(&hist[(size_t)(q0 + qi) * m_pad + t0 + j], hs[i]);
  }
}

// counts[q][i] = sum_{t0(i) <= j <= i} hist[q][j]   (the prefix restarts at every kRT-threshold pass)
__global__ void rank_prefix_kernel(const uint32_t* __restrict__ hist, int nq, int m_pad, int32_t* __restrict__ counts) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int passes = (m_pad + kRT - 1) / kRT;
  if (t >= nq * passes) return;
  const int q = t / passes, ps = t - q * passes;
  const int lo = ps * kRT, hi = (lo + kRT < m_pad) ? lo + kRT : m_pad;
  uint32_t run = 0;
  for (int j = lo; j < hi; ++j) {
    run += hist[(size_t)q * m_pad + j];
    counts[(size_t)q * m_pad + j] = (int32_t)run;
  }
}

template <int DT>
static int rank_count_dt(const SimParams& p, long long idx_offset, const unsigned long long* keys_sorted, int m_pad,
                         uint32_t* hist, cudaStream_t st) {
  const size_t smem = (size_t)kRQ * p.d * sizeof(float) + (size_t)kRQ * kRT * (sizeof(uint64_t) + sizeof(uint32_t));
  if (smem > 220 * 1024) {
    set_error("rank_count: d=%d needs %zu B of shared memory", p.d, smem);
    return RIR_E_ARG;
  }
  RIR_CUDA_OK(ensure_dyn_smem(rank_count_kernel<DT>, smem));
  const int groups = (p.nq + kRQ - 1) / kRQ;
  long long gx = (2ll * sm_count() + groups - 1) / groups;
  if (gx < 1) gx = 1;
  long long rows_per_cta = (p.n + gx - 1) / gx;
  rows_per_cta = (rows_per_cta + 31) / 32 * 32;
  gx = (p.n + rows_per_cta - 1) / rows_per_cta;
  for (int t0 = 0; t0 < m_pad; t0 += kRT) {
    const int mt = (m_pad - t0) < kRT ? (m_pad - t0) : kRT;
    rank_count_kernel<DT><<<dim3((unsigned)gx, (unsigned)groups), kRThreads, smem, st>>>(p, idx_offset, keys_sorted, m_pad, t0,
                                                                                         mt, rows_per_cta, hist);
    RIR_LAUNCH_OK();
  }
  return RIR_OK;
}

}  // namespace rir

using namespace rir;

static int rp_elem(int dtype) { return dtype == RIR_BF16 ? 2 : (dtype == RIR_FP8E4M3 ? 1 : (dtype == RIR_F32 ? 4 : 0)); }

extern "C" int rir_gnd_scores(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                              int64_t n_local, int d, int64_t idx_offset, const int32_t* ids, const int32_t* off,
                              int64_t total_ids, float* scores, void* stream) {
  if (int e = check_arch()) return e;
  const int esz = rp_elem(dtype);
  RIR_REQUIRE(esz != 0 && nq >= 0 && n_local >= 1 && d >= 1, "gnd_scores: bad arguments");
  RIR_REQUIRE(((size_t)d * esz) % 16 == 0, "gnd_scores: row size must be a multiple of 16 bytes");
  if (nq == 0 || total_ids == 0) return RIR_OK;
  RIR_REQUIRE(Q && X && ids && off && scores, "gnd_scores: null pointer");
  SimParams p{};
  p.Q = Q; p.X = X; p.q_scale = q_scale; p.x_scale = x_scale; p.nq = nq; p.n = n_local; p.d = d; p.row_bytes = d * esz;
  const size_t smem = (size_t)d * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RIR_BF16) gnd_score_kernel<RIR_BF16><<<nq, kRThreads, smem, st>>>(p, idx_offset, ids, off, scores);
  else if (dtype == RIR_FP8E4M3) gnd_score_kernel<RIR_FP8E4M3><<<nq, kRThreads, smem, st>>>(p, idx_offset, ids, off, scores);
  else gnd_score_kernel<RIR_F32><<<nq, kRThreads, smem, st>>>(p, idx_offset, ids, off, scores);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

extern "C" int rir_rank_thresholds(const float* scores, const int32_t* ids, const int32_t* off, int nq, int m_pad,
                                   uint64_t* keys_sorted, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(nq >= 0 && m_pad >= 32 && m_pad <= 4096 && (m_pad & (m_pad - 1)) == 0,
              "rank_thresholds: m_pad must be a power of two in [32, 4096] (got %d)", m_pad);
  if (nq == 0) return RIR_OK;
  RIR_REQUIRE(scores && ids && off && keys_sorted, "rank_thresholds: null pointer");
  const size_t smem = (size_t)m_pad * sizeof(uint64_t);
  rank_threshold_kernel<<<nq, kRThreads, smem, (cudaStream_t)stream>>>(
      scores, ids, off, m_pad, reinterpret_cast<unsigned long long*>(keys_sorted));
  RIR_LAUNCH_OK();
  return RIR_OK;
}

extern "C" int rir_rank_count(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                              int64_t n_local, int d, int64_t idx_offset, const uint64_t* keys_sorted, int m_pad,
                              int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_arch()) return e;
  const int esz = rp_elem(dtype);
  RIR_REQUIRE(esz != 0 && nq >= 0 && n_local >= 1 && d >= 1, "rank_count: bad arguments");
  RIR_REQUIRE(((size_t)d * esz) % 16 == 0, "rank_count: row size must be a multiple of 16 bytes");
  RIR_REQUIRE(m_pad >= 32 && m_pad <= 4096 && (m_pad & (m_pad - 1)) == 0, "rank_count: bad m_pad %d", m_pad);
  RIR_REQUIRE(n_local + idx_offset < (1ll << 31) && idx_offset >= 0, "rank_count: global row index exceeds int32");
  if (nq == 0) return RIR_OK;
  RIR_REQUIRE(Q && X && keys_sorted && counts && workspace, "rank_count: null pointer");
  RIR_REQUIRE(workspace_bytes >= (size_t)nq * m_pad * sizeof(uint32_t), "rank_count: workspace too small");
  SimParams p{};
  p.Q = Q; p.X = X; p.q_scale = q_scale; p.x_scale = x_scale; p.nq = nq; p.n = n_local; p.d = d; p.row_bytes = d * esz;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* hist = reinterpret_cast<uint32_t*>(workspace);
  RIR_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)nq * m_pad * sizeof(uint32_t), st));
  const unsigned long long* ks = reinterpret_cast<const unsigned long long*>(keys_sorted);
  int rc;
  if (dtype == RIR_BF16) rc = rank_count_dt<RIR_BF16>(p, idx_offset, ks, m_pad, hist, st);
  else if (dtype == RIR_FP8E4M3) rc = rank_count_dt<RIR_FP8E4M3>(p, idx_offset, ks, m_pad, hist, st);
  else rc = rank_count_dt<RIR_F32>(p, idx_offset, ks, m_pad, hist, st);
  if (rc) return rc;
  const int passes = (m_pad + kRT - 1) / kRT;
  rank_prefix_kernel<<<(nq * passes + 127) / 128, 128, 0, st>>>(hist, nq, m_pad, counts);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

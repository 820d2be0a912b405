// dense_mma.cu — fp32-accurate dense contractions on the bf16 tensor cores, and the fused descriptor head.
//
//   rir_gem_l2_whiten   pool -> [L2] -> whiten (+bias) -> [L2] as ONE call: the tail of GeM.forward_test
//                       (networks/RetrievalNet.py:337-344) and SOLAR.forward_test (:583-590)
//   rir_whiten_prepare  one-time split of the whitening weights (networks/spca.py:215-227 output) for the call above
//   rir_whiten_prepared whitening of already pooled descriptors with prepared weights (networks/RetrievalNet.py:342,588)
//   syrk_split_bf16     centred covariance X_c^T X_c of pcawhitenlearn_shrinkage (networks/backbone.py:47-50), called
//                       from pca_whiten.cu
//
// The contraction itself (dense_nt_kernel): out[m, n] = sum_k A[m, k] B[n, k] with both operands K-major.  The
// reference computes it in fp32 (cuDNN 1x1 conv / np.dot); a plain bf16 or tf32 tensor-core product would miss the
// 1e-5 parity bar, so every fp32 operand is split exactly into a bf16 pair v = v1 + v2 (+ a remainder below 2^-18 |v|)
// and the product is taken as a1 b1 + a1 b2 + a2 b1 — three bf16 tcgen05.mma passes over K accumulated in ONE fp32
// TMEM accumulator (the dropped terms are below 1.2e-5 relative per product and average out far lower).
//
// Kernel shape: one CTA per (128-row tile of A, 256-row tile of B, K split).  Warp 0 = TMA producer (4-stage ring of
// 16 KB + 32 KB, SWIZZLE_128B boxes), warp 1 = TMEM allocator + the single MMA-issuing thread (tcgen05.mma
// kind::f16, M=128, N=256, K=16), warps 2-5 = epilogue (tcgen05.ld -> fp32 partial tile).  Small batches are
// latency / ingest bound — each SM has to pull (128 + 256) x K x 2 B through L2 — so K is split across CTAs until the
// grid fills the 148 SMs; the partial tiles are summed in a fixed order by the finishing kernel (deterministic, no
// atomics), which also applies bias / input norm / output L2 norm.
#include <cuda.h>
#include "rir_common.cuh"
#include "tensor_map.cuh"

namespace rir {

constexpr int kDM = 128;        // A rows per tile (TMEM lanes)
constexpr int kDN = 256;        // B rows per tile (TMEM columns)
constexpr int kDStages = 4;
constexpr int kDABytes = kDM * 128;
constexpr int kDBBytes = kDN * 128;
constexpr int kDStageBytes = kDABytes + kDBBytes;
constexpr int kDThreads = 192;

struct DenseTail {
  uint64_t full[kDStages], empty[kDStages];
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad;
};

struct DenseGeom {
  int kchunks;      // 64-element chunks per K segment
  int cps;          // chunks per split (over the 3 * kchunks chunks of the three passes)
  int mp, np;       // padded partial-tile dimensions: mt * 128, nt * 256
  int upper_only;   // syrk: skip tiles that hold no element with column >= row
};

__global__ void __launch_bounds__(kDThreads, 1)
    dense_nt_kernel(const DenseGeom g, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                    float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
  if (g.upper_only && n_tile * kDN + kDN - 1 < m_tile * kDM) return;  // whole tile below the diagonal (uniform)
  DenseTail* tail = reinterpret_cast<DenseTail*>(smem + (size_t)kDStages * kDStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = 3 * g.kchunks;
  const int c_begin = split * g.cps;
  int c_end = c_begin + g.cps;
  if (c_end > total) c_end = total;
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) {
      printf("librir: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    for (int s = 0; s < kDStages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    mbar_init(&tail->acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmB2);
  }
  if (warp == 1) {
    tmem_alloc(&tail->tmem_base, kDN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;
  pdl_wait();               // the operands come from the previous kernel of the stream (pool / split)
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_last();  // operands are re-read by other tiles: keep them in L2
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int seg = c / g.kchunks, kc = c - seg * g.kchunks;
        mbar_wait(&tail->empty[s], ph ^ 1u);
        mbar_expect_tx(&tail->full[s], (uint32_t)kDStageBytes);
        uint8_t* dst = smem + (size_t)s * kDStageBytes;
        tma_tensor2d_g2s(dst, seg == 2 ? &tmA2 : &tmA1, kc * 64, m_tile * kDM, &tail->full[s], pol);
        tma_tensor2d_g2s(dst + kDABytes, seg == 1 ? &tmB2 : &tmB1, kc * 64, n_tile * kDN, &tail->full[s], pol);
        if (++s == kDStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = bf16, K-major both, N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kDN >> 3) << 17) | ((uint32_t)(kDM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&tail->full[s], ph);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + (size_t)s * kDStageBytes));
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + (size_t)s * kDStageBytes + kDABytes));
#pragma unroll
        for (int j = 0; j < 4; ++j)  // 4 x K=16 inside the 128-byte swizzle atom: +32 bytes = +2 in the address field
          umma_f16(tmem_base, a_desc + 2u * j, b_desc + 2u * j, idesc, (uint32_t)((c > c_begin) || j != 0));
        umma_commit(&tail->empty[s]);
        if (++s == kDStages) { s = 0; ph ^= 1u; }
      }
      umma_commit(&tail->acc_full);
    }
  } else {
    // epilogue: warp w may read TMEM lanes [32 * (w % 4), +32); thread == one row of the tile
    const int ew = warp & 3;
    mbar_wait(&tail->acc_full, 0u);
    tc_fence_after();
    const int row = m_tile * kDM + ew * 32 + lane;
    float* out = partial + ((size_t)split * g.mp + row) * g.np + (size_t)n_tile * kDN;
    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
    uint32_t va[16], vb[16];
    tmem_ld_32x32_x16(taddr, va);
#pragma unroll 1
    for (int c0 = 0; c0 < kDN; c0 += 32) {
      tmem_ld_wait();
      tmem_ld_32x32_x16(taddr + (uint32_t)(c0 + 16), vb);
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(out + c0 + j) = make_float4(__uint_as_float(va[j]), __uint_as_float(va[j + 1]),
                                                               __uint_as_float(va[j + 2]), __uint_as_float(va[j + 3]));
      tmem_ld_wait();
      if (c0 + 32 < kDN) tmem_ld_32x32_x16(taddr + (uint32_t)(c0 + 32), va);
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(out + c0 + 16 + j) = make_float4(__uint_as_float(vb[j]), __uint_as_float(vb[j + 1]),
                                                                    __uint_as_float(vb[j + 2]), __uint_as_float(vb[j + 3]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kDN);
  }
}

// ---------------------------------------------------------------------------------------------
// planning: tiles, K split, workspace
// ---------------------------------------------------------------------------------------------
struct DensePlan {
  int mt, nt, S;
  DenseGeom g;
  size_t partial_bytes;
};

// The tensor core accumulates in fp32 but aligns addends by truncation: a long chain of same-sign products (the
// diagonal of a covariance over 20,000 rows: 3 x 313 chunks x 4 MMAs) drifts by ~1e-4 relative (measured).  K is
// therefore split so that no accumulation chain exceeds kMaxChainChunks chunks; the partial tiles are added by the
// finishing kernels in IEEE fp32, in a fixed order.
constexpr int kMaxChainChunks = 64;

static DensePlan dense_plan(long long M, long long N, long long K, bool upper_only) {
  DensePlan pl;
  pl.mt = (int)((M + kDM - 1) / kDM);
  pl.nt = (int)((N + kDN - 1) / kDN);
  pl.g.kchunks = (int)((K + 63) / 64);
  pl.g.mp = pl.mt * kDM;
  pl.g.np = pl.nt * kDN;
  pl.g.upper_only = upper_only ? 1 : 0;
  long long tiles = 0;
  for (int m = 0; m < pl.mt; ++m)
    for (int n = 0; n < pl.nt; ++n)
      if (!upper_only || n * kDN + kDN - 1 >= m * kDM) ++tiles;
  const int total = 3 * pl.g.kchunks;
  long long S = sm_count() / (tiles > 0 ? tiles : 1);
  if (S < 1) S = 1;
  if (S > 32) S = 32;  // (filling the SMs needs no more)
  const long long s_chain = (total + kMaxChainChunks - 1) / kMaxChainChunks;
  if (S < s_chain) S = s_chain;
  if (S > total) S = total;
  pl.g.cps = (int)((total + S - 1) / S);
  pl.S = (total + pl.g.cps - 1) / pl.g.cps;  // every split non-empty
  pl.partial_bytes = (size_t)pl.S * pl.g.mp * pl.g.np * sizeof(float);
  return pl;
}

// A12 = [2][M][Kp] bf16, B12 = [2][N][Kp] bf16 (Kp % 8 == 0, zero padded); partial = [S][mp][np] fp32
static int launch_dense_nt(const DensePlan& pl, const void* A12, long long M, const void* B12, long long N, int Kp,
                           float* partial, cudaStream_t st) {
  CUtensorMap a1, a2, b1, b2;
  const uint8_t* A = reinterpret_cast<const uint8_t*>(A12);
  const uint8_t* B = reinterpret_cast<const uint8_t*>(B12);
  if (int e = make_rowmajor_map(&a1, A, M, Kp, RIR_BF16, kDM)) return e;
  if (int e = make_rowmajor_map(&a2, A + (size_t)M * Kp * 2, M, Kp, RIR_BF16, kDM)) return e;
  if (int e = make_rowmajor_map(&b1, B, N, Kp, RIR_BF16, kDN)) return e;
  if (int e = make_rowmajor_map(&b2, B + (size_t)N * Kp * 2, N, Kp, RIR_BF16, kDN)) return e;
  const size_t smem = (size_t)kDStages * kDStageBytes + sizeof(DenseTail);
  RIR_CUDA_OK(ensure_dyn_smem(dense_nt_kernel, smem));
  RIR_CUDA_OK(launch_pdl(dense_nt_kernel, dim3((unsigned)pl.mt, (unsigned)pl.nt, (unsigned)pl.S), dim3(kDThreads), smem, st,
                         pl.g, a1, a2, b1, b2, partial));
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// ---------------------------------------------------------------------------------------------
// exact bf16 pair of an fp32 value
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *hi = h;
  *lo = __float2bfloat16_rn(v - __bfloat162float(h));  // the subtraction is exact in fp32
}

// rows [R, K] fp32 (row stride ld) -> out1 / out2 [R, Kp] bf16, zero padded
__global__ void __launch_bounds__(256)
    split_rows_kernel(const float* __restrict__ x, long long R, int K, long long ld, int Kp, __nv_bfloat16* __restrict__ o1,
                      __nv_bfloat16* __restrict__ o2) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = R * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Kp;
    const int c = (int)(i - r * Kp);
    __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
    if (c < K) split_bf16(x[r * ld + c], &h, &l);
    o1[i] = h;
    o2[i] = l;
  }
}

// X [N, D] fp32 -> centred, transposed bf16 pair T1 / T2 [D, Np] (row i = column i of X - mean[i]; Np % 8 == 0)
__global__ void __launch_bounds__(256)
    split_transpose_center_kernel(const float* __restrict__ X, const float* __restrict__ mean, long long N, int D, long long Np,
                                  __nv_bfloat16* __restrict__ t1, __nv_bfloat16* __restrict__ t2) {
  __shared__ float tile[32][33];
  const long long n0 = (long long)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const long long n = n0 + r;
    const int dd = d0 + tx;
    tile[r][tx] = (n < N && dd < D) ? X[(size_t)n * D + dd] - mean[dd] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int dd = d0 + r;
    const long long n = n0 + tx;
    if (dd < D && n < Np) {
      __nv_bfloat16 h, l;
      split_bf16(tile[tx][r], &h, &l);
      t1[(size_t)dd * Np + n] = h;
      t2[(size_t)dd * Np + n] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// finishing kernels
// ---------------------------------------------------------------------------------------------
// one CTA per output row: out[b, :] = L2?( (sum_s partial[s][b][:]) * inv_in + bias ),  inv_in = 1 / max(||pooled[b]||, eps)
__global__ void __launch_bounds__(256)
    whiten_finish_kernel(const float* __restrict__ partial, int S, int mp, int np, const float* __restrict__ bias,
                         const float* __restrict__ pooled, int C, int d_out, int l2_before, int l2_after,
                         float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float s_val;
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];  // fixed order
      s_val = t;
    }
    __syncthreads();
    return s_val;
  };
  float inv_in = 1.f;
  if (l2_before) {
    float ss = 0.f;
    for (int c = threadIdx.x; c < C; c += 256) {
      const float v = pooled[(size_t)b * C + c];
      ss = fmaf(v, v, ss);
    }
    inv_in = 1.f / fmaxf(sqrtf(block_sum(ss)), 1e-12f);
  }
  float ss = 0.f;
  float* o = out + (size_t)b * d_out;
  // 128-bit loads, four splits in flight per thread (the per-element chain over S is latency-bound otherwise: this
  // kernel took 26 us for 256 x 2048 outputs with scalar loads, more than the contraction itself).  The sum over the
  // splits keeps its fixed order s = 0, 1, 2, ...
  const float* prow = partial + (size_t)b * np;
  const size_t sstride = (size_t)mp * np;
  const bool vec = (d_out & 3) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0 &&
                   (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  if (vec) {
    for (int j = threadIdx.x * 4; j < d_out; j += 256 * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = 0;
      for (; s + 4 <= S; s += 4) {
        const float4 a0 = __ldcg(reinterpret_cast<const float4*>(prow + (size_t)(s + 0) * sstride + j));
        const float4 a1 = __ldcg(reinterpret_cast<const float4*>(prow + (size_t)(s + 1) * sstride + j));
        const float4 a2 = __ldcg(reinterpret_cast<const float4*>(prow + (size_t)(s + 2) * sstride + j));
        const float4 a3 = __ldcg(reinterpret_cast<const float4*>(prow + (size_t)(s + 3) * sstride + j));
        acc.x = ((acc.x + a0.x) + a1.x) + a2.x + a3.x; acc.y = ((acc.y + a0.y) + a1.y) + a2.y + a3.y;
        acc.z = ((acc.z + a0.z) + a1.z) + a2.z + a3.z; acc.w = ((acc.w + a0.w) + a1.w) + a2.w + a3.w;
      }
      for (; s < S; ++s) {
        const float4 a0 = __ldcg(reinterpret_cast<const float4*>(prow + (size_t)s * sstride + j));
        acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
      }
      const float4 bb = bias ? *reinterpret_cast<const float4*>(bias + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v;
      v.x = fmaf(acc.x, inv_in, bb.x); v.y = fmaf(acc.y, inv_in, bb.y);
      v.z = fmaf(acc.z, inv_in, bb.z); v.w = fmaf(acc.w, inv_in, bb.w);
      *reinterpret_cast<float4*>(o + j) = v;
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
  } else {
    for (int j = threadIdx.x; j < d_out; j += 256) {
      float v = 0.f;
      for (int s = 0; s < S; ++s) v += __ldcg(prow + (size_t)s * sstride + j);  // fixed order
      v = fmaf(v, inv_in, bias ? bias[j] : 0.f);
      o[j] = v;
      ss = fmaf(v, v, ss);
    }
  }
  if (l2_after) {
    const float denom = fmaxf(sqrtf(block_sum(ss)), 1e-12f);
    if (vec) {
      for (int j = threadIdx.x * 4; j < d_out; j += 256 * 4) {  // own elements
        float4 v = *reinterpret_cast<float4*>(o + j);
        v.x /= denom; v.y /= denom; v.z /= denom; v.w /= denom;
        *reinterpret_cast<float4*>(o + j) = v;
      }
    } else {
      for (int j = threadIdx.x; j < d_out; j += 256) o[j] = o[j] / denom;  // own elements
    }
  }
}

// cov[i][j] = cov[j][i] = (sum_s partial[s][i][j]) / N for j >= i
__global__ void __launch_bounds__(256)
    syrk_finish_kernel(const float* __restrict__ partial, int S, int mp, int np, int D, float inv_n, float* __restrict__ cov) {
  pdl_wait();
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= D || j < i) return;
  float v = 0.f;
  for (int s = 0; s < S; ++s) v += __ldcg(partial + ((size_t)s * mp + i) * np + j);
  v *= inv_n;
  cov[(size_t)i * D + j] = v;
  cov[(size_t)j * D + i] = v;
}

// ---------------------------------------------------------------------------------------------
// host: whitening with prepared weights
// ---------------------------------------------------------------------------------------------
static int pad8(int v) { return (v + 7) / 8 * 8; }
static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct HeadPlan {
  int Cp;
  DensePlan dp;
  size_t off_pooled, off_x12, off_partial, total;
};

static HeadPlan head_plan(int B, int C, int d_out) {
  HeadPlan hp;
  hp.Cp = pad8(C);
  hp.dp = dense_plan(B, d_out, hp.Cp, false);
  size_t o = 0;
  hp.off_pooled = o;  o = align256(o + (size_t)B * C * sizeof(float));
  hp.off_x12 = o;     o = align256(o + (size_t)2 * B * hp.Cp * sizeof(__nv_bfloat16));
  hp.off_partial = o; o = align256(o + hp.dp.partial_bytes);
  hp.total = o;
  return hp;
}

// x12 already holds the split pooled descriptors, pooled the fp32 ones
static int whiten_from_split(const HeadPlan& hp, uint8_t* ws, const void* W12, const float* bias, int B, int C, int d_out,
                             int l2_before, int l2_after, float* out, cudaStream_t st) {
  float* partial = reinterpret_cast<float*>(ws + hp.off_partial);
  if (int e = launch_dense_nt(hp.dp, ws + hp.off_x12, B, W12, d_out, hp.Cp, partial, st)) return e;
  RIR_CUDA_OK(launch_pdl(whiten_finish_kernel, dim3((unsigned)B), dim3(256), 0, st, (const float*)partial, hp.dp.S,
                         hp.dp.g.mp, hp.dp.g.np, bias, (const float*)(ws + hp.off_pooled), C, d_out, l2_before, l2_after,
                         out));
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// used by pca_whiten.cu: cov = Xc^T Xc / N through the split-bf16 tensor-core contraction
size_t syrk_split_workspace(long long N, int D) {
  const long long Np = (N + 7) / 8 * 8;
  const DensePlan dp = dense_plan(D, D, Np, true);
  return align256((size_t)2 * D * Np * sizeof(__nv_bfloat16)) + align256(dp.partial_bytes);
}

int syrk_split_bf16(const float* X, const float* mean, long long N, int D, float* cov, void* workspace, cudaStream_t st) {
  const long long Np = (N + 7) / 8 * 8;
  if (Np >= (1ll << 31)) {
    set_error("pca_covariance: too many descriptors (%lld)", N);
    return RIR_E_ARG;
  }
  const DensePlan dp = dense_plan(D, D, Np, true);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* t1 = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* t2 = t1 + (size_t)D * Np;
  float* partial = reinterpret_cast<float*>(ws + align256((size_t)2 * D * Np * sizeof(__nv_bfloat16)));
  const dim3 gt((unsigned)((Np + 31) / 32), (unsigned)((D + 31) / 32));
  split_transpose_center_kernel<<<gt, 256, 0, st>>>(X, mean, N, D, Np, t1, t2);
  RIR_LAUNCH_OK();
  if (int e = launch_dense_nt(dp, t1, D, t1, D, (int)Np, partial, st)) return e;
  const dim3 gf((unsigned)((D + 255) / 256), (unsigned)D);
  RIR_CUDA_OK(launch_pdl(syrk_finish_kernel, gf, dim3(256), 0, st, (const float*)partial, dp.S, dp.g.mp, dp.g.np, D,
                         1.0f / (float)N, cov));
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// defined in descriptor_build.cu: pooling with the split-bf16 epilogue
int launch_pool_split(const void* x, int dtype, int B, int C, int HW, int mode, float p, float eps, float alpha, float beta,
                      float* pooled, __nv_bfloat16* s1, __nv_bfloat16* s2, int ld_split, cudaStream_t st);
int launch_l2_rows(const float* x, long long n, int d, float eps, float* out, cudaStream_t st);

}  // namespace rir

using namespace rir;

extern "C" size_t rir_whiten_prepared_bytes(int d_out, int C) {
  if (d_out < 1 || C < 1) return 0;
  return (size_t)2 * d_out * pad8(C) * sizeof(__nv_bfloat16);
}

extern "C" int rir_whiten_prepare(const float* W, int d_out, int C, void* W12, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(W && W12, "whiten_prepare: null pointer");
  RIR_REQUIRE(d_out >= 1 && C >= 1, "whiten_prepare: bad shape d_out=%d C=%d", d_out, C);
  RIR_REQUIRE((reinterpret_cast<uintptr_t>(W12) & 15) == 0, "whiten_prepare: W12 must be 16-byte aligned");
  const int Cp = pad8(C);
  __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(W12);
  long long blocks = ((long long)d_out * Cp + 255) / 256;
  if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
  split_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, d_out, C, C, Cp, o1, o1 + (size_t)d_out * Cp);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

extern "C" size_t rir_gem_l2_whiten_workspace(int B, int C, int d_out) {
  if (B < 1 || C < 1) return 0;
  if (d_out < 1) return align256((size_t)B * C * sizeof(float));
  return head_plan(B, C, d_out).total;
}

extern "C" int rir_whiten_prepared(const float* x, const void* W12, const float* bias, int B, int C, int d_out, int l2_before,
                                   int l2_after, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(B >= 0 && C >= 1 && d_out >= 1, "whiten_prepared: bad shape B=%d C=%d d_out=%d", B, C, d_out);
  if (B == 0) return RIR_OK;
  RIR_REQUIRE(x && W12 && out && workspace, "whiten_prepared: null pointer");
  RIR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(W12) & 15) == 0,
              "whiten_prepared: workspace must be 256-byte aligned, W12 16-byte aligned");
  const HeadPlan hp = head_plan(B, C, d_out);
  if (workspace_bytes < hp.total) {
    set_error("whiten_prepared: workspace of %zu B is smaller than the required %zu B", workspace_bytes, hp.total);
    return RIR_E_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* s1 = reinterpret_cast<__nv_bfloat16*>(ws + hp.off_x12);
  long long blocks = ((long long)B * hp.Cp + 255) / 256;
  if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
  split_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, B, C, C, hp.Cp, s1, s1 + (size_t)B * hp.Cp);
  RIR_LAUNCH_OK();
  if (l2_before)  // the finishing kernel derives the input norm from the fp32 rows
    RIR_CUDA_OK(cudaMemcpyAsync(ws + hp.off_pooled, x, (size_t)B * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return whiten_from_split(hp, ws, W12, bias, B, C, d_out, l2_before, l2_after, out, st);
}

extern "C" int rir_gem_l2_whiten(const void* x, int dtype, int B, int C, int HW, int mode, float p, float eps, float alpha,
                                 float beta, const void* W12, const float* bias, int d_out, int l2_before, int l2_after,
                                 float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(B >= 0 && C >= 1 && HW >= 1, "gem_l2_whiten: bad shape B=%d C=%d HW=%d", B, C, HW);
  if (B == 0) return RIR_OK;
  RIR_REQUIRE(x && out && workspace, "gem_l2_whiten: null pointer");
  RIR_REQUIRE(mode == RIR_POOL_GEM || mode == RIR_POOL_MAX || mode == RIR_POOL_AVG, "gem_l2_whiten: bad mode %d", mode);
  RIR_REQUIRE(dtype == RIR_F32 || dtype == RIR_BF16, "gem_l2_whiten: feature maps must be f32 or bf16 (got %d)", dtype);
  RIR_REQUIRE(mode != RIR_POOL_GEM || p > 0.f, "gem_l2_whiten: GeM exponent must be > 0 (got %g)", (double)p);
  RIR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "gem_l2_whiten: feature maps must be 16-byte aligned, the workspace 256-byte aligned");
  RIR_REQUIRE(W12 == nullptr || d_out >= 1, "gem_l2_whiten: d_out must be >= 1 with whitening weights");
  const size_t need = rir_gem_l2_whiten_workspace(B, C, W12 ? d_out : 0);
  if (workspace_bytes < need) {
    set_error("gem_l2_whiten: workspace of %zu B is smaller than the required %zu B", workspace_bytes, need);
    return RIR_E_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  if (W12 == nullptr) {  // pool (+ L2) only: models/gem_pooling.py:86-92
    float* pooled = l2_after ? reinterpret_cast<float*>(ws) : out;
    if (int e = launch_pool_split(x, dtype, B, C, HW, mode, p, eps, alpha, beta, pooled, nullptr, nullptr, 0, st)) return e;
    if (l2_after) return launch_l2_rows(pooled, B, C, 1e-12f, out, st);
    return RIR_OK;
  }
  RIR_REQUIRE((reinterpret_cast<uintptr_t>(W12) & 15) == 0, "gem_l2_whiten: W12 must be 16-byte aligned");
  const HeadPlan hp = head_plan(B, C, d_out);
  __nv_bfloat16* s1 = reinterpret_cast<__nv_bfloat16*>(ws + hp.off_x12);
  if (hp.Cp != C) RIR_CUDA_OK(cudaMemsetAsync(s1, 0, (size_t)2 * B * hp.Cp * sizeof(__nv_bfloat16), st));  // K padding
  if (int e = launch_pool_split(x, dtype, B, C, HW, mode, p, eps, alpha, beta, reinterpret_cast<float*>(ws + hp.off_pooled),
                                s1, s1 + (size_t)B * hp.Cp, hp.Cp, st))
    return e;
  return whiten_from_split(hp, ws, W12, bias, B, C, d_out, l2_before, l2_after, out, st);
}

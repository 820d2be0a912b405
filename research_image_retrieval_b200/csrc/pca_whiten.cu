// pca_whiten.cu — the dense part of PCA-whitening LEARN.
//
// Replaces, in pcawhitenlearn_shrinkage (networks/backbone.py:42-58):
//     m    = X.mean(axis=0, keepdims=True)                         :47
//     Xc   = X - m                                                  :48
//     Xcov = np.dot(Xc.T, Xc); Xcov = (Xcov + Xcov.T) / (2*N)       :49-50
// The symmetric eigen-decomposition that follows (:51-56) stays a library call on the host side (offline, once per
// model; research_image_retrieval_b200/whitening.py), its output feeds rir_whiten.
//
//   col_sum_kernel / col_mean_kernel : two-stage column mean, deterministic (no atomics)
//   covariance                       : C = (X - m)^T (X - m) / N on the tensor cores — the centred descriptors are
//                                      transposed and split into exact bf16 pairs once, then dense_mma.cu's
//                                      split-bf16 tcgen05 contraction (fp32 accumulate in TMEM) runs over the upper
//                                      triangle of tiles and a finishing kernel mirrors it, so C is exactly symmetric
//                                      (== the reference's (Xcov + Xcov.T) / 2).
//   cov_syrk_kernel                  : the round-1 fp32-FMA CUDA-core SYRK, kept as an independent implementation for
//                                      the parity tests and timings (RIR_PCA_FP32=1).
// Work: 2*N*D^2 / 2 flops (x3 bf16 passes); N = 20,000 descriptors x D = 2048: 84 GFLOP fp32-equivalent.
#include <stdlib.h>
#include "rir_common.cuh"

namespace rir {

constexpr int kColThreads = 256;
constexpr int kRowSplits = 64;

__global__ void __launch_bounds__(kColThreads)
    col_sum_kernel(const float* __restrict__ X, long long N, int D, float* __restrict__ partial /*[kRowSplits, D]*/) {
  const int c = blockIdx.x * kColThreads + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= D) return;
  const long long per = (N + kRowSplits - 1) / kRowSplits;
  const long long lo = r * per, hi = (lo + per < N) ? lo + per : N;
  float acc = 0.f;
  for (long long n = lo; n < hi; ++n) acc += X[(size_t)n * D + c];  // consecutive threads -> consecutive columns
  partial[(size_t)r * D + c] = acc;
}

__global__ void __launch_bounds__(kColThreads)
    col_mean_kernel(const float* __restrict__ partial, long long N, int D, float* __restrict__ mean) {
  const int c = blockIdx.x * kColThreads + threadIdx.x;
  if (c >= D) return;
  double acc = 0.0;  // 64 partial sums: combine in fp64, fixed order
  for (int r = 0; r < kRowSplits; ++r) acc += (double)partial[(size_t)r * D + c];
  mean[c] = (float)(acc / (double)N);
}

constexpr int kCovTile = 128;
constexpr int kCovK = 16;
constexpr int kCovThreads = 256;

__global__ void __launch_bounds__(kCovThreads)
    cov_syrk_kernel(const float* __restrict__ X, const float* __restrict__ mean, long long N, int D,
                    float* __restrict__ C) {
  // upper-triangular tile index -> (bi, bj), bj >= bi
  const int T = (D + kCovTile - 1) / kCovTile;
  int t = blockIdx.x, bi = 0;
  while (t >= T - bi) {
    t -= T - bi;
    ++bi;
  }
  const int bj = bi + t;
  __shared__ __align__(16) float As[kCovK][kCovTile];
  __shared__ __align__(16) float Bs[kCovK][kCovTile];
  const int tid = threadIdx.x;
  const int ti = tid / 16, tj = tid % 16;  // 16 x 16 threads, 8 x 8 outputs each
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  // loader: 16 rows x 128 columns per operand = 2048 floats, 8 per thread (two float4 when aligned)
  const int lr = tid / 16;         // row inside the K step
  const int lc = (tid % 16) * 8;   // first of 8 columns
  const int ci = bi * kCovTile + lc, cj = bj * kCovTile + lc;
  float mi[8], mj[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mi[e] = (ci + e < D) ? mean[ci + e] : 0.f;
    mj[e] = (cj + e < D) ? mean[cj + e] : 0.f;
  }
  for (long long n0 = 0; n0 < N; n0 += kCovK) {
    const long long n = n0 + lr;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = 0.f, b = 0.f;
      if (n < N) {
        if (ci + e < D) a = X[(size_t)n * D + ci + e] - mi[e];
        if (cj + e < D) b = X[(size_t)n * D + cj + e] - mj[e];
      }
      As[lr][lc + e] = a;
      Bs[lr][lc + e] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kCovK; ++k) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ti * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ti * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tj * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tj * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  const float inv = 1.f / (float)N;
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    const int i = bi * kCovTile + ti * 8 + x;
    if (i >= D) continue;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      const int j = bj * kCovTile + tj * 8 + y;
      if (j >= D) continue;
      const float v = acc[x][y] * inv;
      if (bi == bj) {
        if (j >= i) {  // diagonal tile: keep the upper triangle, mirror it
          C[(size_t)i * D + j] = v;
          C[(size_t)j * D + i] = v;
        }
      } else {
        C[(size_t)i * D + j] = v;
        C[(size_t)j * D + i] = v;
      }
    }
  }
}

// dense_mma.cu
size_t syrk_split_workspace(long long N, int D);
int syrk_split_bf16(const float* X, const float* mean, long long N, int D, float* cov, void* workspace, cudaStream_t st);

static bool pca_use_fp32() {
  static const int v = getenv("RIR_PCA_FP32") ? atoi(getenv("RIR_PCA_FP32")) : 0;
  return v != 0;
}
static size_t col_partial_bytes(int D) { return ((size_t)kRowSplits * (size_t)D * sizeof(float) + 255) / 256 * 256; }

}  // namespace rir

using namespace rir;

extern "C" size_t rir_pca_covariance_workspace(int64_t N, int D) {
  if (D < 1 || N < 1) return 0;
  return col_partial_bytes(D) + syrk_split_workspace(N, D);
}

extern "C" int rir_pca_covariance(const float* X, int64_t N, int D, float* mean, float* cov, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(N >= 1 && D >= 1, "pca_covariance: bad shape N=%lld D=%d", (long long)N, D);
  RIR_REQUIRE(X && mean && cov && workspace, "pca_covariance: null pointer");
  RIR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pca_covariance: workspace must be 256-byte aligned");
  RIR_REQUIRE(workspace_bytes >= rir_pca_covariance_workspace(N, D), "pca_covariance: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  const dim3 g1((D + kColThreads - 1) / kColThreads, kRowSplits);
  col_sum_kernel<<<g1, kColThreads, 0, st>>>(X, N, D, partial);
  RIR_LAUNCH_OK();
  col_mean_kernel<<<(D + kColThreads - 1) / kColThreads, kColThreads, 0, st>>>(partial, N, D, mean);
  RIR_LAUNCH_OK();
  if (pca_use_fp32()) {
    const int T = (D + kCovTile - 1) / kCovTile;
    cov_syrk_kernel<<<T * (T + 1) / 2, kCovThreads, 0, st>>>(X, mean, N, D, cov);
    RIR_LAUNCH_OK();
    return RIR_OK;
  }
  return syrk_split_bf16(X, mean, N, D, cov, reinterpret_cast<uint8_t*>(workspace) + col_partial_bytes(D), st);
}

// descriptor_build.cu — feature maps -> global descriptors.
//
//   rir_pool             gem / GeMPooling / G2Pooling / spoc / MAC   (networks/RetrievalNet.py:318-325,359-365;
//                        models/gem_pooling.py:12-23; models/senet_g2.py:132-153; models/spoc.py:33-35;
//                        models/ultron_modules/ultron.py:193-205)
//   rir_l2_normalize     F.normalize(x, p=2, dim=-1)                 (networks/RetrievalNet.py:343,587,589; ...)
//   rir_whiten           1x1 conv / Linear whitening, W x + b         (networks/RetrievalNet.py:342,588; networks/spca.py:61-64)
//   rir_scale_mean_l2    multi-scale mean + re-normalise              (utils/helpfunc.py:31-44)
//   rir_pack_descriptors fp32 -> bf16 / fp8(+row scale) database rows
//
// The pooling kernel is the only heavy one (cfg-4: 2.15 GB of fp32 feature maps): HBM-bound, one warp per (b,c)
// plane, 128-bit streaming loads with 4 independent loads in flight per lane, warp-shuffle reduction.
// Algorithmic bytes = B*C*HW*sizeof(in) + B*C*4.
#include "rir_common.cuh"

namespace rir {

// ---------------------------------------------------------------------------------------------
// pooling
// ---------------------------------------------------------------------------------------------
enum PowKind : int { kPow1 = 0, kPow2 = 1, kPow3 = 2, kPowGeneric = 3 };

template <int MODE, int PK>
struct PoolOp {
  float p, eps;
  __device__ __forceinline__ float init() const { return MODE == RIR_POOL_MAX ? -INFINITY : 0.f; }
  __device__ __forceinline__ float fold(float acc, float x) const {
    if (MODE == RIR_POOL_MAX) return fmaxf(acc, x);
    const float v = fmaxf(x, eps);  // clamp(min=eps)
    if (MODE == RIR_POOL_AVG || PK == kPow1) return acc + v;
    if (PK == kPow2) return fmaf(v, v, acc);
    if (PK == kPow3) return fmaf(v * v, v, acc);
    return acc + exp2f(p * __log2f(v));  // v > 0 after the clamp
  }
  __device__ __forceinline__ float merge(float a, float b) const { return MODE == RIR_POOL_MAX ? fmaxf(a, b) : a + b; }
  __device__ __forceinline__ float finish(float acc, int hw) const {
    if (MODE == RIR_POOL_MAX) return acc;
    const float m = acc / (float)hw;
    if (MODE == RIR_POOL_AVG || PK == kPow1) return m;
    if (PK == kPow2) return sqrtf(m);
    if (PK == kPow3) return cbrtf(m);
    return powf(m, 1.0f / p);
  }
};

template <int DT>
__device__ __forceinline__ void unpack16(const uint4& v, float* f);
template <>
__device__ __forceinline__ void unpack16<RIR_F32>(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <>
__device__ __forceinline__ void unpack16<RIR_BF16>(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

constexpr int kPoolThreads = 256;

// Optional split epilogue (fused descriptor head, dense_mma.cu): besides the fp32 value, plane (b, c) is written as an
// exact bf16 pair v = s1 + s2 (+ a remainder below 2^-18 |v|) at [b * ld_split + c] — the operand layout of the
// tensor-core whitening that follows, so the pooled descriptors are never re-read and re-packed by another launch.
template <int DT, int MODE, int PK>
__global__ void __launch_bounds__(kPoolThreads)
    pool_kernel(const void* __restrict__ x, long long planes, int hw, float p, float eps, float alpha, float beta,
                float* __restrict__ out, __nv_bfloat16* __restrict__ s1, __nv_bfloat16* __restrict__ s2, int C,
                int ld_split) {
  pdl_wait();  // (no-op unless launched behind another kernel with programmatic stream serialization)
  pdl_launch_dependents();
  constexpr int ESZ = DT == RIR_F32 ? 4 : 2;
  constexpr int EPC = 16 / ESZ;
  const PoolOp<MODE, PK> op{p, eps};
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kPoolThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kPoolThreads / 32);
  const size_t plane_bytes = (size_t)hw * ESZ;
  const bool vec = (plane_bytes & 15) == 0;
  const int chunks = (int)(plane_bytes >> 4);
  for (long long pl = warp0; pl < planes; pl += nwarps) {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(x) + (size_t)pl * plane_bytes;
    float acc = op.init();
    if (vec) {
      int c = lane;
      for (; c + 96 < chunks; c += 128) {  // 4 independent 16-byte loads in flight
        const uint4 v0 = ldg_stream_16B(base + (size_t)c * 16);
        const uint4 v1 = ldg_stream_16B(base + (size_t)(c + 32) * 16);
        const uint4 v2 = ldg_stream_16B(base + (size_t)(c + 64) * 16);
        const uint4 v3 = ldg_stream_16B(base + (size_t)(c + 96) * 16);
        float f[4][EPC];
        unpack16<DT>(v0, f[0]); unpack16<DT>(v1, f[1]); unpack16<DT>(v2, f[2]); unpack16<DT>(v3, f[3]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < EPC; ++e) acc = op.fold(acc, f[u][e]);
      }
      for (; c < chunks; c += 32) {
        const uint4 v0 = ldg_stream_16B(base + (size_t)c * 16);
        float f[EPC];
        unpack16<DT>(v0, f);
#pragma unroll
        for (int e = 0; e < EPC; ++e) acc = op.fold(acc, f[e]);
      }
    } else {
      for (int i = lane; i < hw; i += 32) {
        float v;
        if (DT == RIR_F32) v = reinterpret_cast<const float*>(base)[i];
        else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
        acc = op.fold(acc, v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = op.merge(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (lane == 0) {
      const float v = fmaf(alpha, op.finish(acc, hw), beta);
      out[pl] = v;
      if (s1 != nullptr) {
        const long long b = pl / C;
        const size_t o = (size_t)b * ld_split + (size_t)(pl - b * C);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        s1[o] = h;
        s2[o] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
}

struct SplitOut {
  __nv_bfloat16 *s1, *s2;
  int C, ld;
};

template <int DT, int MODE, int PK>
static int launch_pool(const void* x, long long planes, int hw, float p, float eps, float alpha, float beta,
                       float* out, cudaStream_t st, const SplitOut& so) {
  const long long warps_needed = planes;
  long long blocks = (warps_needed + (kPoolThreads / 32) - 1) / (kPoolThreads / 32);
  const long long max_blocks = (long long)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > max_blocks) blocks = max_blocks;
  pool_kernel<DT, MODE, PK><<<(unsigned)blocks, kPoolThreads, 0, st>>>(x, planes, hw, p, eps, alpha, beta, out, so.s1,
                                                                       so.s2, so.C, so.ld);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

template <int DT>
static int dispatch_pool(const void* x, long long planes, int hw, int mode, float p, float eps, float alpha,
                         float beta, float* out, cudaStream_t st, const SplitOut& so) {
  if (mode == RIR_POOL_MAX) return launch_pool<DT, RIR_POOL_MAX, kPow1>(x, planes, hw, p, eps, alpha, beta, out, st, so);
  if (mode == RIR_POOL_AVG) return launch_pool<DT, RIR_POOL_AVG, kPow1>(x, planes, hw, p, eps, alpha, beta, out, st, so);
  if (p == 3.0f) return launch_pool<DT, RIR_POOL_GEM, kPow3>(x, planes, hw, p, eps, alpha, beta, out, st, so);
  if (p == 2.0f) return launch_pool<DT, RIR_POOL_GEM, kPow2>(x, planes, hw, p, eps, alpha, beta, out, st, so);
  if (p == 1.0f) return launch_pool<DT, RIR_POOL_GEM, kPow1>(x, planes, hw, p, eps, alpha, beta, out, st, so);
  return launch_pool<DT, RIR_POOL_GEM, kPowGeneric>(x, planes, hw, p, eps, alpha, beta, out, st, so);
}

// pooling with the optional split-bf16 epilogue (arguments validated by the callers)
int launch_pool_split(const void* x, int dtype, int B, int C, int HW, int mode, float p, float eps, float alpha, float beta,
                      float* pooled, __nv_bfloat16* s1, __nv_bfloat16* s2, int ld_split, cudaStream_t st) {
  const SplitOut so{s1, s2, C, ld_split};
  const long long planes = (long long)B * C;
  if (dtype == RIR_F32) return dispatch_pool<RIR_F32>(x, planes, HW, mode, p, eps, alpha, beta, pooled, st, so);
  return dispatch_pool<RIR_BF16>(x, planes, HW, mode, p, eps, alpha, beta, pooled, st, so);
}

// ---------------------------------------------------------------------------------------------
// row-wise L2 normalisation (one warp per row, second read served by L1/L2)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ x, long long n, int d, float eps,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  const bool vec = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  for (long long r = warp0; r < n; r += nwarps) {
    const float* xr = x + (size_t)r * d;
    float* orow = out + (size_t)r * d;
    float ss = 0.f;
    if (vec) {
      for (int i = lane; i < (d >> 2); i += 32) {
        const float4 v = reinterpret_cast<const float4*>(xr)[i];
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
      }
    } else {
      for (int i = lane; i < d; i += 32) ss = fmaf(xr[i], xr[i], ss);
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), eps);
    if (vec) {
      for (int i = lane; i < (d >> 2); i += 32) {
        float4 v = reinterpret_cast<const float4*>(xr)[i];
        v.x /= denom; v.y /= denom; v.z /= denom; v.w /= denom;
        reinterpret_cast<float4*>(orow)[i] = v;
      }
    } else {
      for (int i = lane; i < d; i += 32) orow[i] = xr[i] / denom;
    }
  }
}

static int launch_l2(const float* x, long long n, int d, float eps, float* out, cudaStream_t st) {
  if (n == 0) return RIR_OK;
  long long blocks = (n + 7) / 8;
  const long long max_blocks = (long long)sm_count() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  l2_normalize_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, d, eps, out);
  RIR_LAUNCH_OK();
  return RIR_OK;
}
int launch_l2_rows(const float* x, long long n, int d, float eps, float* out, cudaStream_t st) {
  return launch_l2(x, n, d, eps, out, st);
}

// ---------------------------------------------------------------------------------------------
// whitening: out[B, d_out] = x[B, C] * W[d_out, C]^T + bias   (fp32 CUDA-core SGEMM, 64x64x16 tiles, 4x4 per thread)
// fp32 FMA keeps the 1e-5 parity bar against the reference's fp32 conv; the op is <1% of the pooling bytes.
// ---------------------------------------------------------------------------------------------
constexpr int kWBM = 64, kWBN = 64, kWBK = 16;

__global__ void __launch_bounds__(256)
    whiten_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, int B,
                  int C, int d_out, float* __restrict__ out) {
  __shared__ float As[kWBK][kWBM + 4];
  __shared__ float Bs[kWBK][kWBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a 4 x 4 micro tile
  const int m0 = blockIdx.y * kWBM, n0 = blockIdx.x * kWBN;
  float acc[4][4] = {};
  // loader mapping: 64 rows x 16 k per tile = 1024 elements, 4 per thread
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < C; k0 += kWBK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = k0 + lk + u;
      const int am = m0 + lr, bn = n0 + lr;
      As[lk + u][lr] = (am < B && kk < C) ? x[(size_t)am * C + kk] : 0.f;
      Bs[lk + u][lr] = (bn < d_out && kk < C) ? W[(size_t)bn * C + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kWBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d_out) out[(size_t)m * d_out + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// multi-scale mean + L2 (one warp per image)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    scale_mean_l2_kernel(const float* __restrict__ v, const uint8_t* __restrict__ keep, long long N, int S, int D,
                         float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long n = warp0; n < N; n += nwarps) {
    int kept = 0;
    for (int s = 0; s < S; ++s) kept += (keep == nullptr || keep[(size_t)n * S + s]) ? 1 : 0;
    const float cnt = (float)kept;  // kept == 0 -> 0/0 = NaN, like `vec /= (len(ms)-drop)` in the reference
    float ss = 0.f;
    for (int i = lane; i < D; i += 32) {
      float a = 0.f;
      for (int s = 0; s < S; ++s)
        if (keep == nullptr || keep[(size_t)n * S + s]) a += v[((size_t)n * S + s) * D + i];
      a = a / cnt;
      out[(size_t)n * D + i] = a;
      ss = fmaf(a, a, ss);
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    __syncwarp();
    for (int i = lane; i < D; i += 32) out[(size_t)n * D + i] = out[(size_t)n * D + i] / denom;
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 / fp8(+per-row scale)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pack_kernel(const float* __restrict__ v, long long n, int d, int dtype_out, void* __restrict__ out,
                float* __restrict__ scale_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long r = warp0; r < n; r += nwarps) {
    const float* row = v + (size_t)r * d;
    if (dtype_out == RIR_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + (size_t)r * d;
      for (int i = lane; i < d; i += 32) o[i] = __float2bfloat16_rn(row[i]);
    } else {
      float amax = 0.f;
      for (int i = lane; i < d; i += 32) amax = fmaxf(amax, fabsf(row[i]));
      amax = warp_max(amax);
      const float scale = amax > 0.f ? amax / 448.0f : 1.0f;
      if (lane == 0 && scale_out) scale_out[r] = scale;
      __nv_fp8_storage_t* o = reinterpret_cast<__nv_fp8_storage_t*>(out) + (size_t)r * d;
      for (int i = lane; i < d; i += 32) o[i] = __nv_cvt_float_to_fp8(row[i] / scale, __NV_SATFINITE, __NV_E4M3);
    }
  }
}

}  // namespace rir

// =============================================================================================
// C ABI
// =============================================================================================
using namespace rir;

extern "C" int rir_pool(const void* x, int dtype, int B, int C, int HW, int mode, float p, float eps, float alpha,
                        float beta, float* out, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(B >= 0 && C >= 1 && HW >= 1, "pool: bad shape B=%d C=%d HW=%d", B, C, HW);
  if (B == 0) return RIR_OK;
  RIR_REQUIRE(x && out, "pool: null pointer");
  RIR_REQUIRE(mode == RIR_POOL_GEM || mode == RIR_POOL_MAX || mode == RIR_POOL_AVG, "pool: bad mode %d", mode);
  RIR_REQUIRE(dtype == RIR_F32 || dtype == RIR_BF16, "pool: dtype must be f32 or bf16 (got %d)", dtype);
  RIR_REQUIRE(mode != RIR_POOL_GEM || p > 0.f, "pool: GeM exponent must be > 0 (got %g)", (double)p);
  RIR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "pool: feature maps must be 16-byte aligned");
  const long long planes = (long long)B * C;
  if (planes == 0) return RIR_OK;
  return launch_pool_split(x, dtype, B, C, HW, mode, p, eps, alpha, beta, out, nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int rir_l2_normalize(const float* x, int64_t n_rows, int d, float eps, float* out, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(x && out, "l2_normalize: null pointer");
  RIR_REQUIRE(n_rows >= 0 && d >= 1, "l2_normalize: bad shape");
  return launch_l2(x, n_rows, d, eps, out, (cudaStream_t)stream);
}

extern "C" int rir_whiten(const float* x, const float* W, const float* bias, int B, int C, int d_out, int l2_after,
                          float* out, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(x && W && out, "whiten: null pointer");
  RIR_REQUIRE(B >= 0 && C >= 1 && d_out >= 1, "whiten: bad shape B=%d C=%d d_out=%d", B, C, d_out);
  if (B == 0) return RIR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((d_out + kWBN - 1) / kWBN, (B + kWBM - 1) / kWBM);
  whiten_kernel<<<grid, 256, 0, st>>>(x, W, bias, B, C, d_out, out);
  RIR_LAUNCH_OK();
  if (l2_after) return launch_l2(out, B, d_out, 1e-12f, out, st);
  return RIR_OK;
}

extern "C" int rir_scale_mean_l2(const float* v, const uint8_t* keep, int64_t N, int S, int D, float* out,
                                 void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(v && out, "scale_mean_l2: null pointer");
  RIR_REQUIRE(N >= 0 && S >= 1 && D >= 1, "scale_mean_l2: bad shape");
  if (N == 0) return RIR_OK;
  long long blocks = (N + 7) / 8;
  const long long max_blocks = (long long)sm_count() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  scale_mean_l2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(v, keep, N, S, D, out);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

extern "C" int rir_pack_descriptors(const float* v, int64_t n, int d, int dtype_out, void* out, float* scale_out,
                                    void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(v && out, "pack_descriptors: null pointer");
  RIR_REQUIRE(n >= 0 && d >= 1, "pack_descriptors: bad shape");
  RIR_REQUIRE(dtype_out == RIR_BF16 || dtype_out == RIR_FP8E4M3, "pack_descriptors: bad dtype %d", dtype_out);
  RIR_REQUIRE(dtype_out != RIR_FP8E4M3 || scale_out, "pack_descriptors: fp8 needs scale_out");
  if (n == 0) return RIR_OK;
  long long blocks = (n + 7) / 8;
  const long long max_blocks = (long long)sm_count() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(v, n, d, dtype_out, out, scale_out);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

// query_expansion.cu — alpha query expansion.
//
//   q' = L2( q + sum_{j < kq} max(s_j, 0)^alpha * x_{idx_j} )          then sim_topk again with q'
//
// The reference tree has no alpha-QE; the nearest skeleton (search -> gather top-k rows -> renormalise -> search
// again) is IterativeQueryExpansion.expand_query, reference/manus/1_SPARSE/sparse_model.py:374-405.  The formula is
// Radenovic et al.'s alpha-QE from the revisitop / cirtorch lineage the reference's memo.md:42 points to.
//
// Sharded database: every rank calls rir_aqe_accumulate on the merged global top-k; it adds only the rows it owns,
// the partial sums are all-reduced by the host, then rir_aqe_finalize adds q and renormalises.
#include "rir_common.cuh"

namespace rir {

template <int DT>
__device__ __forceinline__ float load_elem(const void* base, size_t i);
template <>
__device__ __forceinline__ float load_elem<RIR_BF16>(const void* base, size_t i) {
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
}
template <>
__device__ __forceinline__ float load_elem<RIR_FP8E4M3>(const void* base, size_t i) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw(reinterpret_cast<const __nv_fp8_storage_t*>(base)[i], __NV_E4M3);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}
template <>
__device__ __forceinline__ float load_elem<RIR_F32>(const void* base, size_t i) {
  return reinterpret_cast<const float*>(base)[i];
}

template <int DT>
__global__ void __launch_bounds__(256)
    aqe_accumulate_kernel(const void* __restrict__ X, const float* __restrict__ x_scale, long long n_local,
                          long long idx_offset, int d, const float* __restrict__ sc, const int32_t* __restrict__ ix,
                          int ld_topk, int kq, float alpha, float* __restrict__ acc) {
  const int q = blockIdx.x;
  for (int j = 0; j < kq; ++j) {  // fixed order -> deterministic sums
    const long long gid = ix[(size_t)q * ld_topk + j];
    const long long loc = gid - idx_offset;
    if (gid < 0 || loc < 0 || loc >= n_local) continue;  // uniform per block
    const float s = fmaxf(sc[(size_t)q * ld_topk + j], 0.f);
    float w = powf(s, alpha);
    if (x_scale) w *= x_scale[loc];
    for (int i = threadIdx.x; i < d; i += blockDim.x)
      acc[(size_t)q * d + i] = fmaf(w, load_elem<DT>(X, (size_t)loc * d + i), acc[(size_t)q * d + i]);
  }
}

template <int DT>
__global__ void __launch_bounds__(256)
    aqe_finalize_kernel(const void* __restrict__ Q, const float* __restrict__ q_scale, const float* __restrict__ acc,
                        int d, float* __restrict__ out_f32, void* __restrict__ out_q, float* __restrict__ out_scale) {
  extern __shared__ float row[];  // [d]
  __shared__ float red[8];
  __shared__ float bcast[2];
  const int q = blockIdx.x;
  const float qs = q_scale ? q_scale[q] : 1.f;
  float ss = 0.f, amax = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float v = load_elem<DT>(Q, (size_t)q * d + i) * qs + acc[(size_t)q * d + i];
    row[i] = v;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    bcast[0] = fmaxf(sqrtf(t), 1e-12f);
  }
  __syncthreads();
  const float denom = bcast[0];
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float v = row[i] / denom;
    row[i] = v;
    amax = fmaxf(amax, fabsf(v));
    if (out_f32) out_f32[(size_t)q * d + i] = v;
  }
  if (out_q == nullptr) return;
  if (DT == RIR_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_q) + (size_t)q * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) o[i] = __float2bfloat16_rn(row[i]);
    if (out_scale && threadIdx.x == 0) out_scale[q] = 1.f;
  } else if (DT == RIR_FP8E4M3) {
    amax = warp_max(amax);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t = fmaxf(t, red[w]);
      bcast[1] = t > 0.f ? t / 448.0f : 1.0f;
      if (out_scale) out_scale[q] = bcast[1];
    }
    __syncthreads();
    const float scale = bcast[1];
    __nv_fp8_storage_t* o = reinterpret_cast<__nv_fp8_storage_t*>(out_q) + (size_t)q * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x)
      o[i] = __nv_cvt_float_to_fp8(row[i] / scale, __NV_SATFINITE, __NV_E4M3);
  }
}

}  // namespace rir

using namespace rir;

extern "C" int rir_aqe_accumulate(const void* X, int dtype, const float* x_scale, int64_t n_local, int64_t idx_offset,
                                  int d, const float* sc, const int32_t* ix, int nq, int ld_topk, int kq, float alpha,
                                  float* acc, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(X && sc && ix && acc, "aqe_accumulate: null pointer");
  RIR_REQUIRE(nq >= 0 && d >= 1 && kq >= 0 && ld_topk >= kq && n_local >= 0, "aqe_accumulate: bad shape");
  if (nq == 0 || kq == 0) return RIR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RIR_BF16)
    aqe_accumulate_kernel<RIR_BF16><<<nq, 256, 0, st>>>(X, x_scale, n_local, idx_offset, d, sc, ix, ld_topk, kq, alpha, acc);
  else if (dtype == RIR_FP8E4M3)
    aqe_accumulate_kernel<RIR_FP8E4M3><<<nq, 256, 0, st>>>(X, x_scale, n_local, idx_offset, d, sc, ix, ld_topk, kq, alpha, acc);
  else if (dtype == RIR_F32)
    aqe_accumulate_kernel<RIR_F32><<<nq, 256, 0, st>>>(X, x_scale, n_local, idx_offset, d, sc, ix, ld_topk, kq, alpha, acc);
  else {
    set_error("aqe_accumulate: bad dtype %d", dtype);
    return RIR_E_ARG;
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

extern "C" int rir_aqe_finalize(const void* Q, int dtype, const float* q_scale, const float* acc, int nq, int d,
                                float* out_f32, void* out_q, float* out_scale, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(Q && acc, "aqe_finalize: null pointer");
  RIR_REQUIRE(nq >= 0 && d >= 1 && d <= 16384, "aqe_finalize: bad shape");
  if (nq == 0) return RIR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)d * sizeof(float);
  if (dtype == RIR_BF16) {
    RIR_CUDA_OK(ensure_dyn_smem(aqe_finalize_kernel<RIR_BF16>, smem));
    aqe_finalize_kernel<RIR_BF16><<<nq, 256, smem, st>>>(Q, q_scale, acc, d, out_f32, out_q, out_scale);
  } else if (dtype == RIR_FP8E4M3) {
    RIR_CUDA_OK(ensure_dyn_smem(aqe_finalize_kernel<RIR_FP8E4M3>, smem));
    aqe_finalize_kernel<RIR_FP8E4M3><<<nq, 256, smem, st>>>(Q, q_scale, acc, d, out_f32, out_q, out_scale);
  } else if (dtype == RIR_F32) {
    RIR_REQUIRE(out_q == nullptr, "aqe_finalize: out_q must be NULL for fp32 queries (use out_f32)");
    RIR_CUDA_OK(ensure_dyn_smem(aqe_finalize_kernel<RIR_F32>, smem));
    aqe_finalize_kernel<RIR_F32><<<nq, 256, smem, st>>>(Q, q_scale, acc, d, out_f32, out_q, out_scale);
  } else {
    set_error("aqe_finalize: bad dtype %d", dtype);
    return RIR_E_ARG;
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

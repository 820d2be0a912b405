// dot_rows.cuh — CUDA-core fp32 dot products of database rows against an fp32 query held in shared memory.
// Shared by the exact / redo / rescore kernels (sim_topk_select.cu) and the full-ranking position kernels
// (rank_positions.cu): one definition, so a row scores bit-identically wherever it is scored.
#pragma once
#include "sim_topk.cuh"

namespace rir {

// ---------------------------------------------------------------------------------------------
// CUDA-core dot products (fallback / redo / rescore paths): one warp per row, fp32 FMA
// ---------------------------------------------------------------------------------------------
// acc += <16 bytes of a row, matching 16-byte chunk c of the fp32 query in shared memory>, fixed FMA order
template <int DT>
__device__ __forceinline__ float chunk_fma(const uint4& v, const float* qs, int c, float acc);

template <>
__device__ __forceinline__ float chunk_fma<RIR_BF16>(const uint4& v, const float* qs, int c, float acc) {
  const float4 q0 = reinterpret_cast<const float4*>(qs)[2 * c], q1 = reinterpret_cast<const float4*>(qs)[2 * c + 1];
  acc = fmaf(__uint_as_float(v.x << 16), q0.x, acc);
  acc = fmaf(__uint_as_float(v.x & 0xffff0000u), q0.y, acc);
  acc = fmaf(__uint_as_float(v.y << 16), q0.z, acc);
  acc = fmaf(__uint_as_float(v.y & 0xffff0000u), q0.w, acc);
  acc = fmaf(__uint_as_float(v.z << 16), q1.x, acc);
  acc = fmaf(__uint_as_float(v.z & 0xffff0000u), q1.y, acc);
  acc = fmaf(__uint_as_float(v.w << 16), q1.z, acc);
  acc = fmaf(__uint_as_float(v.w & 0xffff0000u), q1.w, acc);
  return acc;
}
template <>
__device__ __forceinline__ float chunk_fma<RIR_F32>(const uint4& v, const float* qs, int c, float acc) {
  const float4 q0 = reinterpret_cast<const float4*>(qs)[c];
  acc = fmaf(__uint_as_float(v.x), q0.x, acc);
  acc = fmaf(__uint_as_float(v.y), q0.y, acc);
  acc = fmaf(__uint_as_float(v.z), q0.z, acc);
  acc = fmaf(__uint_as_float(v.w), q0.w, acc);
  return acc;
}
__device__ __forceinline__ void fp8x4_to_float(uint32_t w, float* f) {
  const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w & 0xffffu), __NV_E4M3);
  const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w >> 16), __NV_E4M3);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
template <>
__device__ __forceinline__ float chunk_fma<RIR_FP8E4M3>(const uint4& v, const float* qs, int c, float acc) {
  const float* qp = qs + (size_t)c * 16;
  float f[16];
  fp8x4_to_float(v.x, f); fp8x4_to_float(v.y, f + 4); fp8x4_to_float(v.z, f + 8); fp8x4_to_float(v.w, f + 12);
#pragma unroll
  for (int e = 0; e < 16; ++e) acc = fmaf(f[e], qp[e], acc);
  return acc;
}

template <int DT>
__device__ __forceinline__ float dot_row(const uint8_t* row, const float* qs, int chunks, int lane) {
  float acc = 0.f;
  for (int c = lane; c < chunks; c += 32) acc = chunk_fma<DT>(ldg_stream_16B(row + (size_t)c * 16), qs, c, acc);
  return warp_sum(acc);
}

// Four consecutive rows at once: the four rows' loads of a chunk are independent, so four times as many bytes are
// in flight per warp (these paths are latency-bound).  Per row the arithmetic (and its order) is dot_row's.
template <int DT>
__device__ __forceinline__ void dot_rows4(const uint8_t* row0, size_t stride, int nvalid, const float* qs, int chunks,
                                          int lane, float (&out)[4]) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane; c < chunks; c += 32) {
    uint4 v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
      v[r] = r < nvalid ? ldg_stream_16B(row0 + (size_t)r * stride + (size_t)c * 16) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = chunk_fma<DT>(v[r], qs, c, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) out[r] = warp_sum(acc[r]);
}

// query row q -> fp32 shared memory (q_scale folded in)
template <int DT>
__device__ __forceinline__ void load_query_f32(const SimParams& p, int q, float* qs) {
  for (int i = threadIdx.x; i < p.d; i += blockDim.x) {
    float v;
    if (DT == RIR_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.Q)[(size_t)q * p.d + i]);
    else if (DT == RIR_F32) v = reinterpret_cast<const float*>(p.Q)[(size_t)q * p.d + i];
    else {
      const __half_raw h =
          __nv_cvt_fp8_to_halfraw(reinterpret_cast<const __nv_fp8_storage_t*>(p.Q)[(size_t)q * p.d + i], __NV_E4M3);
      v = __half2float(*reinterpret_cast<const __half*>(&h));
    }
    if (p.q_scale) v *= p.q_scale[q];
    qs[i] = v;
  }
}

}  // namespace rir

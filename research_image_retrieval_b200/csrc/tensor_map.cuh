// tensor_map.cuh — CUtensorMap construction shared by the tcgen05 kernels (defined in sim_topk_mma.cu).
#pragma once
#include <cuda.h>
#include "rir_common.cuh"

namespace rir {

// 2-D tiled map over a row-major [rows, d] array of bf16 (RIR_BF16) or bytes (RIR_FP8E4M3): box = 128 bytes of a row x
// box_rows rows, SWIZZLE_128B, out-of-bounds elements read as zero.  Descriptors are cached per host thread.
int make_rowmajor_map(CUtensorMap* m, const void* base, long long rows, int d, int dtype, int box_rows);

// K-major, SWIZZLE_128B shared-memory matrix descriptor of a tile whose rows are 128 bytes apart in 8-row groups of
// 1024 bytes (what the maps above deliver)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

}  // namespace rir

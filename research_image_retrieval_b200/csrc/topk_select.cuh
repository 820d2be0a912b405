// topk_select.cuh — block-wide exact top-k selection on 64-bit ranking keys.
//
// Keys are distinct (the low word encodes the row index), so "the k largest keys" is a well defined set:
// it IS the reference ordering `np.argsort(-sim)` / `torch.topk` (iris_evaluate.py:386,
// reference/manus/7_AdaptiveHybridModel/modified/adaptive_hybrid_retrieval_complete.py:428) with the
// tie rule of SURVEY.md §7.3 (equal score -> lower index first).
//
//   block_select_topk : MSB-first radix select (8-bit digits, starting below the common prefix of min/max so
//                       narrow score ranges do not collapse into one histogram bin) -> compaction of the
//                       survivors into shared memory -> bitonic sort, descending.
#pragma once
#include "rir_common.cuh"

namespace rir {

// sort n (power of two) keys in shared memory, descending. All threads of the block must call.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* s, int n) {
  const int nthreads = blockDim.x;
  if (n <= 256) {
    // Small lists (the final k <= 256 of every search): ONE warp sorts with warp-level barriers — 28 block-wide
    // barriers of a 512-thread CTA cost more than the whole sort (the select kernels are latency-bound).
    __syncthreads();
    if (threadIdx.x < 32) {
      for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          __syncwarp();
          for (int t = threadIdx.x; t < (n >> 1); t += 32) {
            const int lo = 2 * t - (t & (stride - 1));
            const int hi = lo + stride;
            const bool desc = ((lo & size) == 0);
            const uint64_t a = s[lo], b = s[hi];
            if ((a < b) == desc) {
              s[lo] = b;
              s[hi] = a;
            }
          }
        }
      }
    }
    __syncthreads();
    return;
  }
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += nthreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = s[lo], b = s[hi];
        if ((a < b) == desc) {
          s[lo] = b;
          s[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

struct SelectScratch {
  uint32_t hist[256];
  unsigned long long kmin, kmax;
  unsigned long long prefix;  // value of (key >> shift) for the bucket holding the k-th key
  int shift;
  int need;
  int bucket_count;
  int out_count;
};

// Select the k largest of m keys produced by key_at(i), i in [0,m); write them sorted descending to dst[0..k),
// pad dst[k..kpad) with 0 (kpad = power of two >= k, dst lives in shared memory).  Returns (to all threads) the
// number of real keys written, min(m, k).  blockDim.x must be a multiple of 32 and >= 64.
template <class KeyAt>
__device__ int block_select_topk(KeyAt key_at, int m, int k, uint64_t* dst, int kpad, SelectScratch* sc) {
  const int tid = threadIdx.x, nthreads = blockDim.x;
  __syncthreads();
  if (m <= k) {
    for (int i = tid; i < kpad; i += nthreads) dst[i] = (i < m) ? key_at(i) : 0ull;
    block_bitonic_sort_desc(dst, kpad);
    return m;
  }
  // ---- min / max of the keys -> common prefix ----
  if (tid == 0) {
    sc->kmin = ~0ull;
    sc->kmax = 0ull;
    sc->out_count = 0;
  }
  __syncthreads();
  {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int i = tid; i < m; i += nthreads) {
      const unsigned long long key = key_at(i);
      lo = key < lo ? key : lo;
      hi = key > hi ? key : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long lo2 = __shfl_xor_sync(0xffffffffu, lo, o);
      const unsigned long long hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = lo2 < lo ? lo2 : lo;
      hi = hi2 > hi ? hi2 : hi;
    }
    if ((tid & 31) == 0) {
      atomicMin(&sc->kmin, lo);
      atomicMax(&sc->kmax, hi);
    }
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned long long diff = sc->kmin ^ sc->kmax;  // != 0: keys are distinct and m > k >= 1
    const int hb = 63 - __clzll((long long)diff);
    sc->shift = hb + 1;  // bits [shift,64) are common to all keys
    sc->prefix = (sc->shift >= 64) ? 0ull : (sc->kmax >> sc->shift);
    sc->need = k;
    sc->bucket_count = m;
  }
  __syncthreads();
  // ---- radix passes ----
  while (true) {
    const int shift_hi = sc->shift;
    const int need = sc->need;
    const int bucket = sc->bucket_count;
    if (shift_hi == 0 || bucket == need) break;  // the whole current bucket belongs to the top-k
    const int shift_lo = shift_hi > 8 ? shift_hi - 8 : 0;
    const uint32_t mask = (1u << (shift_hi - shift_lo)) - 1u;
    const unsigned long long prefix = sc->prefix;
    for (int i = tid; i < 256; i += nthreads) sc->hist[i] = 0;
    __syncthreads();
    // (Merging same-digit lanes with __match_any_sync before the atomic was tried: slower, 19.8 vs 17.1 us.)
    for (int i = tid; i < m; i += nthreads) {
      const unsigned long long key = key_at(i);
      const unsigned long long top = (shift_hi >= 64) ? 0ull : (key >> shift_hi);
      if (top == prefix) atomicAdd(&sc->hist[(uint32_t)(key >> shift_lo) & mask], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns digits [255-8l-7, 255-8l], walked from the top
      uint32_t h[8];
      uint32_t lane_sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        h[j] = sc->hist[255 - 8 * tid - j];
        lane_sum += h[j];
      }
      uint32_t incl = lane_sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      const uint32_t excl = incl - lane_sum;
      if (excl < (uint32_t)need && incl >= (uint32_t)need) {
        uint32_t cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (cum < (uint32_t)need && cum + h[j] >= (uint32_t)need) {
            const uint32_t digit = 255 - 8 * tid - j;
            sc->prefix = (prefix << (shift_hi - shift_lo)) | digit;
            sc->need = need - (int)cum;
            sc->bucket_count = (int)h[j];
          }
          cum += h[j];
        }
        sc->shift = shift_lo;
      }
    }
    __syncthreads();
  }
  // ---- compaction: exactly k keys satisfy (key >> shift) >= prefix ----
  {
    const int shift = sc->shift;
    const unsigned long long prefix = sc->prefix;
    for (int i = tid; i < kpad; i += nthreads) dst[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < m; i += nthreads) {
      const unsigned long long key = key_at(i);
      const unsigned long long top = (shift >= 64) ? 0ull : (key >> shift);
      if (top >= prefix) {
        const int slot = atomicAdd(&sc->out_count, 1);
        if (slot < kpad) dst[slot] = key;
      }
    }
  }
  block_bitonic_sort_desc(dst, kpad);
  return k;
}

__host__ __device__ __forceinline__ int pow2_ceil_int(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace rir

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
      // key 0 is padding ("nothing"): with fewer than k real keys among m > k inputs the radix select ends at
      // prefix 0, which every key satisfies — padding must not race the real keys for the kpad slots
      if (top >= prefix && key != 0ull) {
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

// ---------------------------------------------------------------------------------------------
// Fast path for the select that follows every scan: k <= 256 out of a few thousand candidates staged in shared memory.
// ONE histogram pass over a range-normalised score (kBuckets linear buckets between the smallest and largest ordered
// score) finds the bucket holding the k-th key; keys above it are in for sure, the keys of that bucket compete for the
// remaining places; both groups are ranked by counting (no sort, ~8 block barriers instead of ~30).
// Returns k (dst[0..k) sorted descending, dst[k..kpad) = 0), or -1 when the shape is degenerate for this scheme (all
// scores equal, or an over-full boundary bucket) and the caller must use block_select_topk.
// ---------------------------------------------------------------------------------------------
constexpr int kBuckets = 2048;
constexpr int kBoundaryMax = 512;   // keys of the boundary bucket that can be ranked here
constexpr int kFastSelectMaxK = 256;

struct BucketScratch {
  uint32_t hist[kBuckets];
  uint64_t sure[kFastSelectMaxK];
  uint64_t bnd[kBoundaryMax];
  uint32_t warp_tot[32];
  uint32_t lo, hi;
  int bstar, above, cb;
  int n_sure, n_bnd;
};

__device__ __forceinline__ int block_select_bucket(const uint64_t* keys, int m, int k, uint64_t* dst, int kpad,
                                                   BucketScratch* bs) {
  const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  // ---- range of the ordered scores ----
  if (tid == 0) {
    bs->lo = 0xFFFFFFFFu;
    bs->hi = 0u;
    bs->n_sure = 0;
    bs->n_bnd = 0;
    bs->bstar = -1;
  }
  for (int i = tid; i < kBuckets; i += nthreads) bs->hist[i] = 0u;
  __syncthreads();
  {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int i = tid; i < m; i += nthreads) {
      const uint32_t h = (uint32_t)(keys[i] >> 32);
      lo = min(lo, h);
      hi = max(hi, h);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      atomicMin(&bs->lo, lo);
      atomicMax(&bs->hi, hi);
    }
  }
  __syncthreads();
  const uint32_t lo = bs->lo, range = bs->hi - lo;
  if (range == 0u) return -1;  // (uniform: every thread sees the same value)
  int shift = 32 - __clz(range) - 11;  // (range >> shift) < 2048
  if (shift < 0) shift = 0;
  // ---- histogram ----
  for (int i = tid; i < m; i += nthreads) atomicAdd(&bs->hist[((uint32_t)(keys[i] >> 32) - lo) >> shift], 1u);
  __syncthreads();
  // ---- bucket of the k-th key, walking from the top: thread t owns buckets [NB-1-per*t-(per-1), NB-1-per*t] ----
  {
    const int per = kBuckets / nthreads;  // 4 at 512 threads, 16 at 128
    uint32_t mine = 0;
    for (int j = 0; j < per; ++j) mine += bs->hist[kBuckets - 1 - per * tid - j];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) bs->warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += bs->warp_tot[w];
    (void)nwarps;
    const uint32_t excl = base + incl - mine;
    if (excl < (uint32_t)k && excl + mine >= (uint32_t)k) {  // exactly one thread
      uint32_t cum = excl;
      for (int j = 0; j < per; ++j) {
        const int b = kBuckets - 1 - per * tid - j;
        const uint32_t h = bs->hist[b];
        if (cum < (uint32_t)k && cum + h >= (uint32_t)k) {
          bs->bstar = b;
          bs->above = (int)cum;
          bs->cb = (int)h;
        }
        cum += h;
      }
    }
  }
  __syncthreads();
  const int bstar = bs->bstar, above = bs->above, cb = bs->cb;
  if (bstar < 0 || cb > kBoundaryMax || above >= kFastSelectMaxK) return -1;
  // ---- compaction ----
  for (int i = tid; i < m; i += nthreads) {
    const uint64_t key = keys[i];
    const int b = (int)(((uint32_t)(key >> 32) - lo) >> shift);
    if (b > bstar) bs->sure[atomicAdd(&bs->n_sure, 1)] = key;
    else if (b == bstar) bs->bnd[atomicAdd(&bs->n_bnd, 1)] = key;
  }
  for (int i = k + tid; i < kpad; i += nthreads) dst[i] = 0ull;
  __syncthreads();
  // ---- ranks by counting (keys are distinct) ----
  const int need = k - above;
  for (int t = tid; t < above + cb; t += nthreads) {
    if (t < above) {
      const uint64_t key = bs->sure[t];
      int r = 0;
      for (int j = 0; j < above; ++j) r += bs->sure[j] > key ? 1 : 0;
      dst[r] = key;
    } else {
      const uint64_t key = bs->bnd[t - above];
      int r = 0;
      for (int j = 0; j < cb; ++j) r += bs->bnd[j] > key ? 1 : 0;
      if (r < need) dst[above + r] = key;
    }
  }
  __syncthreads();
  return k;
}

// number of keys strictly greater than `key` in a list sorted descending (0 = padding, always at the end)
__device__ __forceinline__ int count_greater_desc(const uint64_t* list, int n, uint64_t key) {
  int lo = 0, hi = n;  // first position whose key is <= `key`
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (list[mid] > key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Merge G lists of k keys each, every list sorted descending with 0-padding at its end, real keys distinct across
// lists: the rank of a key is its position in its own list plus, per other list, the number of greater keys (binary
// search) — no sort.  lists[G][k] in shared memory; writes dst[0..k) (0 = fewer than k real keys in total).
__device__ __forceinline__ void block_merge_sorted_lists(const uint64_t* lists, int G, int k, uint64_t* dst) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = 0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < G * k; i += blockDim.x) {
    const int g = i / k, j = i - g * k;
    const uint64_t key = lists[i];
    if (key == 0ull) continue;
    int r = j;
    for (int o = 0; o < G && r < k; ++o)
      if (o != g) r += count_greater_desc(lists + (size_t)o * k, k, key);
    if (r < k) dst[r] = key;
  }
  __syncthreads();
}

}  // namespace rir

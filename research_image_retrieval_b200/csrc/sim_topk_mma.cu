// sim_topk_mma.cu — Q·Xᵀ on the 5th-gen tensor cores with a fused threshold-filter epilogue.
//
// Replaces torch.mm(query_features, gallery_features.t()) (iris_evaluate.py:383) + the ranking that follows it
// (iris_evaluate.py:386; torch.topk at reference/manus/7_AdaptiveHybridModel/modified/
// adaptive_hybrid_retrieval_complete.py:428).  The [nq, n] score matrix lives only in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer, database ring : cp.async.bulk.tensor 2-D tiles (256 rows x 128 B of K, 128B swizzle)
//   warp 3      TMA producer, query ring    : MB x (<=128 queries x 128 B of K) (L2 resident)
//   warp 1      MMA issuer   : one thread issues tcgen05.mma (M=128, N=256, K=16|32) — fp32 accumulators in TMEM
//   warp 2      TMEM allocator
//   warps 4..   epilogue (4 warps per query block): tcgen05.ld 32 lanes x 16 columns, double buffered; thread == one
//               query; packed-bf16 pre-filter against the query's threshold tau -> per-lane bitmask; the rare
//               survivors are picked out of registers with a select tree, checked exactly in fp32, staged per thread
//               and appended to the candidate list with one deferred atomicAdd per batch
// Work item = (database tile, query super-block of MB x 128 queries).
//   MB = 1 : two 256-column accumulators, the epilogue of one tile overlaps the MMAs of the next
//   MB = 2 : two query blocks share every database chunk in shared memory (512 TMEM columns, one buffer: the epilogue
//            is not overlapped) — used from 2048 queries.
// What bounds the tensor regime (measured on B200): the bytes each SM ingests per flop — a 128x256 tile
// needs 48 KB per 512 MMA clocks (33% of the tensor pipe), a CTA pair 32 KB (52%), a pair with MB = 2 48 KB per 1024.
// Tried for the two-block shape (round 2): splitting its un-overlapped 512-column epilogue over twice the warps (two
// threads per query, half the columns each; 640 threads cap the kernel at 96 registers, small spills) — parity green
// but SLOWER: 1024 queries 1240 -> 1153 TFLOP/s, 2048: 1307 -> 1231, 4096: 1194 -> 1213, fp8 2028 -> 1899.  With the
// epilogue switched off the shape reaches 1545 TFLOP/s (95 % of burst): the dead time is TMEM-read / survivor bound,
// not issue-slot bound per warp.
// For single-block batches the query box is trimmed to the real number of queries.
//
// Thread-block clusters of 2:
//   one super-block   : the CTAs of a cluster take DIFFERENT database tiles and SHARE the query chunk — each CTA
//                       loads half of its rows and TMA-multicasts them to both shared memories;
//   several           : CTA PAIR (TWO, cta_group::2) — the two CTAs take the SAME database tile and neighbouring
//                       super-blocks; ONE M=256 tcgen05.mma issued by the leader spans both SMs, each CTA stages only
//                       its 128 of the tile's 256 rows and its own queries; every TMA load completes on the leader's
//                       barriers, tcgen05.commit (multicast) frees ring slots and publishes accumulators in both CTAs.
//                       (Without TWO — RIR_MMA_TWO=0 — the database chunk is multicast instead.  RIR_MMA_CLUSTER4=1:
//                       two pairs per cluster that multicast the query chunk to each other; measured slower.)
// A multicast slot may only be refilled once every CTA that shares it has consumed it: the MMA thread's
// tcgen05.commit for that ring is multicast to the `empty` barrier of all sharers.
//
// Fused mode (kModeFused): no separate sample / threshold launches.  The first `ra_rounds` rounds of the scan are
// the sample: the epilogue keeps the best kFusedTopT keys per (query, tile) in sample_keys (a pre-pass over the
// tile's unit maxima gives the collecting pass its starting threshold); ONE grid barrier follows (all CTAs are
// co-resident: one per SM); the epilogue threads of every CTA then derive tau for a few queries (3-pass radix over
// the top 24 key bits: tau = lower edge of the bucket holding the k-th best kept key — still a lower bound of the
// true k-th best, so the filter stays exact) and publish it in place over a sentinel; the remaining rounds poll
// their query's tau and filter.  The TMA / MMA warps keep running through the barrier (prefetching the next tiles).
// The first-phase keys are merged into the candidate list by final_select_kernel (sim_topk_select.cu).
#include <cuda.h>
#include <stdlib.h>
#include <atomic>
#include "sim_topk.cuh"

namespace rir {

constexpr int kTileM = 128;  // queries per block (TMEM lanes)
constexpr int kTileN = 256;  // database rows per tile (TMEM columns per accumulator)
constexpr int kABytes = kTileM * 128;  // 16 KB: 128 queries x 128 B of K
constexpr int kBBytes = kTileN * 128;  // 32 KB: 256 database rows x 128 B of K
constexpr int kMaxSlots = 8;
constexpr int kTmemCols = 512;
constexpr int kPend = 16;    // per-thread staged candidates before one atomicAdd reserves their slots (>= 16: a
                             // 16-column unit's survivors always fit after a reservation)
constexpr int kMaxTopT = 8;  // keys kept per (query, sample tile)
static_assert(kFusedTopT <= kMaxTopT, "fused sample keeps at most kMaxTopT keys");

// Two independent TMA rings: the database ring is deep (its loads come from HBM: ~4 us loaded latency, so bytes in
// flight decide the achieved bandwidth), the query ring holds L2 hits.
struct MmaSmemTail {
  uint64_t full_a[kMaxSlots], empty_a[kMaxSlots];
  uint64_t full_b[kMaxSlots], empty_b[kMaxSlots];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float xs[2][kTileN];
  uint32_t hist[256];      // radix histogram of the fused threshold
  uint32_t tau_prefix;     // radix-select state shared by the epilogue threads
  uint32_t tau_need;
  uint32_t tau_nz;
  uint32_t pad2;
};

struct MmaGeom {
  int nqb;            // query blocks of 128
  int nsb;            // query super-blocks of MB blocks
  long long ntiles;   // database tiles (scan) or sample blocks (sample mode)
  int kchunks;        // ceil(d / elements-per-128B)
  uint32_t idesc;     // tcgen05 instruction descriptor
  int x_streamed_once;
  int na, nb;         // ring depths (query chunks / database chunks)
  int csize;          // cluster size: 1, 2 or 4
  int npairs;         // CTA pairs per cluster (cta_group::2 only): 1, or 2 = two pairs that multicast the query chunk
  int share;          // 0 none, 1 query chunk shared (CTAs differ in tile), 2 database chunk shared (differ in super-block)
  int nqg;            // super-block groups = ceil(nsb / csize) when share == 2, else nsb
  long long rounds;   // persistent-loop trips, identical for every CTA (dummy items keep clusters in lock step)
  int a_rows;         // query rows per block actually loaded (multiple of 8 * sharers; 128 unless one small block)
  int fused;          // kModeFused
  int ra_rounds;      // fused: rounds [0, ra_rounds) are the sample phase
  int tile_n;         // database rows per tile: 256, or 128 when a small shard would leave the last round mostly idle
  int debug;          // development only (RIR_MMA_DEBUG): bit0 = skip the MMAs, bit1 = skip the epilogue (timing
                      // decomposition of mainloop vs epilogue; results are garbage)
  // Range mode (fused scan of one query block): CTA b owns the CONTIGUOUS rows [b * range_rows, (b + 1) * range_rows)
  // and walks them as tiles of h_main, ..., h_main, h_last rows (both multiples of 32, <= 256) — every CTA streams the
  // same number of bytes, so the last round is as full as the others (148 CTAs x 3.32 tiles used to run as 4 rounds
  // with the last one 32 % occupied and HBM under-subscribed).  Each height has its tensor map (box = h rows) and its
  // N = h tcgen05.mma.  0 = classic mode (256-row tiles dealt round-robin).
  int range_rows, h_main, h_last;
};

// range mode: rows and height of CTA b's tile in round rd
__device__ __forceinline__ void range_tile(const MmaGeom& g, int b, long long rd, long long* row0, int* h) {
  *row0 = (long long)b * g.range_rows + rd * g.h_main;
  *h = rd == g.rounds - 1 ? g.h_last : g.h_main;
}

enum { kShareNone = 0, kShareQ = 1, kShareX = 2 };

// which (virtual tile, super-block) this CTA works on in round `rd`; v may be >= ntiles (dummy item: all-OOB loads)
__device__ __forceinline__ void item_of(const MmaGeom& g, long long rd, int cluster_id, int nclusters, int crank,
                                        long long* v, int* sb) {
  const long long j = rd * nclusters + cluster_id;
  if (g.share == kShareQ) {
    *v = j * g.csize + crank;
    *sb = 0;
  } else if (g.share == kShareX) {
    // clusters of 2: the two CTAs take the same tile and neighbouring super-blocks.  CTA pairs in clusters of 4
    // (g.npairs == 2): pair p = crank >> 1 takes tile 2*tp + p, the CTAs of a pair take super-blocks 2*qg + (crank & 1)
    *v = (j / g.nqg) * g.npairs + (crank >> 1) * (g.npairs - 1);
    *sb = (int)(j % g.nqg) * (g.csize / g.npairs) + (g.npairs > 1 ? (crank & 1) : crank);
  } else {
    *v = j / g.nsb;
    *sb = (int)(j % g.nsb);
  }
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(ignored)=1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

// development timeline (rir_profile_timeline): one (meta, time) pair per call; no-op when not armed
__device__ __forceinline__ void tl_mark(const SimParams& p, int event, long long round) {
  if (p.timeline == nullptr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const unsigned long long i = atomicAdd(p.timeline, 1ull);
  if (i < (unsigned long long)p.timeline_cap) {
    p.timeline[1 + 2 * i] = ((unsigned long long)blockIdx.x << 32) | ((unsigned long long)event << 16) |
                            (unsigned long long)(round & 0xffff);
    p.timeline[2 + 2 * i] = t;
  }
}

// one thread per CTA: arrive on a global counter and wait until all `expected` CTAs have (bounded: a bug must trap)
__device__ __forceinline__ void grid_arrive_wait(uint32_t* ctr, uint32_t expected) {
  // Arrive with a release reduction (no return value to wait for), poll with relaxed loads (an acquire load per poll
  // would invalidate L1 every time), fence once at the end.  Under a bandwidth-saturating scan every L2 round trip
  // costs microseconds (tools/timeline.py), so the barrier is kept to the minimum number of them.
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
  const long long t0 = clock64();
  while (true) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= expected) break;
    __nanosleep(20);
    if (clock64() - t0 > 8000000000ll) {
      printf("librir: grid barrier timed out (block %d: %u of %u arrived)\n", (int)blockIdx.x, v, expected);
      __trap();
    }
  }
  __threadfence();
}

constexpr int kMaxFusedKeys = 1280;  // first-phase keys per query the in-kernel threshold can hold in registers

// All epilogue threads (NE = 128 or 256): tau for query q = lower edge of the 24-bit bucket that holds the k-th best
// kept key.  The keys are read once into registers; three 8-bit radix passes run on a shared-memory histogram.
// Publishes tau_score[q] over its sentinel (release) — readers poll the value, there is no second grid barrier.
template <int NE>
__device__ __forceinline__ void cta_fused_tau(const SimParams& p, int q, int k, MmaSmemTail* tail, int etid) {
  constexpr int KPT = kMaxFusedKeys / NE;
  const unsigned long long* keys = p.sample_keys + (size_t)q * p.sample_m;
  const int m = p.sample_m;
  auto bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NE) : "memory"); };
  unsigned long long kk[KPT];
  int nz = 0;
#pragma unroll
  for (int i = 0; i < KPT; ++i) {
    const int idx = etid + i * NE;
    kk[i] = idx < m ? __ldcg(keys + idx) : 0ull;
    nz += kk[i] != 0ull ? 1 : 0;
  }
  for (int i = etid; i < 256; i += NE) tail->hist[i] = 0u;
  if (etid == 0) tail->tau_nz = 0u;
  bar();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
  if ((etid & 31) == 0 && nz) atomicAdd(&tail->tau_nz, (uint32_t)nz);
  uint32_t prefix = 0u, need = (uint32_t)k;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = 56 - 8 * pass;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const unsigned long long key = kk[i];
      if (key != 0ull && (pass == 0 || (uint32_t)(key >> (shift + 8)) == prefix))
        atomicAdd(&tail->hist[(uint32_t)(key >> shift) & 255u], 1u);
    }
    bar();
    if (etid < 32) {
      // lane l owns digits [255-8l-7, 255-8l], walked from the top
      uint32_t h[8];
      uint32_t lane_sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        h[j] = tail->hist[255 - 8 * etid - j];
        lane_sum += h[j];
      }
      uint32_t incl = lane_sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (etid >= o) incl += v;
      }
      const uint32_t excl = incl - lane_sum;
      if (excl < need && incl >= need) {  // at most one lane (none when fewer than `need` keys: handled via tau_nz)
        uint32_t cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (cum < need && cum + h[j] >= need) {
            tail->tau_prefix = (prefix << 8) | (255u - 8u * etid - j);
            tail->tau_need = need - cum;
          }
          cum += h[j];
        }
      }
    }
    bar();
    prefix = tail->tau_prefix;
    need = tail->tau_need;
    for (int i = etid; i < 256; i += NE) tail->hist[i] = 0u;
    bar();
  }
  if (etid == 0) {
    // tau_score[q] itself is the flag: the host pre-set it to the sentinel kTauUnset (a NaN pattern no threshold can
    // be); readers poll the value — one L2 round trip less than a separate flag, and under a bandwidth-saturating scan
    // every dependent L2 access costs microseconds.
    const float tau = tail->tau_nz >= (uint32_t)k ? ordered_to_float(prefix << 8) : -INFINITY;
    p.tau_idx[q] = 0xFFFFFFFFu;  // every index passes at score == tau
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.tau_score + q), "r"(__float_as_uint(tau)) : "memory");
  }
}

// value j (runtime) of a register array, by a select tree (no dynamic indexing -> no local memory)
__device__ __forceinline__ float pick16(const float (&v)[16], int j) {
  float a[8], b[4], c[2];
#pragma unroll
  for (int t = 0; t < 8; ++t) a[t] = (j & 1) ? v[2 * t + 1] : v[2 * t];
#pragma unroll
  for (int t = 0; t < 4; ++t) b[t] = (j & 2) ? a[2 * t + 1] : a[2 * t];
#pragma unroll
  for (int t = 0; t < 2; ++t) c[t] = (j & 4) ? b[2 * t + 1] : b[2 * t];
  return (j & 8) ? c[1] : c[0];
}

template <int DT, int MB, int TWO>
__global__ void __launch_bounds__(128 + 128 * MB, 1)
    sim_mma_kernel(const SimParams p, const MmaGeom g, const __grid_constant__ CUtensorMap tmQ,
                   const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kNumBuf = 2 / MB;             // TMEM accumulator buffers
  constexpr int kBufCols = kTileN * MB;       // columns per buffer
  constexpr int kASlot = kABytes * MB;        // query ring slot stride
  constexpr int kBSlot = TWO ? kBBytes / 2 : kBBytes;  // CTA pair: each CTA holds its 128 of the tile's 256 rows
  constexpr int kEpiThreads = 128 * MB;
  constexpr int kEpiWarps = 4 * MB;
  // layout: B ring [nb][32 KB] | A ring [na][MB x 16 KB] | tail
  uint8_t* ring_b = smem;
  uint8_t* ring_a = smem + (size_t)g.nb * kBSlot;
  MmaSmemTail* tail = reinterpret_cast<MmaSmemTail*>(ring_a + (size_t)g.na * kASlot);
  constexpr int kElemsPerChunk = (DT == RIR_BF16) ? 64 : 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = g.csize > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / g.csize;
  const int nclusters = gridDim.x / g.csize;
  const uint16_t cmask = (uint16_t)((1u << g.csize) - 1u);
  const int ca = (!TWO && g.share == kShareQ) ? g.csize : 1;  // CTAs sharing one query chunk (multicast)
  const int cb = (!TWO && g.share == kShareX) ? g.csize : 1;  // CTAs sharing one database chunk (multicast)
  // CTA pair (TWO): the leader (cluster rank 0) issues every MMA for both SMs; both CTAs' TMA loads complete on the
  // LEADER's full barriers; tcgen05.commit releases ring slots / publishes accumulators in both CTAs.
  // In clusters of 4 two pairs (ranks {0,1} and {2,3}) work on neighbouring tiles with the SAME queries: each CTA
  // loads half of its query block and multicasts it to the CTA of the same parity in the other pair.
  const int pr = crank & 1;                          // rank inside the pair
  const int pi = crank >> 1;                         // pair inside the cluster
  const bool leader = !TWO || pr == 0;
  const uint32_t leader_rank = (uint32_t)(crank & ~1);
  const uint16_t pair_mask = (uint16_t)(0x3u << (2 * pi));

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) {  // SWIZZLE_128B tiles need 1024-byte aligned shared memory
      printf("librir: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    for (int s = 0; s < kMaxSlots; ++s) {
      mbar_init(&tail->full_a[s], 1);
      // a shared slot is free once all sharers' MMAs have read it (pairs sharing the query chunk: one commit per pair)
      mbar_init(&tail->empty_a[s], TWO ? g.npairs : ca);
      mbar_init(&tail->full_b[s], 1);
      mbar_init(&tail->empty_b[s], cb);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], TWO ? 2 * kEpiWarps : kEpiWarps);  // pair: both CTAs' epilogues free the leader
    }
    mbar_fence_init();
  }
  if ((warp == 0 || warp == 3) && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmXr);
  }
  if (warp == 2) {
    if (TWO) {
      tmem_alloc_2sm(&tail->tmem_base, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(&tail->tmem_base, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (g.csize > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;
  // PDL: everything above overlapped the tail of the previous kernel of the stream (the previous search's select /
  // merge, or the query pack); from here on its results (packed queries, clean workspace header) are visible.
  pdl_wait();
  pdl_launch_dependents();  // the select kernel's CTAs move in as this kernel's CTAs retire
  if (threadIdx.x == 0) tl_mark(p, 0, 0);

  // virtual tile -> first database row.  Dummy tiles (v >= ntiles) start past the last row (all-OOB loads).
  auto tile_row0 = [&](long long v) -> long long {
    if (v >= g.ntiles) return ((p.n + g.tile_n - 1) / g.tile_n) * (long long)g.tile_n;
    if (p.mode == kModeSample) return sample_block_row0((int)v, p.nblk, p.sblk);
    if (g.fused) return ((v * p.perm_mul) % p.perm_n) * (long long)g.tile_n;
    return v * (long long)g.tile_n;
  };

  if (warp == 0) {
    // ===================== TMA producer: database chunks =====================
    if (lane == 0) {
      uint64_t pol_x;
      if (g.x_streamed_once) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_x));
      else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_x));
      int s = 0;
      uint32_t ph = 0;
      const int slice_rows = g.tile_n / cb, slice_bytes = g.tile_n * 128 / cb;
      const uint32_t b_bytes = (uint32_t)(g.tile_n * 128);
      for (long long rd = 0; rd < g.rounds; ++rd) {
        long long v;
        int sb;
        item_of(g, rd, cluster_id, nclusters, crank, &v, &sb);
        int row0 = (int)tile_row0(v);
        if (!TWO && g.range_rows > 0) {  // range mode: my own contiguous rows, ragged tile through its own map
          long long r0;
          int h;
          range_tile(g, (int)blockIdx.x, rd, &r0, &h);
          for (int kc = 0; kc < g.kchunks; ++kc) {
            mbar_wait(&tail->empty_b[s], ph ^ 1u);
            if (kc == 0) tl_mark(p, 1, rd);
            mbar_expect_tx(&tail->full_b[s], (uint32_t)(h * 128));
            tma_tensor2d_g2s(ring_b + (size_t)s * kBSlot, rd == g.rounds - 1 ? &tmXr : &tmX, kc * kElemsPerChunk, (int)r0,
                             &tail->full_b[s], pol_x);
            if (++s == g.nb) { s = 0; ph ^= 1u; }
          }
          continue;
        }
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->empty_b[s], ph ^ 1u);
          if (kc == 0) tl_mark(p, 1, rd);
          uint8_t* dst = ring_b + (size_t)s * kBSlot;
          if (TWO) {  // my 128 rows into my shared memory; both halves complete on the leader's barrier
            if (leader) mbar_expect_tx(&tail->full_b[s], b_bytes);
            tma_tensor2d_g2s_2sm(dst, &tmX, kc * kElemsPerChunk, row0 + pr * (g.tile_n / 2),
                                 mapa_u32(smem_u32(&tail->full_b[s]), leader_rank), pol_x);
          } else {
            mbar_expect_tx(&tail->full_b[s], b_bytes);
            if (cb == 1)
              tma_tensor2d_g2s(dst, &tmX, kc * kElemsPerChunk, row0, &tail->full_b[s], pol_x);
            else  // my 1/cb of the rows, delivered to every CTA of the cluster
              tma_tensor2d_g2s_mcast(dst + (size_t)crank * slice_bytes, &tmX, kc * kElemsPerChunk,
                                     row0 + crank * slice_rows, &tail->full_b[s], cmask, pol_x);
          }
          if (++s == g.nb) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== TMA producer: query chunks (L2 resident) =====================
    if (lane == 0) {
      uint64_t pol_q;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_q));
      int s = 0;
      uint32_t ph = 0;
      const int slice_rows = g.a_rows / ca, slice_bytes = slice_rows * 128;
      const uint32_t tx_bytes = (uint32_t)(MB * g.a_rows * 128);
      for (long long rd = 0; rd < g.rounds; ++rd) {
        long long v;
        int sb;
        item_of(g, rd, cluster_id, nclusters, crank, &v, &sb);
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->empty_a[s], ph ^ 1u);
          if (TWO) {
            if (leader) mbar_expect_tx(&tail->full_a[s], 2u * tx_bytes);
          } else {
            mbar_expect_tx(&tail->full_a[s], tx_bytes);
          }
#pragma unroll
          for (int m = 0; m < MB; ++m) {
            uint8_t* dst = ring_a + (size_t)s * kASlot + (size_t)m * kABytes;
            const int qrow0 = (sb * MB + m) * kTileM;
            if (TWO && g.npairs == 1)
              tma_tensor2d_g2s_2sm(dst, &tmQ, kc * kElemsPerChunk, qrow0,
                                   mapa_u32(smem_u32(&tail->full_a[s]), leader_rank), pol_q);
            else if (TWO)  // my half of the block's rows, also delivered to the same-parity CTA of the other pair
              tma_tensor2d_g2s_2sm_mcast(dst + (size_t)pi * (kABytes / 2), &tmQ, kc * kElemsPerChunk,
                                         qrow0 + pi * (kTileM / 2), mapa_u32(smem_u32(&tail->full_a[s]), leader_rank),
                                         (uint16_t)((1u << pr) | (1u << (pr + 2))), pol_q);
            else if (ca == 1)
              tma_tensor2d_g2s(dst, &tmQ, kc * kElemsPerChunk, qrow0, &tail->full_a[s], pol_q);
            else
              tma_tensor2d_g2s_mcast(dst + (size_t)crank * slice_bytes, &tmQ, kc * kElemsPerChunk,
                                     qrow0 + crank * slice_rows, &tail->full_a[s], cmask, pol_q);
          }
          if (++s == g.na) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA pair: the leader only) =====================
    if (lane == 0 && leader) {
      int sa = 0, sb_ = 0;
      uint32_t pha = 0, phb = 0;
      int ab = 0;
      uint32_t aph = 0;
      for (long long rd = 0; rd < g.rounds; ++rd) {
        mbar_wait(&tail->tmem_empty[ab], aph ^ 1u);
        tc_fence_after();
        tl_mark(p, 2, rd);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * kBufCols);
        uint32_t idesc = g.idesc;
        if (!TWO && g.range_rows > 0) {  // ragged tile: N = its height
          long long r0;
          int h;
          range_tile(g, (int)blockIdx.x, rd, &r0, &h);
          idesc = (g.idesc & ~(0x3Fu << 17)) | ((uint32_t)(h >> 3) << 17);
        }
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->full_a[sa], pha);
          mbar_wait(&tail->full_b[sb_], phb);
          tc_fence_after();
          const uint64_t b_desc = make_smem_desc(smem_u32(ring_b + (size_t)sb_ * kBSlot));
#pragma unroll
          for (int m = 0; m < MB; ++m) {
            if (g.debug & 1) break;
            const uint64_t a_desc = make_smem_desc(smem_u32(ring_a + (size_t)sa * kASlot + (size_t)m * kABytes));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // advance 32 bytes of K inside the 128B swizzle atom: +2 in the (addr >> 4) field
              const uint32_t dcol = d_tmem + (uint32_t)(m * kTileN);
              const uint32_t acc = (uint32_t)((kc | j) != 0);
              if (TWO) {
                if (DT == RIR_BF16) umma_f16_2sm(dcol, a_desc + 2u * j, b_desc + 2u * j, idesc, acc);
                else umma_f8_2sm(dcol, a_desc + 2u * j, b_desc + 2u * j, idesc, acc);
              } else {
                if (DT == RIR_BF16) umma_f16(dcol, a_desc + 2u * j, b_desc + 2u * j, idesc, acc);
                else umma_f8(dcol, a_desc + 2u * j, b_desc + 2u * j, idesc, acc);
              }
            }
          }
          // both slots are free once these MMAs have read them; a shared slot is released in every sharer
          if (TWO) {
            umma_commit_2sm(&tail->empty_a[sa], cmask);       // every CTA that wrote into / shares this slot
            umma_commit_2sm(&tail->empty_b[sb_], pair_mask);  // database halves are private to the pair
          } else {
            if (ca == 1) umma_commit(&tail->empty_a[sa]); else umma_commit_mcast(&tail->empty_a[sa], cmask);
            if (cb == 1) umma_commit(&tail->empty_b[sb_]); else umma_commit_mcast(&tail->empty_b[sb_], cmask);
          }
          if (++sa == g.na) { sa = 0; pha ^= 1u; }
          if (++sb_ == g.nb) { sb_ = 0; phb ^= 1u; }
        }
        // accumulators complete -> epilogue (of both CTAs of a pair)
        if (TWO) umma_commit_2sm(&tail->tmem_full[ab], pair_mask); else umma_commit(&tail->tmem_full[ab]);
        tl_mark(p, 3, rd);
        if (++ab == kNumBuf) { ab = 0; aph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread == query =====================
    const int ewi = warp - 4;      // epilogue warp index
    const int mblk = ewi >> 2;     // query block inside the super-block
    const int ew = ewi & 3;        // == warp % 4: the TMEM lane quarter this warp may read
    const int etid = (int)threadIdx.x - 128;
    int ab = 0;
    uint32_t aph = 0;
    const size_t sample_ld = (size_t)p.sblk * kSampleBlockRows;
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); };
    // Survivors are staged per thread and their slots reserved with ONE atomicAdd per flush: the atomic's L2 round
    // trip (~1 us) must not be paid per candidate inside the tile loop (it was 96% of the kernel before).
    // The reservation is also DEFERRED: the atomic is issued when a staging buffer fills, its result is first used
    // when the other buffer fills (or at the end of the query's rounds), so its latency overlaps the tile loop.
    unsigned long long pend[2][kPend];
    int npend = 0;       // entries in pend[cur]
    int cur = 0;
    int out_n = 0;       // entries of pend[cur ^ 1] whose slots are reserved (query out_q, base out_base), not yet written
    uint32_t out_base = 0;
    int out_q = 0;
    int pend_q = 0;      // query of the entries in pend[cur]
    auto drain = [&]() {  // write the reserved batch
      for (int i = 0; i < out_n; ++i)
        if (out_base + (uint32_t)i < (uint32_t)p.cap) p.cand[(size_t)out_q * p.cap + out_base + i] = pend[cur ^ 1][i];
      out_n = 0;
    };
    // staging buffer full, or the thread moves on to another query: reserve slots for it and switch buffers.  The
    // batch reserved LAST time is written first — a tile ago, so that atomic has long returned.
    auto reserve = [&]() {
      drain();
      out_base = atomicAdd(&p.cnt[pend_q], (uint32_t)npend);
      out_n = npend;
      out_q = pend_q;
      cur ^= 1;
      npend = 0;
    };
    auto flush = [&]() {  // everything out (end of the scan / before the grid barrier)
      if (npend > 0) reserve();
      drain();
    };
    // fused: publish the first-phase keys, then compute (and publish) tau for this CTA's share of the queries
    auto fused_threshold = [&]() {
      if (etid == 0) tl_mark(p, 6, 0);
      // bar.sync orders every epilogue thread's key stores before thread 0's release-arrive (cumulativity: the same
      // pattern as cooperative groups' grid sync) — no per-thread fence, one L2 round trip less
      epi_bar();
      if (etid == 0) grid_arrive_wait(&p.gbar[0], gridDim.x);
      epi_bar();
      if (etid == 0) tl_mark(p, 7, 0);
      if (p.nq >= (int)gridDim.x) {
        for (int q = (int)blockIdx.x; q < p.nq; q += (int)gridDim.x) cta_fused_tau<kEpiThreads>(p, q, p.k, tail, etid);
      } else if ((int)blockIdx.x < 4 * p.nq) {
        // fewer queries than CTAs: up to four CTAs compute the same query — they publish the same value and the readers
        // get the fastest replica (the tail of the L2 latency under load is what they wait for).  Not more: 148 CTAs
        // reading one query's keys at once is an L2 hot spot (1 query, 126 k-row shard: +4 us).
        cta_fused_tau<kEpiThreads>(p, (int)blockIdx.x % p.nq, p.k, tail, etid);
      }
      if (etid == 0) tl_mark(p, 8, 0);
    };
    bool tau_ready = !g.fused;
    const bool scaled = p.q_scale != nullptr || p.x_scale != nullptr;  // bf16 rows: scores are the raw accumulators
    // sample / first phase: this thread's (query's) best `topt` keys of the current tile
    unsigned long long top[kMaxTopT];
    for (long long rd = 0; rd < g.rounds; ++rd) {
      if (g.fused && rd == g.ra_rounds) {
        flush();
        fused_threshold();
        tau_ready = true;
      }
      const int mode = g.fused ? (rd < g.ra_rounds ? (int)kModeSample : (int)kModeScanFilter) : p.mode;
      long long v;
      int sb;
      item_of(g, rd, cluster_id, nclusters, crank, &v, &sb);
      long long row0 = tile_row0(v);
      int tile_h = g.tile_n;                      // rows (= accumulator columns) of this round's tile
      bool tile_real = v < g.ntiles;
      if (!TWO && g.range_rows > 0) {
        range_tile(g, (int)blockIdx.x, rd, &row0, &tile_h);
        tile_real = row0 < p.n;
        v = blockIdx.x;                           // first-phase slot of this CTA in sample_keys
      }
      const int q = (sb * MB + mblk) * kTileM + ew * 32 + lane;
      const bool qvalid = q < p.nq && tile_real;
      if (q != pend_q) {
        if (npend > 0) reserve();
        pend_q = q;
      }
#pragma unroll
      for (int i = 0; i < kMaxTopT; ++i) top[i] = 0ull;
      float ts = INFINITY;
      uint32_t ti = 0;
      float qsc = 1.f;
      if (qvalid) {
        if (mode == kModeScanFilter && g.fused) {  // published by whichever CTA computed it: poll the value itself
          uint32_t bits;
          const long long t0 = clock64();
          while (true) {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(bits) : "l"(p.tau_score + q) : "memory");
            if (bits != kTauUnset) break;
            __nanosleep(20);
            if (clock64() - t0 > 8000000000ll) {
              printf("librir: fused threshold of query %d never arrived (block %d)\n", q, (int)blockIdx.x);
              __trap();
            }
          }
          ts = __uint_as_float(bits);
          ti = 0xFFFFFFFFu;
        } else if (mode == kModeScanFilter) {
          ts = __ldcg(&p.tau_score[q]);
          ti = __ldcg(&p.tau_idx[q]);
        }
        if (p.q_scale) qsc = p.q_scale[q];
      }
      // the barrier below lets a warp run at most one round ahead of the slowest: two buffers, by round parity
      const int xb = (int)(rd & 1);
      const __nv_bfloat162 ts2 = __float2bfloat162_rn(ts);
      if (p.x_scale) {  // stage this tile's row scales (uniform branch)
        for (int j = etid; j < tile_h; j += kEpiThreads) {
          const long long row = row0 + j;
          tail->xs[xb][j] = row < p.n ? p.x_scale[row] : 0.f;
        }
        epi_bar();
      }
      mbar_wait(&tail->tmem_full[ab], aph);
      tc_fence_after();
      if (etid == 0) tl_mark(p, 4, rd);
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * kBufCols + mblk * kTileN);
      // 16 columns at a time, double buffered: the tcgen05.ld of the next unit is in flight while this one is
      // processed (TMEM loads take a few hundred clocks while the MMAs of the other accumulator are running).
      // Survivors are found with a per-lane bitmask — no divergent per-column branches (the unrolled 32-way version
      // of this code took ~27k clocks per tile, 10x the budget) — and each lane then walks ITS bits, fetching the
      // score from registers with a select tree, so the score array is never indexed dynamically.
      auto scores16 = [&](const uint32_t (&vv)[16], int c0, float (&sc)[16]) {
        if (p.x_scale) {
#pragma unroll
          for (int j = 0; j < 16; ++j) sc[j] = __uint_as_float(vv[j]) * qsc * tail->xs[xb][c0 + j];
        } else if (scaled) {
#pragma unroll
          for (int j = 0; j < 16; ++j) sc[j] = __uint_as_float(vv[j]) * qsc;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) sc[j] = __uint_as_float(vv[j]);
        }
      };
      // First-phase pre-pass: the 8th largest of the tile's 16 unit maxima is a lower bound of the lane's final 8th
      // best score (8 distinct columns reach it), so the collecting pass can start with that threshold instead of
      // -inf: ~12 insertions per lane instead of ~50 (and far less divergence: every lane inserts exactly once per
      // unit here).  Costs a second read of the tile from TMEM.
      float t8[kMaxTopT];
#pragma unroll
      for (int i = 0; i < kMaxTopT; ++i) t8[i] = -INFINITY;
      auto prepass16 = [&](const uint32_t (&vv)[16], int c0) {
        float sc[16];
        scores16(vv, c0, sc);
        if (row0 + c0 + 16 > p.n) {  // last, partial tile: columns past the shard do not count
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (row0 + c0 + j >= p.n) sc[j] = -INFINITY;
        }
        float m = sc[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) m = fmaxf(m, sc[j]);
#pragma unroll
        for (int i = 0; i < kMaxTopT; ++i) {  // branch-free insertion of m into the descending list t8
          const float hi = fmaxf(t8[i], m);
          m = fminf(t8[i], m);
          t8[i] = hi;
        }
      };
      float thr0 = -INFINITY;  // set by the pre-pass (first-phase rounds only)
      auto process16 = [&](const uint32_t (&vv)[16], int c0) {
        float sc[16];
        scores16(vv, c0, sc);
        if (mode == kModeScanFilter) {
          // pre-filter on packed bf16 pairs (3 instructions per 2 columns): round-to-nearest is monotonic, so
          // score >= ts implies bf16(score) >= bf16(ts) — it can only let extra columns through, which the exact fp32
          // test below rejects.  bit j = column 2j, bit 16+j = column 2j+1.
          uint32_t mask = 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            mask |= __hge2_mask(__floats2bfloat162_rn(sc[2 * j], sc[2 * j + 1]), ts2) & (0x00010001u << j);
          // qvalid: lanes of queries that do not exist hold whatever the (trimmed) query box left in shared memory
          if (!qvalid) mask = 0u;
          // room for all of this unit's survivors is made BEFORE the loop, so the loop body stays tiny
          if (mask != 0u && npend + __popc(mask) > kPend) reserve();
#pragma unroll 1
          while (mask) {  // rare: ~k*n/S survivors per query over the whole scan
            const int b = __ffs(mask) - 1;
            mask &= mask - 1u;
            const int j = ((b & 15) << 1) | (b >> 4);
            const float s1 = pick16(sc, j);
            const long long row = row0 + c0 + j;
            if (row < p.n && passes(s1, (uint32_t)row, ts, ti)) pend[cur][npend++] = make_key(s1, (uint32_t)row);
          }
        } else if (qvalid) {
          if (mode == kModeSample && p.topt > 0) {
            // columns that can enter this lane's running top list: score >= its current last kept score
            unsigned long long last = top[0];
#pragma unroll
            for (int i = 1; i < kMaxTopT; ++i)
              if (i < p.topt) last = top[i];
            const float thr = last != 0ull ? fmaxf(key_score(last), thr0) : thr0;
            uint32_t mask = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) mask |= (sc[j] >= thr) ? (1u << j) : 0u;
#pragma unroll 1
            while (mask) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1u;
              const long long row = row0 + c0 + j;
              unsigned long long key = make_key(pick16(sc, j), (uint32_t)row);
              if (row < p.n && key > last) {
                // (a rank-based insertion — position by independent compares, every slot rewritten — was tried to
                //  shorten the dependency chain: slower, 30 vs 24.5 us for the first-phase tile.)
#pragma unroll
                for (int i = 0; i < kMaxTopT; ++i) {  // insertion into the sorted (descending) top list
                  if (i < p.topt && key > top[i]) {
                    const unsigned long long tmp = top[i];
                    top[i] = key;
                    key = tmp;
                  }
                }
                last = top[0];
#pragma unroll
                for (int i = 1; i < kMaxTopT; ++i)
                  if (i < p.topt) last = top[i];
              }
            }
          } else if (mode == kModeSample) {
            float4* out = reinterpret_cast<float4*>(p.sample_scores + (size_t)q * sample_ld + (size_t)v * kTileN + c0);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 o;
              o.x = (row0 + c0 + j + 0 < p.n) ? sc[j + 0] : -INFINITY;
              o.y = (row0 + c0 + j + 1 < p.n) ? sc[j + 1] : -INFINITY;
              o.z = (row0 + c0 + j + 2 < p.n) ? sc[j + 2] : -INFINITY;
              o.w = (row0 + c0 + j + 3 < p.n) ? sc[j + 3] : -INFINITY;
              out[j >> 2] = o;
            }
          } else {  // kModeScanAll: slot == row (cap >= n), no atomics
            unsigned long long* out = p.cand + (size_t)q * p.cap + (size_t)(row0 + c0);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (row0 + c0 + j < p.n) out[j] = make_key(sc[j], (uint32_t)(row0 + c0 + j));
          }
        }
      };
      if (!(g.debug & 2)) {
        uint32_t va[16], vb[16];
        // (uniform) first-phase pre-pass — pays off once many lanes collect at the same time (divergence); with a
        // handful of queries the second TMEM pass costs more than it saves (1 query, 126 k-row shard: +4 us)
        if (mode == kModeSample && p.topt == kMaxTopT && tile_h >= 16 * kMaxTopT && p.nq >= 16) {
          tmem_ld_32x32_x16(taddr, va);
#pragma unroll 1
          for (int c0 = 0; c0 < tile_h; c0 += 32) {
            tmem_ld_wait();
            tmem_ld_32x32_x16(taddr + (uint32_t)(c0 + 16), vb);
            prepass16(va, c0);
            tmem_ld_wait();
            if (c0 + 32 < tile_h) tmem_ld_32x32_x16(taddr + (uint32_t)(c0 + 32), va);
            prepass16(vb, c0 + 16);
          }
          thr0 = t8[kMaxTopT - 1];
        }
        tmem_ld_32x32_x16(taddr, va);
#pragma unroll 1
        for (int c0 = 0; c0 < tile_h; c0 += 32) {
          tmem_ld_wait();
          tmem_ld_32x32_x16(taddr + (uint32_t)(c0 + 16), vb);
          process16(va, c0);
          tmem_ld_wait();
          if (c0 + 32 < tile_h) tmem_ld_32x32_x16(taddr + (uint32_t)(c0 + 32), va);
          process16(vb, c0 + 16);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (etid == 0) tl_mark(p, 5, rd);
      if (lane == 0) {
        if (TWO) mbar_arrive_cluster(mapa_u32(smem_u32(&tail->tmem_empty[ab]), leader_rank));  // my pair's MMA thread
        else mbar_arrive(&tail->tmem_empty[ab]);
      }
      if (++ab == kNumBuf) { ab = 0; aph ^= 1u; }
      // slot = (query, sample tile): EVERY slot is written.  Range mode: the last CTA(s) may own no real row at all —
      // their slots must read "nothing" (top[] is still all zero), not whatever an earlier search left there (a stale
      // key above the true k-th score would raise tau and drop real neighbours).
      if (mode == kModeSample && p.topt > 0 && (qvalid || (g.range_rows > 0 && q < p.nq))) {
#pragma unroll
        for (int i = 0; i < kMaxTopT; ++i)
          if (i < p.topt) p.sample_keys[(size_t)q * p.sample_m + (size_t)v * p.topt + i] = top[i];
      }
    }
    flush();
    if (!tau_ready) fused_threshold();
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) tl_mark(p, 9, 0);
  if (g.csize > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it / signal it
  if (warp == 2) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// Encoding a tensor map costs a few microseconds of host time per call; a search loop re-uses the same database and
// query buffers, so the last few descriptors are kept per host thread.
struct MapCacheEntry {
  const void* base;
  long long rows;
  int d, dtype, box_rows;
  CUtensorMap map;
};
static thread_local MapCacheEntry g_map_cache[16];
static thread_local int g_map_cache_next = 0;

static int make_rowmajor_map_uncached(CUtensorMap* m, const void* base, long long rows, int d, int dtype, int box_rows);

int make_rowmajor_map(CUtensorMap* m, const void* base, long long rows, int d, int dtype, int box_rows) {
  for (const MapCacheEntry& e : g_map_cache)
    if (e.base == base && e.rows == rows && e.d == d && e.dtype == dtype && e.box_rows == box_rows && base != nullptr) {
      *m = e.map;
      return RIR_OK;
    }
  if (int e = make_rowmajor_map_uncached(m, base, rows, d, dtype, box_rows)) return e;
  MapCacheEntry& slot = g_map_cache[g_map_cache_next];
  g_map_cache_next = (g_map_cache_next + 1) % 16;
  slot = MapCacheEntry{base, rows, d, dtype, box_rows, *m};
  return RIR_OK;
}

static int make_rowmajor_map_uncached(CUtensorMap* m, const void* base, long long rows, int d, int dtype, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return RIR_E_CUDA;
  }
  const int esz = dtype == RIR_BF16 ? 2 : 1;
  const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)d * esz};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, dtype == RIR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                         const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d=%d dtype=%d)", (int)r, rows, d, dtype);
    return RIR_E_CUDA;
  }
  return RIR_OK;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static int cluster_size_for(int sms) {
  // pairs only: clusters of 4 cannot be placed on 16 of the 148 SMs (GPC sizes 16/18/20), which costs a second wave
  int c = (sms % 2 == 0) ? 2 : 1;
  const int o = env_int("RIR_MMA_CLUSTER", 0);  // tuning override (development only)
  if ((o == 1 || o == 2) && sms % o == 0) c = o;
  return c;
}

static long long gcd_ll(long long a, long long b) {
  while (b) {
    const long long t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// Launch attributes of the scan.  The FUSED scan spins on a grid barrier, so all its CTAs must be co-resident: it is
// launched with cudaLaunchAttributeCooperative — a grid that cannot be fully resident (another search running on a
// second stream, a foreign kernel holding SMs) waits at the launch instead of dead-locking inside the barrier.  PDL
// (programmatic stream serialization) lets the kernel's prologue overlap the previous kernel's tail.  Which
// combinations the driver accepts together with the cluster attribute is probed once per process: mode 3 =
// cooperative + PDL, 2 = cooperative only, 1 = neither accepted -> the fused route is disabled and the three-launch
// route (no in-kernel grid barrier) is used instead.
static std::atomic<int> g_fused_launch_mode{0};  // 0 = not probed yet
bool mma_fused_disabled() { return g_fused_launch_mode == 1; }

template <int DT, int MB, int TWO>
static int launch_mma_t(const SimParams& p, const MmaGeom& g, const CUtensorMap& tmQ, const CUtensorMap& tmX,
                        const CUtensorMap& tmXr, unsigned grid, cudaStream_t st) {
  const size_t smem_bytes =
      (size_t)g.nb * (TWO ? kBBytes / 2 : kBBytes) + (size_t)g.na * kABytes * MB + sizeof(MmaSmemTail);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128 + 128 * MB);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  RIR_CUDA_OK(ensure_dyn_smem(sim_mma_kernel<DT, MB, TWO>, smem_bytes));
  auto launch_with = [&](bool coop, bool pdl) -> cudaError_t {
    cudaLaunchAttribute attr[3];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)g.csize;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    if (coop) {
      attr[na].id = cudaLaunchAttributeCooperative;
      attr[na].val.cooperative = 1;
      ++na;
    }
    if (pdl && pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, sim_mma_kernel<DT, MB, TWO>, p, g, tmQ, tmX, tmXr);
  };
  if (!g.fused) {
    RIR_CUDA_OK(launch_with(false, true));
    RIR_LAUNCH_OK();
    return RIR_OK;
  }
  static const int forced = env_int("RIR_FUSED_LAUNCH_MODE", 0);  // development override: 2, 3, or 4 = no cooperative
  // Nsight Compute cannot replay a cooperative launch of clustered CTAs (measured on this pool: the profiled process
  // dies with "LaunchFailed" on the first such kernel).  Under the profiler every kernel runs alone on the device —
  // the co-residency the cooperative attribute exists to guarantee holds trivially — so the attribute is dropped when
  // ncu's injection environment is present.  Nothing else changes (same kernel, same grid, same barrier).
  static const bool profiled = getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr ||
                               getenv("NV_NSIGHT_INJECTION_PORT_BASE") != nullptr;
  if (forced == 4 || (profiled && forced == 0)) {
    RIR_CUDA_OK(launch_with(false, true));
    RIR_LAUNCH_OK();
    return RIR_OK;
  }
  if (g_fused_launch_mode == 0 && forced >= 2 && forced <= 3) g_fused_launch_mode = forced;
  if (g_fused_launch_mode == 0 || g_fused_launch_mode == 3) {
    const cudaError_t e = launch_with(true, true);
    if (e == cudaSuccess) {
      g_fused_launch_mode = 3;
      RIR_LAUNCH_OK();
      return RIR_OK;
    }
    cudaGetLastError();
    if (g_fused_launch_mode == 3) {
      set_error("fused scan: cooperative launch failed: %s", cudaGetErrorString(e));
      return RIR_E_CUDA;
    }
  }
  if (g_fused_launch_mode == 0 || g_fused_launch_mode == 2) {
    const cudaError_t e = launch_with(true, false);
    if (e == cudaSuccess) {
      g_fused_launch_mode = 2;
      RIR_LAUNCH_OK();
      return RIR_OK;
    }
    cudaGetLastError();
    if (g_fused_launch_mode == 2) {
      set_error("fused scan: cooperative launch failed: %s", cudaGetErrorString(e));
      return RIR_E_CUDA;
    }
    set_error("fused scan: the driver rejects a cooperative launch of %u CTAs in clusters of %d (%s); "
              "using the three-launch route", grid, g.csize, cudaGetErrorString(e));
  }
  g_fused_launch_mode = 1;
  return RIR_E_NOFUSE;
}

template <int DT>
static int launch_mma_d(const SimParams& p, const MmaGeom& g, const CUtensorMap& tmQ, const CUtensorMap& tmX,
                        const CUtensorMap& tmXr, unsigned grid, int mb, int two, cudaStream_t st) {
  if (two)
    return mb == 1 ? launch_mma_t<DT, 1, 1>(p, g, tmQ, tmX, tmXr, grid, st) : launch_mma_t<DT, 2, 1>(p, g, tmQ, tmX, tmXr, grid, st);
  return mb == 1 ? launch_mma_t<DT, 1, 0>(p, g, tmQ, tmX, tmXr, grid, st) : launch_mma_t<DT, 2, 0>(p, g, tmQ, tmX, tmXr, grid, st);
}

// ---------------------------------------------------------------------------------------------
// launch shape: blocks per CTA, CTA pairs, cluster size — shared by the planner (mma_can_fuse) and the launcher
// ---------------------------------------------------------------------------------------------
struct MmaShape {
  int nqb, mb, nsb;
  int csize, npairs, share, two;
  int na, nb;
  int max_ctas;  // CTAs that can be co-resident with this shape (the grid of a full launch)
};

template <int MB>
static int max_active_clusters4() {  // clusters of 4 CTAs of the pair kernel that fit on the device at once
  static int cached = -1;
  if (cached >= 0) return cached;
  const int nslots = MB == 1 ? 6 : 4;
  const size_t smem_bytes = (size_t)nslots * (kBBytes / 2) + (size_t)nslots * kABytes * MB + sizeof(MmaSmemTail);
  auto kern = sim_mma_kernel<RIR_BF16, MB, 1>;
  int n = 0;
  if (ensure_dyn_smem(kern, smem_bytes) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(128 + 128 * MB);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
  }
  cudaGetLastError();
  cached = n;
  return cached;
}

static MmaShape plan_shape(int nq) {
  const int sms = sm_count();
  MmaShape h;
  h.nqb = (nq + kTileM - 1) / kTileM;
  // Measured on B200 (1M x 2048 bf16), CTA pairs: one block per CTA (epilogue overlapped with the next tile's MMAs)
  // 1185 TFLOP/s at 1024 queries and 1162 at 4096; two blocks per CTA (a third fewer bytes ingested per flop, but its
  // epilogue is not overlapped) 1174 and 1245 — and 1415 with the epilogue switched off.  Without pairs 1074-1107.
  h.mb = h.nqb >= 16 ? 2 : 1;
  {
    const int o = env_int("RIR_MMA_MB", 0);  // tuning override (development only)
    if (o == 1 || o == 2) h.mb = o;
  }
  h.nsb = (h.nqb + h.mb - 1) / h.mb;
  h.csize = cluster_size_for(sms);
  h.npairs = 1;
  h.share = h.csize == 1 ? kShareNone : (h.nsb == 1 ? kShareQ : kShareX);
  // CTA pair (cta_group::2): the two CTAs of a pair work on the same database tile and different super-blocks;
  // one M=256 MMA spans both SMs, each CTA stages only its 128 of the tile's 256 rows
  h.two = (h.share == kShareX && env_int("RIR_MMA_TWO", 1) != 0) ? 1 : 0;
  h.na = h.mb == 1 ? 4 : 3;  // ring depths: ~192 KB of shared memory either way
  h.nb = h.mb == 1 ? 4 : 3;
  h.max_ctas = sms - sms % h.csize;
  if (h.two) {
    h.na = h.nb = h.mb == 1 ? 6 : 4;
    // Optional (RIR_MMA_CLUSTER4=1): two pairs per cluster that multicast the query chunk to each other.  Measured on
    // B200 it halves the L2 reads of the query operand but is SLOWER overall (1024 queries: 1141 vs 1208 TFLOP/s;
    // 4096: 1123 vs 1189): clusters of 4 cannot be placed on all 148 SMs (GPC sizes), and the mainloop turned out to be
    // bound by the bytes each SM ingests (~33 B/clk), which multicast does not reduce.  Parity-tested, off by default.
    const int want = env_int("RIR_MMA_CLUSTER4", 0);
    const int n4 = want == 0 ? 0 : (h.mb == 1 ? max_active_clusters4<1>() : max_active_clusters4<2>());
    if (want != 0 && n4 * 4 * 100 >= sms * 85) {
      h.csize = 4;
      h.npairs = 2;
      h.max_ctas = n4 * 4 > sms ? sms - sms % 4 : n4 * 4;
    }
  }
  return h;
}

bool mma_can_fuse(int nq, long long n, int k) {
  if (env_int("RIR_MMA_FUSED", 1) == 0 || mma_fused_disabled()) return false;
  const MmaShape h = plan_shape(nq);
  const long long ntiles = (n + kTileN - 1) / kTileN;
  // the first phase (one tile per CTA and query) must hold >= 2k keys and be a small part of the scan
  return (long long)h.max_ctas * kFusedTopT >= 2ll * k && h.max_ctas * kFusedTopT <= kMaxFusedKeys &&
         ntiles >= 2ll * h.max_ctas;
}

int launch_sim_mma(SimParams& p, int dtype, cudaStream_t st) {
  if (dtype != RIR_BF16 && dtype != RIR_FP8E4M3) {
    set_error("sim_topk(mma): unsupported dtype %d", dtype);
    return RIR_E_ARG;
  }
  if (p.n >= (1ll << 31) - kTileN) {
    set_error("sim_topk(mma): shard too large (%lld rows)", p.n);
    return RIR_E_ARG;
  }
  const int sms = sm_count();
  const MmaShape h = plan_shape(p.nq);
  const int mb = h.mb, two = h.two;
  MmaGeom g;
  g.nqb = h.nqb;
  g.nsb = h.nsb;
  // Tile height: 256 rows.  128-row tiles were tried for small shards (8-way sharded cfg-2: 492 tiles / 148 CTAs =
  // 3.3 -> 4 rounds) to shorten the under-filled last round, but measured SLOWER (scan 0.121 vs 0.106 ms at 125,916
  // rows, 0.215 vs 0.172 ms at 251,831): the per-tile handshakes and epilogue outweigh the quantisation gain.
  // Kept behind RIR_MMA_TILE128=1 for experiments.
  g.tile_n = kTileN;
  if (p.mode == kModeFused && g.nqb == 1 && (p.n + kTileN - 1) / kTileN < 8ll * sms &&
      (p.n + 127) / 128 >= 2ll * sms && env_int("RIR_MMA_TILE128", 0) != 0)
    g.tile_n = 128;
  g.ntiles = p.mode == kModeSample ? p.sblk : (p.n + g.tile_n - 1) / g.tile_n;
  const int epc = dtype == RIR_BF16 ? 64 : 128;
  g.kchunks = (p.d + epc - 1) / epc;
  const uint32_t fmt = dtype == RIR_BF16 ? 1u : 0u;  // F16F32Format::BF16 = 1 ; MXF8F6F4Format::E4M3 = 0
  g.idesc = (1u << 4) /*D = f32*/ | (fmt << 7) | (fmt << 10) | ((uint32_t)(g.tile_n >> 3) << 17) |
            ((uint32_t)(kTileM >> 4) << 24);
  g.x_streamed_once = (g.nsb == 1);
  g.na = h.na;
  g.nb = h.nb;
  g.csize = h.csize;
  g.npairs = h.npairs;
  g.share = h.share;
  if (const char* e = getenv("RIR_MMA_RINGS")) {  // tuning override "na,nb" (development only)
    int a = 0, b = 0;
    if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && b >= 1 && a <= kMaxSlots && b <= kMaxSlots &&
        (size_t)b * (two ? kBBytes / 2 : kBBytes) + (size_t)a * kABytes * mb + sizeof(MmaSmemTail) <= 227 * 1024) {
      g.na = a;
      g.nb = b;
    }
  }
  if (two)
    g.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  const int ca = (!two && g.share == kShareQ) ? g.csize : 1, cb = two ? 2 : (g.share == kShareX ? g.csize : 1);
  // one small query block: load only the rows that exist (multiple of 8 rows per sharer: 1024-byte swizzle atoms)
  g.a_rows = kTileM;
  if (g.nqb == 1 && mb == 1 && env_int("RIR_MMA_TRIM", 1) != 0) {
    const int unit = 8 * ca;
    g.a_rows = (p.nq + unit - 1) / unit * unit;
    if (g.a_rows > kTileM) g.a_rows = kTileM;
  }
  // super-blocks per cluster item: a pair (or a 2-cluster) takes two neighbouring super-blocks
  g.nqg = g.share == kShareX ? (g.nsb + 1) / 2 : g.nsb;
  long long cluster_items;
  if (g.share == kShareQ) cluster_items = (g.ntiles + g.csize - 1) / g.csize;
  else if (g.share == kShareX) cluster_items = ((g.ntiles + g.npairs - 1) / g.npairs) * g.nqg;
  else cluster_items = g.ntiles * g.nsb;
  if (cluster_items <= 0) return RIR_OK;
  long long nclusters = h.max_ctas / g.csize;
  if (nclusters > cluster_items) nclusters = cluster_items;
  g.rounds = (cluster_items + nclusters - 1) / nclusters;
  const unsigned grid = (unsigned)(nclusters * g.csize);

  g.debug = env_int("RIR_MMA_DEBUG", 0);
  g.fused = p.mode == kModeFused ? 1 : 0;
  g.ra_rounds = 0;
  if (g.fused) {
    // first phase: every (query, tile v < grid) pair — one tile per CTA and query
    if ((int)grid != h.max_ctas || g.ntiles < 2ll * grid) {
      set_error("sim_topk(mma): fused scan needs a full grid (internal error: mma_can_fuse not honoured)");
      return RIR_E_ARG;
    }
    // rounds until every query has met `grid` tiles: a cluster item covers csize tiles (shared query chunk), or
    // npairs tiles x 2 super-blocks (shared database chunk / CTA pairs: 2 * nqg rounds either way)
    g.ra_rounds = g.share == kShareQ ? 1 : (g.share == kShareX ? g.nqg * 2 : g.nsb);
    p.topt = kFusedTopT;
    p.fused_tiles = (int)grid;
    p.sample_m = (int)grid * kFusedTopT;
    p.perm_n = g.ntiles;
    p.tile_rows = g.tile_n;
    long long mul = g.ntiles / grid;  // consecutive virtual tiles land ~ntiles/grid apart
    if (mul < 1) mul = 1;
    while (gcd_ll(mul, g.ntiles) != 1) ++mul;
    if (env_int("RIR_MMA_PERM", 1) == 0) mul = 1;  // tuning override (development only)
    p.perm_mul = mul;
  }
  // Range mode (see MmaGeom): one query block, no CTA pairs — every CTA streams its own contiguous rows.
  g.range_rows = g.h_main = g.h_last = 0;
  p.range_rows = p.first_rows = 0;
  if (g.fused && g.nsb == 1 && !two && g.tile_n == kTileN && env_int("RIR_MMA_RANGE", 1) != 0) {
    long long R = (p.n + grid - 1) / grid;
    R = (R + 31) / 32 * 32;
    int t = (int)(R % kTileN);
    if (t == 0) t = kTileN;
    if (t < 64) {  // a 32-row tile is not worth a round: lengthen the ranges by 32 rows (the last CTA takes the slack)
      R += 64 - t;
      t = 64;
    }
    const int rounds = (int)((R - t) / kTileN) + 1;  // >= 2: the fused scan needs >= 2 tiles per CTA
    // Tile heights.  256, ..., 256, t leaves a short last tile whose k-chunks are small: with 4 ring slots in flight
    // that round moves 12 KB instead of 32 KB per slot and runs latency-bound (measured on a 125,916-row shard: 14 us
    // for 0.38 of a tile).  Spreading the rows evenly — h_main = R / rounds rounded up to 32, the last tile takes the
    // rest — keeps every chunk near full size (864 rows: 224, 224, 224, 192 instead of 256, 256, 256, 96).  Measured,
    // 70 queries x 125,916 rows: scan 94.3 -> 90.8 us (CTA done 89.2 -> 81.8 us), 1 query 88.4 -> 84.1 us.
    int h_main = (int)(((R + rounds - 1) / rounds + 31) / 32 * 32);
    if (h_main > kTileN) h_main = kTileN;
    int h_last = (int)(R - (long long)h_main * (rounds - 1));
    if (h_last < 64 || env_int("RIR_MMA_EVEN_TILES", 1) == 0) {
      h_main = kTileN;
      h_last = t;
    }
    if (rounds >= 2 && R * (long long)grid < (1ll << 31)) {
      g.range_rows = (int)R;
      g.h_main = h_main;
      g.h_last = h_last;
      g.rounds = rounds;
      p.range_rows = (int)R;
      p.first_rows = h_main;
    }
  }
  CUtensorMap tmQ, tmX, tmXr;
  if (int e = make_rowmajor_map(&tmQ, p.Q, p.nq, p.d, dtype, two ? g.a_rows / g.npairs : g.a_rows / ca)) return e;
  if (int e = make_rowmajor_map(&tmX, p.X, p.n, p.d, dtype, g.range_rows > 0 ? g.h_main : g.tile_n / cb)) return e;
  tmXr = tmX;
  if (g.range_rows > 0 && g.h_last != g.h_main)
    if (int e = make_rowmajor_map(&tmXr, p.X, p.n, p.d, dtype, g.h_last)) return e;
  if (dtype == RIR_BF16) return launch_mma_d<RIR_BF16>(p, g, tmQ, tmX, tmXr, grid, mb, two, st);
  return launch_mma_d<RIR_FP8E4M3>(p, g, tmQ, tmX, tmXr, grid, mb, two, st);
}

}  // namespace rir

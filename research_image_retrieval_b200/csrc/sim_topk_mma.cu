// sim_topk_mma.cu — Q·Xᵀ on the 5th-gen tensor cores with a fused threshold-filter epilogue.
//
// Replaces torch.mm(query_features, gallery_features.t()) (iris_evaluate.py:383) + the ranking that follows it
// (iris_evaluate.py:386; torch.topk at reference/manus/7_AdaptiveHybridModel/modified/
// adaptive_hybrid_retrieval_complete.py:428).  The [nq, n] score matrix lives only in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer, database ring : cp.async.bulk.tensor 2-D tiles (256 rows x 128 B of K, 128B swizzle)
//   warp 3    TMA producer, query ring    : 128 queries x 128 B of K (L2 resident)
//   warp 1    MMA issuer   : one thread issues tcgen05.mma (M=128, N=256, K=16|32) — fp32 accumulators in TMEM,
//                            double buffered (2 x 256 columns) so the epilogue overlaps the next tile
//   warp 2    TMEM allocator
//   warps 4-7 epilogue     : tcgen05.ld 32 lanes x 32 columns; thread == one query; running max against the
//                            query's threshold tau; rare survivors are staged per thread and appended to the
//                            candidate list with one atomicAdd per flush
// Work item = (database tile, query block); items are tile-major so a database tile is pulled from HBM once and
// re-read from L2 by the other query blocks.
//
// Thread-block clusters (2 or 4 CTAs) cut the L2->SM traffic, which is what bounds this kernel once the epilogue
// is cheap (measured: ~7.2 TB/s L2->SM on B200):
//   one query block   : the CTAs of a cluster take DIFFERENT database tiles and SHARE the query chunk — each CTA
//                       loads 1/C of its rows and TMA-multicasts them to all C shared memories;
//   several blocks    : the CTAs of a cluster take the SAME database tile and different query blocks — the
//                       database chunk is the multicast operand.
// A multicast slot may only be refilled once every CTA of the cluster has consumed it: the MMA warp's
// tcgen05.commit for that ring is multicast to the `empty` barrier of all C CTAs (barrier count C).
#include <cuda.h>
#include <stdlib.h>
#include "sim_topk.cuh"

namespace rir {

constexpr int kMmaThreads = 256;
constexpr int kTileM = 128;  // queries per block (TMEM lanes)
constexpr int kTileN = 256;  // database rows per tile (TMEM columns per accumulator)
constexpr int kABytes = kTileM * 128;  // 16 KB: 128 queries x 128 B of K
constexpr int kBBytes = kTileN * 128;  // 32 KB: 256 database rows x 128 B of K
constexpr int kMaxSlots = 8;
constexpr int kTmemCols = 512;
constexpr int kPend = 8;
constexpr int kMaxTopT = 8;  // per-thread staged candidates before one atomicAdd reserves their slots

// Two independent TMA rings: the database ring is deep (its loads come from HBM: ~4 us loaded latency, so bytes in
// flight decide the achieved bandwidth), the query ring is shallow when the kernel is HBM-bound (its chunks are L2
// hits) and deeper when it is tensor-bound.
struct MmaSmemTail {
  uint64_t full_a[kMaxSlots], empty_a[kMaxSlots];
  uint64_t full_b[kMaxSlots], empty_b[kMaxSlots];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float xs[2][kTileN];
};

struct MmaGeom {
  int nqb;            // query blocks of 128
  long long ntiles;   // database tiles (scan) or sample blocks (sample mode)
  int kchunks;        // ceil(d / elements-per-128B)
  uint32_t idesc;     // tcgen05 instruction descriptor
  int x_streamed_once;
  int na, nb;         // ring depths (query chunks / database chunks)
  int csize;          // cluster size: 1, 2 or 4
  int share;          // 0 none, 1 query chunk shared (CTAs differ in tile), 2 database chunk shared (differ in query block)
  int nqg;            // query-block groups = ceil(nqb / csize) when share == 2, else nqb
  long long rounds;   // persistent-loop trips, identical for every CTA (dummy items keep clusters in lock step)
};

enum { kShareNone = 0, kShareQ = 1, kShareX = 2 };

// which (tile, query block) this CTA works on in round `rd`; tile may be >= ntiles (dummy item: all-OOB loads)
__device__ __forceinline__ void item_of(const MmaGeom& g, long long rd, int cluster_id, int nclusters, int crank,
                                        long long* tile, int* qb) {
  const long long j = rd * nclusters + cluster_id;
  if (g.share == kShareQ) {
    *tile = j * g.csize + crank;
    *qb = 0;
  } else if (g.share == kShareX) {
    *tile = j / g.nqg;
    *qb = (int)(j % g.nqg) * g.csize + crank;
  } else {
    *tile = j / g.nqb;
    *qb = (int)(j % g.nqb);
  }
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(ignored)=1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

template <int DT>
__global__ void __launch_bounds__(kMmaThreads, 1)
    sim_mma_kernel(const SimParams p, const MmaGeom g, const __grid_constant__ CUtensorMap tmQ,
                   const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: B ring [nb][32 KB] | A ring [na][16 KB] | tail
  uint8_t* ring_b = smem;
  uint8_t* ring_a = smem + (size_t)g.nb * kBBytes;
  MmaSmemTail* tail = reinterpret_cast<MmaSmemTail*>(ring_a + (size_t)g.na * kABytes);
  constexpr int kElemsPerChunk = (DT == RIR_BF16) ? 64 : 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = g.csize > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / g.csize;
  const int nclusters = gridDim.x / g.csize;
  const uint16_t cmask = (uint16_t)((1u << g.csize) - 1u);
  const int ca = g.share == kShareQ ? g.csize : 1;  // CTAs sharing one query chunk
  const int cb = g.share == kShareX ? g.csize : 1;  // CTAs sharing one database chunk

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) {  // SWIZZLE_128B tiles need 1024-byte aligned shared memory
      printf("librir: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    for (int s = 0; s < kMaxSlots; ++s) {
      mbar_init(&tail->full_a[s], 1);
      mbar_init(&tail->empty_a[s], ca);  // a shared slot is free once all sharers' MMAs have read it
      mbar_init(&tail->full_b[s], 1);
      mbar_init(&tail->empty_b[s], cb);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], 4);
    }
    mbar_fence_init();
  }
  if ((warp == 0 || warp == 3) && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 2) {
    tmem_alloc(&tail->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (g.csize > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  auto tile_row0 = [&](long long t) -> long long {  // dummy tiles (t >= ntiles) start past the last row
    if (t >= g.ntiles) return ((p.n + kTileN - 1) / kTileN) * (long long)kTileN;
    return p.mode == kModeSample ? sample_block_row0((int)t, p.nblk, p.sblk) : t * (long long)kTileN;
  };

  if (warp == 0) {
    // ===================== TMA producer: database chunks =====================
    if (lane == 0) {
      uint64_t pol_x;
      if (g.x_streamed_once) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_x));
      else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_x));
      int s = 0;
      uint32_t ph = 0;
      const int slice_rows = kTileN / cb, slice_bytes = kBBytes / cb;
      for (long long rd = 0; rd < g.rounds; ++rd) {
        long long t;
        int qb;
        item_of(g, rd, cluster_id, nclusters, crank, &t, &qb);
        const int row0 = (int)tile_row0(t);
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->empty_b[s], ph ^ 1u);
          mbar_expect_tx(&tail->full_b[s], kBBytes);
          uint8_t* dst = ring_b + (size_t)s * kBBytes;
          if (cb == 1)
            tma_tensor2d_g2s(dst, &tmX, kc * kElemsPerChunk, row0, &tail->full_b[s], pol_x);
          else  // my 1/cb of the rows, delivered to every CTA of the cluster
            tma_tensor2d_g2s_mcast(dst + (size_t)crank * slice_bytes, &tmX, kc * kElemsPerChunk,
                                   row0 + crank * slice_rows, &tail->full_b[s], cmask, pol_x);
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
      const int slice_rows = kTileM / ca, slice_bytes = kABytes / ca;
      for (long long rd = 0; rd < g.rounds; ++rd) {
        long long t;
        int qb;
        item_of(g, rd, cluster_id, nclusters, crank, &t, &qb);
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->empty_a[s], ph ^ 1u);
          mbar_expect_tx(&tail->full_a[s], kABytes);
          uint8_t* dst = ring_a + (size_t)s * kABytes;
          if (ca == 1)
            tma_tensor2d_g2s(dst, &tmQ, kc * kElemsPerChunk, qb * kTileM, &tail->full_a[s], pol_q);
          else
            tma_tensor2d_g2s_mcast(dst + (size_t)crank * slice_bytes, &tmQ, kc * kElemsPerChunk,
                                   qb * kTileM + crank * slice_rows, &tail->full_a[s], cmask, pol_q);
          if (++s == g.na) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int ab = 0;
      uint32_t aph = 0;
      for (long long rd = 0; rd < g.rounds; ++rd) {
        mbar_wait(&tail->tmem_empty[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * kTileN);
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->full_a[sa], pha);
          mbar_wait(&tail->full_b[sb], phb);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc(smem_u32(ring_a + (size_t)sa * kABytes));
          const uint64_t b_desc = make_smem_desc(smem_u32(ring_b + (size_t)sb * kBBytes));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // advance 32 bytes of K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            if (DT == RIR_BF16) umma_f16(d_tmem, a_desc + 2u * j, b_desc + 2u * j, g.idesc, (uint32_t)((kc | j) != 0));
            else umma_f8(d_tmem, a_desc + 2u * j, b_desc + 2u * j, g.idesc, (uint32_t)((kc | j) != 0));
          }
          // both slots are free once these MMAs have read them; a shared slot is released in every sharer
          if (ca == 1) umma_commit(&tail->empty_a[sa]); else umma_commit_mcast(&tail->empty_a[sa], cmask);
          if (cb == 1) umma_commit(&tail->empty_b[sb]); else umma_commit_mcast(&tail->empty_b[sb], cmask);
          if (++sa == g.na) { sa = 0; pha ^= 1u; }
          if (++sb == g.nb) { sb = 0; phb ^= 1u; }
        }
        umma_commit(&tail->tmem_full[ab]);  // accumulator complete -> epilogue
        if (++ab == 2) { ab = 0; aph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread == query =====================
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    int ab = 0;
    uint32_t aph = 0;
    const size_t sample_ld = (size_t)p.sblk * kSampleBlockRows;
    // Survivors are staged per thread and their slots reserved with ONE atomicAdd per flush: the atomic's L2 round
    // trip (~1 us) must not be paid per candidate inside the tile loop (it was 96% of the kernel before).
    unsigned long long pend[kPend];
    int npend = 0;
    int pend_q = 0;
    auto flush = [&]() {
      if (npend > 0) {
        const uint32_t base = atomicAdd(&p.cnt[pend_q], (uint32_t)npend);
        for (int i = 0; i < npend; ++i)
          if (base + (uint32_t)i < (uint32_t)p.cap) p.cand[(size_t)pend_q * p.cap + base + i] = pend[i];
        npend = 0;
      }
    };
    // sample mode with topt: this thread's (query's) best `topt` keys of the current sample tile
    unsigned long long top[kMaxTopT];
    for (long long rd = 0; rd < g.rounds; ++rd) {
      long long t;
      int qb;
      item_of(g, rd, cluster_id, nclusters, crank, &t, &qb);
      const long long row0 = tile_row0(t);
      const int q = qb * kTileM + ew * 32 + lane;
      const bool qvalid = q < p.nq && t < g.ntiles;
      if (q != pend_q) {
        flush();
        pend_q = q;
      }
#pragma unroll
      for (int i = 0; i < kMaxTopT; ++i) top[i] = 0ull;
      float ts = INFINITY;
      uint32_t ti = 0;
      float qsc = 1.f;
      if (qvalid) {
        if (p.mode == kModeScanFilter) {
          ts = p.tau_score[q];
          ti = p.tau_idx[q];
        }
        if (p.q_scale) qsc = p.q_scale[q];
      }
      if (p.x_scale) {  // stage this tile's row scales (uniform branch)
        const int e = ew * 32 + lane;
        for (int j = e; j < kTileN; j += 128) {
          const long long row = row0 + j;
          tail->xs[ab][j] = row < p.n ? p.x_scale[row] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(&tail->tmem_full[ab], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * kTileN);
#pragma unroll 1
      for (int c0 = 0; c0 < kTileN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        float sc[32];
        if (p.x_scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sc[j] = __uint_as_float(v[j]) * qsc * tail->xs[ab][c0 + j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) sc[j] = __uint_as_float(v[j]) * qsc;
        }
        if (p.mode == kModeScanFilter) {
          float mx = sc[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) mx = fmaxf(mx, sc[j]);
          if (mx >= ts) {  // rare: ~k*n/S survivors per query over the whole scan
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (sc[j] >= ts) {
                const long long row = row0 + c0 + j;
                if (row < p.n && passes(sc[j], (uint32_t)row, ts, ti)) {
                  if (npend == kPend) flush();
                  pend[npend++] = make_key(sc[j], (uint32_t)row);
                }
              }
            }
          }
        } else if (qvalid) {
          if (p.mode == kModeSample && p.topt > 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const long long row = row0 + c0 + j;
              if (row < p.n) {
                unsigned long long key = make_key(sc[j], (uint32_t)row);
                if (key > top[kMaxTopT - 1] || p.topt < kMaxTopT) {
#pragma unroll
                  for (int i = 0; i < kMaxTopT; ++i) {  // insertion into the sorted (descending) top list
                    if (i < p.topt && key > top[i]) {
                      const unsigned long long tmp = top[i];
                      top[i] = key;
                      key = tmp;
                    }
                  }
                }
              }
            }
          } else if (p.mode == kModeSample) {
            float4* out = reinterpret_cast<float4*>(p.sample_scores + (size_t)q * sample_ld + (size_t)t * kTileN + c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
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
            for (int j = 0; j < 32; ++j)
              if (row0 + c0 + j < p.n) out[j] = make_key(sc[j], (uint32_t)(row0 + c0 + j));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tmem_empty[ab]);
      if (++ab == 2) { ab = 0; aph ^= 1u; }
      if (p.mode == kModeSample && p.topt > 0 && qvalid) {  // slot = (query, sample tile): every slot is written
#pragma unroll
        for (int i = 0; i < kMaxTopT; ++i)
          if (i < p.topt) p.sample_keys[(size_t)q * p.sample_m + (size_t)t * p.topt + i] = top[i];
      }
    }
    flush();
  }

  tc_fence_before();
  __syncthreads();
  if (g.csize > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it / signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

static int make_rowmajor_map(CUtensorMap* m, const void* base, long long rows, int d, int dtype, int box_rows) {
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

int launch_sim_mma(const SimParams& p, int dtype, cudaStream_t st) {
  if (dtype != RIR_BF16 && dtype != RIR_FP8E4M3) {
    set_error("sim_topk(mma): unsupported dtype %d", dtype);
    return RIR_E_ARG;
  }
  if (p.n >= (1ll << 31) - kTileN) {
    set_error("sim_topk(mma): shard too large (%lld rows)", p.n);
    return RIR_E_ARG;
  }
  MmaGeom g;
  g.nqb = (p.nq + kTileM - 1) / kTileM;
  g.ntiles = p.mode == kModeSample ? p.sblk : (p.n + kTileN - 1) / kTileN;
  const int epc = dtype == RIR_BF16 ? 64 : 128;
  g.kchunks = (p.d + epc - 1) / epc;
  const uint32_t fmt = dtype == RIR_BF16 ? 1u : 0u;  // F16F32Format::BF16 = 1 ; MXF8F6F4Format::E4M3 = 0
  g.idesc = (1u << 4) /*D = f32*/ | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTileN >> 3) << 17) |
            ((uint32_t)(kTileM >> 4) << 24);
  g.x_streamed_once = (g.nqb == 1);
  // ring depths: measured on B200 (70 queries, 1M x 2048 bf16) 4+4 >= 3+5 >= 2+6 within 3% — the kernel is bound by
  // bytes delivered to the SMs (~6.7 TB/s for HBM fills + L2 hits together), not by bytes in flight
  g.na = 4;
  g.nb = 4;
  // clusters: share the query chunk across tiles (one block) or the database chunk across query blocks (several)
  const int sms = sm_count();
  // pairs only: clusters of 4 cannot be placed on 16 of the 148 SMs (GPC sizes 16/18/20), which costs a second wave
  g.csize = (sms % 2 == 0) ? 2 : 1;
  g.share = g.csize == 1 ? kShareNone : (g.nqb == 1 ? kShareQ : kShareX);
  if (const char* e = getenv("RIR_MMA_RINGS")) {  // tuning override "na,nb" (development only)
    int a = 0, b = 0;
    if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && b >= 1 && a <= kMaxSlots && b <= kMaxSlots &&
        (size_t)b * kBBytes + (size_t)a * kABytes + sizeof(MmaSmemTail) <= 227 * 1024) {
      g.na = a;
      g.nb = b;
    }
  }
  if (const char* e = getenv("RIR_MMA_CLUSTER")) {  // tuning override: cluster size 1 / 2 / 4 (development only)
    const int c = atoi(e);
    if ((c == 1 || c == 2 || c == 4) && sms % c == 0) {
      g.csize = c;
      g.share = c == 1 ? kShareNone : (g.nqb == 1 ? kShareQ : kShareX);
    }
  }
  g.nqg = g.share == kShareX ? (g.nqb + g.csize - 1) / g.csize : g.nqb;
  long long cluster_items;
  if (g.share == kShareQ) cluster_items = (g.ntiles + g.csize - 1) / g.csize;
  else if (g.share == kShareX) cluster_items = g.ntiles * g.nqg;
  else cluster_items = g.ntiles * g.nqb;
  if (cluster_items <= 0) return RIR_OK;
  long long nclusters = sms / g.csize;
  if (nclusters > cluster_items) nclusters = cluster_items;
  g.rounds = (cluster_items + nclusters - 1) / nclusters;
  const int ca = g.share == kShareQ ? g.csize : 1, cb = g.share == kShareX ? g.csize : 1;
  CUtensorMap tmQ, tmX;
  if (int e = make_rowmajor_map(&tmQ, p.Q, p.nq, p.d, dtype, kTileM / ca)) return e;
  if (int e = make_rowmajor_map(&tmX, p.X, p.n, p.d, dtype, kTileN / cb)) return e;
  const size_t smem_bytes = (size_t)g.nb * kBBytes + (size_t)g.na * kABytes + sizeof(MmaSmemTail);

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nclusters * g.csize));
  cfg.blockDim = dim3(kMmaThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)g.csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (dtype == RIR_BF16) {
    RIR_CUDA_OK(cudaFuncSetAttribute(sim_mma_kernel<RIR_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
    RIR_CUDA_OK(cudaLaunchKernelEx(&cfg, sim_mma_kernel<RIR_BF16>, p, g, tmQ, tmX));
  } else {
    RIR_CUDA_OK(cudaFuncSetAttribute(sim_mma_kernel<RIR_FP8E4M3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
    RIR_CUDA_OK(cudaLaunchKernelEx(&cfg, sim_mma_kernel<RIR_FP8E4M3>, p, g, tmQ, tmX));
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

}  // namespace rir

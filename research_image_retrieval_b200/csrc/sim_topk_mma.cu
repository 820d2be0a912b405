// sim_topk_mma.cu — Q·Xᵀ on the 5th-gen tensor cores with a fused threshold-filter epilogue.
//
// Replaces torch.mm(query_features, gallery_features.t()) (iris_evaluate.py:383) + the ranking that follows it
// (iris_evaluate.py:386; torch.topk at reference/manus/7_AdaptiveHybridModel/modified/
// adaptive_hybrid_retrieval_complete.py:428).  The [nq, n] score matrix lives only in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer : cp.async.bulk.tensor 2-D tiles, 128B-swizzled, into a 4-stage shared-memory ring
//                            (A = 128 queries x 128 B of K,  B = 256 database rows x 128 B of K)
//   warp 1    MMA issuer   : one thread issues tcgen05.mma (M=128, N=256, K=16|32) — fp32 accumulators in TMEM,
//                            double buffered (2 x 256 columns) so the epilogue overlaps the next tile
//   warp 2    TMEM allocator
//   warps 4-7 epilogue     : tcgen05.ld 32 lanes x 32 columns; thread == one query; running max against the
//                            query's threshold tau, rare survivors appended to the candidate list
// Work item = (database tile, query block); items are tile-major so a database tile is pulled from HBM once and
// re-read from L2 by the other query blocks.
#include <cuda.h>
#include "sim_topk.cuh"

namespace rir {

constexpr int kMmaThreads = 256;
constexpr int kTileM = 128;  // queries per block (TMEM lanes)
constexpr int kTileN = 256;  // database rows per tile (TMEM columns per accumulator)
constexpr int kStages = 4;
constexpr int kABytes = kTileM * 128;  // 16 KB
constexpr int kBBytes = kTileN * 128;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;

struct MmaSmemTail {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float xs[2][kTileN];
};
constexpr size_t kMmaSmemBytes = 1024 /*alignment slack*/ + (size_t)kStages * kStageBytes + sizeof(MmaSmemTail);

struct MmaGeom {
  int nqb;            // query blocks of 128
  long long ntiles;   // database tiles (scan) or sample blocks (sample mode)
  int kchunks;        // ceil(d / elements-per-128B)
  uint32_t idesc;     // tcgen05 instruction descriptor
  int x_streamed_once;
};

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
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  MmaSmemTail* tail = reinterpret_cast<MmaSmemTail*>(smem + (size_t)kStages * kStageBytes);
  constexpr int kElemsPerChunk = (DT == RIR_BF16) ? 64 : 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nitems = g.ntiles * g.nqb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], 4);
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 2) {
    tmem_alloc(&tail->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  auto tile_row0 = [&](long long t) -> long long {
    return p.mode == kModeSample ? sample_block_row0((int)t, p.nblk, p.sblk) : t * (long long)kTileN;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint64_t pol_q, pol_x;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_q));
      if (g.x_streamed_once) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_x));
      else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_x));
      int s = 0;
      uint32_t ph = 0;
      for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        const long long t = it / g.nqb;
        const int qb = (int)(it - t * g.nqb);
        const int row0 = (int)tile_row0(t);
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->empty[s], ph ^ 1u);
          uint8_t* a = smem + (size_t)s * kStageBytes;
          uint8_t* b = a + kABytes;
          mbar_expect_tx(&tail->full[s], kStageBytes);
          tma_tensor2d_g2s(a, &tmQ, kc * kElemsPerChunk, qb * kTileM, &tail->full[s], pol_q);
          tma_tensor2d_g2s(b, &tmX, kc * kElemsPerChunk, row0, &tail->full[s], pol_x);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int ab = 0;
      uint32_t aph = 0;
      for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        mbar_wait(&tail->tmem_empty[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * kTileN);
        for (int kc = 0; kc < g.kchunks; ++kc) {
          mbar_wait(&tail->full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * kStageBytes);
          const uint64_t a_desc = make_smem_desc(a_addr);
          const uint64_t b_desc = make_smem_desc(a_addr + kABytes);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // advance 32 bytes of K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            if (DT == RIR_BF16) umma_f16(d_tmem, a_desc + 2u * j, b_desc + 2u * j, g.idesc, (uint32_t)((kc | j) != 0));
            else umma_f8(d_tmem, a_desc + 2u * j, b_desc + 2u * j, g.idesc, (uint32_t)((kc | j) != 0));
          }
          umma_commit(&tail->empty[s]);  // frees the smem stage once these MMAs have read it
          if (++s == kStages) { s = 0; ph ^= 1u; }
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
    for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
      const long long t = it / g.nqb;
      const int qb = (int)(it - t * g.nqb);
      const long long row0 = tile_row0(t);
      const int q = qb * kTileM + ew * 32 + lane;
      const bool qvalid = q < p.nq;
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
          if (mx >= ts) {  // rare
#pragma unroll 1
            for (int j = 0; j < 32; ++j) {
              const long long row = row0 + c0 + j;
              if (row < p.n && passes(sc[j], (uint32_t)row, ts, ti)) push_candidate(p, q, sc[j], (uint32_t)row);
            }
          }
        } else if (qvalid) {
          if (p.mode == kModeSample) {
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
    }
  }

  tc_fence_before();
  __syncthreads();
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
  CUtensorMap tmQ, tmX;
  if (int e = make_rowmajor_map(&tmQ, p.Q, p.nq, p.d, dtype, kTileM)) return e;
  if (int e = make_rowmajor_map(&tmX, p.X, p.n, p.d, dtype, kTileN)) return e;
  MmaGeom g;
  g.nqb = (p.nq + kTileM - 1) / kTileM;
  g.ntiles = p.mode == kModeSample ? p.sblk : (p.n + kTileN - 1) / kTileN;
  const int epc = dtype == RIR_BF16 ? 64 : 128;
  g.kchunks = (p.d + epc - 1) / epc;
  const uint32_t fmt = dtype == RIR_BF16 ? 1u : 0u;  // F16F32Format::BF16 = 1 ; MXF8F6F4Format::E4M3 = 0
  g.idesc = (1u << 4) /*D = f32*/ | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTileN >> 3) << 17) |
            ((uint32_t)(kTileM >> 4) << 24);
  g.x_streamed_once = (g.nqb == 1);
  const long long nitems = g.ntiles * g.nqb;
  if (nitems <= 0) return RIR_OK;
  long long grid = nitems < (long long)sm_count() ? nitems : (long long)sm_count();
  if (dtype == RIR_BF16) {
    RIR_CUDA_OK(cudaFuncSetAttribute(sim_mma_kernel<RIR_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kMmaSmemBytes));
    sim_mma_kernel<RIR_BF16><<<(unsigned)grid, kMmaThreads, kMmaSmemBytes, st>>>(p, g, tmQ, tmX);
  } else {
    RIR_CUDA_OK(cudaFuncSetAttribute(sim_mma_kernel<RIR_FP8E4M3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kMmaSmemBytes));
    sim_mma_kernel<RIR_FP8E4M3><<<(unsigned)grid, kMmaThreads, kMmaSmemBytes, st>>>(p, g, tmQ, tmX);
  }
  RIR_LAUNCH_OK();
  return RIR_OK;
}

}  // namespace rir

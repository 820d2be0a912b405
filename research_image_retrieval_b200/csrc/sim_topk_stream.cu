// sim_topk_stream.cu — HBM-bound similarity scan for tiny query batches (nq <= 8).
//
// Replaces the GEMV-shaped case of torch.mm(q, g.t()) (iris_evaluate.py:383) where tensor cores cannot help:
// every database byte is read exactly once and used for <= 8 dot products.
//
// Design (B200): one persistent CTA per SM.  A single elected producer thread streams contiguous row blocks of
// the database into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), so
// 128-192 KB per SM are in flight independent of register pressure.  Eight consumer warps take RPW rows each
// from a landed stage, convert bf16/fp8 -> fp32 in registers and FMA against the fp32 query block held in
// shared memory.  Row scores are warp-reduced, compared with the per-query threshold tau and the rare survivors
// appended to the candidate list (sim_topk.cuh).  Algorithmic bytes per row = d * sizeof(element).
#include "sim_topk.cuh"

namespace rir {

constexpr int kStreamConsumerWarps = 8;
constexpr int kStreamThreads = (kStreamConsumerWarps + 1) * 32;
constexpr int kStreamMaxStages = 8;

template <int DT> struct Elem;
template <> struct Elem<RIR_BF16> {
  static constexpr int kPer16B = 8;
  // 16 bytes (8 bf16) -> 8 floats, memory order
  __device__ static __forceinline__ void cvt(const uint4& v, float (&f)[8]) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
  }
  __device__ static __forceinline__ float load1(const void* base, size_t i) {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
  }
};
template <> struct Elem<RIR_FP8E4M3> {
  static constexpr int kPer16B = 16;
  __device__ static __forceinline__ void cvt4(uint32_t w, float* f) {
    const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w & 0xffffu), __NV_E4M3);
    const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w >> 16), __NV_E4M3);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  __device__ static __forceinline__ void cvt(const uint4& v, float (&f)[16]) {
    cvt4(v.x, f); cvt4(v.y, f + 4); cvt4(v.z, f + 8); cvt4(v.w, f + 12);
  }
  __device__ static __forceinline__ float load1(const void* base, size_t i) {
    const __half_raw h = __nv_cvt_fp8_to_halfraw(reinterpret_cast<const __nv_fp8_storage_t*>(base)[i], __NV_E4M3);
    return __half2float(*reinterpret_cast<const __half*>(&h));
  }
};

template <> struct Elem<RIR_F32> {
  static constexpr int kPer16B = 4;
  __device__ static __forceinline__ void cvt(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ float load1(const void* base, size_t i) {
    return reinterpret_cast<const float*>(base)[i];
  }
};

struct StreamGeom {
  int rows_per_stage;  // kStreamConsumerWarps * RPW
  int stages;
  int stage_bytes;
  long long nstages_total;  // work items
};

template <int DT, int QB, int RPW>
__global__ void __launch_bounds__(kStreamThreads, 1) sim_stream_kernel(const SimParams p, const StreamGeom g) {
  using E = Elem<DT>;
  constexpr int EPC = E::kPer16B;
  extern __shared__ __align__(128) uint8_t smem[];
  // layout: ring[stages][stage_bytes] | q_f32[QB][d] | full[stages] | empty[stages]
  uint8_t* ring = smem;
  float* qs = reinterpret_cast<float*>(smem + (size_t)g.stages * g.stage_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(qs + (size_t)QB * p.d);
  uint64_t* empty = full + kStreamMaxStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = g.rows_per_stage;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kStreamConsumerWarps);
    }
    mbar_fence_init();
  }
  // query block -> fp32 shared (q_scale folded in); queries beyond nq are zero
  for (int i = threadIdx.x; i < QB * p.d; i += blockDim.x) {
    const int q = i / p.d, c = i - q * p.d;
    float v = 0.f;
    if (q < p.nq) {
      v = E::load1(p.Q, (size_t)(p.q0 + q) * p.d + c);
      if (p.q_scale) v *= p.q_scale[p.q0 + q];
    }
    // layout [q][h][chunk][4]: lanes walking consecutive 16-byte chunks read consecutive float4 (no bank conflicts)
    const int ch = c / EPC, w = c - ch * EPC;
    qs[(((size_t)q * (EPC / 4) + (w >> 2)) * (p.row_bytes >> 4) + ch) * 4 + (w & 3)] = v;
  }
  __syncthreads();

  const int spb = kSampleBlockRows / R;  // stages per sample block
  auto stage_row0 = [&](long long s) -> long long {
    if (p.mode == kModeSample) {
      const int j = (int)(s / spb), sub = (int)(s % spb);
      return sample_block_row0(j, p.nblk, p.sblk) + (long long)sub * R;
    }
    return s * (long long)R;
  };

  if (warp == kStreamConsumerWarps) {
    // ===== producer: one elected lane streams this CTA's stages into the ring =====
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (long long s = blockIdx.x; s < g.nstages_total; s += gridDim.x) {
        const long long row0 = stage_row0(s);
        long long rows = p.n - row0;
        rows = rows < 0 ? 0 : (rows > R ? R : rows);
        mbar_wait(&empty[slot], phase ^ 1u);
        if (rows > 0) {
          const uint32_t bytes = (uint32_t)rows * (uint32_t)p.row_bytes;
          mbar_expect_tx(&full[slot], bytes);
          tma_bulk_g2s(ring + (size_t)slot * g.stage_bytes,
                       reinterpret_cast<const uint8_t*>(p.X) + (size_t)row0 * p.row_bytes, bytes, &full[slot]);
        } else {
          mbar_arrive(&full[slot]);
        }
        if (++slot == g.stages) { slot = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ===== consumers =====
  float ts[QB];
  uint32_t ti[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    ts[q] = -INFINITY;
    ti[q] = 0xFFFFFFFFu;
    if (p.mode == kModeScanFilter && q < p.nq) {
      ts[q] = p.tau_score[p.q0 + q];
      ti[q] = p.tau_idx[p.q0 + q];
    }
  }
  const int chunks = p.row_bytes >> 4;
  int slot = 0;
  uint32_t phase = 0;
  unsigned long long best[QB];  // sample mode with topt: this warp's best key per query (held by lane 0)
#pragma unroll
  for (int q = 0; q < QB; ++q) best[q] = 0ull;
  for (long long s = blockIdx.x; s < g.nstages_total; s += gridDim.x) {
    const long long row0 = stage_row0(s);
    mbar_wait(&full[slot], phase);
    const uint8_t* st = ring + (size_t)slot * g.stage_bytes + (size_t)(warp * RPW) * p.row_bytes;
    float acc[RPW][QB];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int q = 0; q < QB; ++q) acc[r][q] = 0.f;

    for (int c = lane; c < chunks; c += 32) {
      float xf[RPW][EPC];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const uint4 xv = *reinterpret_cast<const uint4*>(st + (size_t)r * p.row_bytes + (size_t)c * 16);
        E::cvt(xv, xf[r]);
      }
#pragma unroll
      for (int q = 0; q < QB; ++q) {
#pragma unroll
        for (int h = 0; h < EPC / 4; ++h) {
          const float4 qv = reinterpret_cast<const float4*>(qs)[((size_t)q * (EPC / 4) + h) * chunks + c];
#pragma unroll
          for (int r = 0; r < RPW; ++r) {
            acc[r][q] = fmaf(xf[r][4 * h + 0], qv.x, acc[r][q]);
            acc[r][q] = fmaf(xf[r][4 * h + 1], qv.y, acc[r][q]);
            acc[r][q] = fmaf(xf[r][4 * h + 2], qv.z, acc[r][q]);
            acc[r][q] = fmaf(xf[r][4 * h + 3], qv.w, acc[r][q]);
          }
        }
      }
    }
    // the stage's shared memory is no longer needed by this warp
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
    if (++slot == g.stages) { slot = 0; phase ^= 1u; }

#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int q = 0; q < QB; ++q) acc[r][q] = warp_sum(acc[r][q]);

    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const long long row = row0 + warp * RPW + r;
        const bool valid = row < p.n;
        const float xs = (valid && p.x_scale) ? p.x_scale[row] : 1.f;
#pragma unroll
        for (int q = 0; q < QB; ++q) {
          if (q >= p.nq) continue;
          const float score = acc[r][q] * xs;
          if (p.mode == kModeSample) {
            if (p.topt > 0) {
              if (valid) {
                const unsigned long long key = make_key(score, (uint32_t)row);
                best[q] = key > best[q] ? key : best[q];
              }
            } else {
              const long long pos = s * (long long)R + warp * RPW + r;  // dense position inside the sample
              p.sample_scores[(size_t)(p.q0 + q) * ((size_t)p.sblk * kSampleBlockRows) + pos] =
                  valid ? score : -INFINITY;
            }
          } else if (valid) {
            if (p.mode == kModeScanAll)  // slot == row (cap >= n): no atomics, the select kernel reads n entries
              p.cand[(size_t)(p.q0 + q) * p.cap + (size_t)row] = make_key(score, (uint32_t)row);
            else if (passes(score, (uint32_t)row, ts[q], ti[q]))
              push_candidate(p, p.q0 + q, score, (uint32_t)row);
          }
        }
      }
    }
  }
  if (p.mode == kModeSample && p.topt > 0 && lane == 0) {
#pragma unroll
    for (int q = 0; q < QB; ++q)
      if (q < p.nq)
        p.sample_keys[(size_t)(p.q0 + q) * p.sample_m + (size_t)blockIdx.x * kStreamConsumerWarps + warp] = best[q];
  }
}

template <int DT, int QB, int RPW>
static int launch_stream_t(const SimParams& p, cudaStream_t st) {
  StreamGeom g;
  g.rows_per_stage = kStreamConsumerWarps * RPW;
  g.stage_bytes = g.rows_per_stage * p.row_bytes;
  const size_t fixed = (size_t)QB * p.d * sizeof(float) + 2 * kStreamMaxStages * sizeof(uint64_t) + 128;
  const size_t budget = 220 * 1024;
  int stages = (int)((budget - fixed) / (size_t)g.stage_bytes);
  if (stages > kStreamMaxStages) stages = kStreamMaxStages;
  if (stages < 2) return RIR_E_ARG;
  g.stages = stages;
  if (p.mode == kModeSample)
    g.nstages_total = (long long)p.sblk * (kSampleBlockRows / g.rows_per_stage);
  else
    g.nstages_total = (p.n + g.rows_per_stage - 1) / g.rows_per_stage;
  const size_t smem = (size_t)stages * g.stage_bytes + fixed;
  auto kern = sim_stream_kernel<DT, QB, RPW>;
  RIR_CUDA_OK(ensure_dyn_smem(kern, smem));
  long long grid = g.nstages_total < (long long)sm_count() ? g.nstages_total : (long long)sm_count();
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kStreamThreads, smem, st>>>(p, g);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

template <int DT, int QB>
static int launch_stream_q(const SimParams& p, cudaStream_t st) {
  // rows per warp: largest of {4,2,1} that still leaves >= 3 ring stages
  const size_t fixed = (size_t)QB * p.d * sizeof(float) + 2 * kStreamMaxStages * sizeof(uint64_t) + 128;
  const size_t budget = 220 * 1024 - fixed;
  const size_t row8 = (size_t)kStreamConsumerWarps * p.row_bytes;
  if (budget / (row8 * 4) >= 3) return launch_stream_t<DT, QB, 4>(p, st);
  if (budget / (row8 * 2) >= 3) return launch_stream_t<DT, QB, 2>(p, st);
  if (budget / row8 >= 2) return launch_stream_t<DT, QB, 1>(p, st);
  set_error("sim_topk(stream): descriptor dimension %d too large for the shared-memory ring", p.d);
  return RIR_E_ARG;
}

template <int DT>
static int launch_stream_d(const SimParams& p, cudaStream_t st) {
  if (p.nq <= 1) return launch_stream_q<DT, 1>(p, st);
  if (p.nq <= 2) return launch_stream_q<DT, 2>(p, st);
  if (p.nq <= 4) return launch_stream_q<DT, 4>(p, st);
  if (p.nq <= 8) return launch_stream_q<DT, 8>(p, st);
  set_error("sim_topk(stream): at most 8 queries per launch (got %d)", p.nq);
  return RIR_E_ARG;
}

int launch_sim_stream(const SimParams& p, int dtype, cudaStream_t st) {
  if (dtype == RIR_BF16) return launch_stream_d<RIR_BF16>(p, st);
  if (dtype == RIR_FP8E4M3) return launch_stream_d<RIR_FP8E4M3>(p, st);
  if (dtype == RIR_F32) return launch_stream_d<RIR_F32>(p, st);
  set_error("sim_topk(stream): unsupported dtype %d", dtype);
  return RIR_E_ARG;
}

}  // namespace rir

// evaluate_map.cu — revisited-protocol mAP / mP@k on the GPU.
//
// Replaces compute_ap (utils/evaluate.py:4-34), compute_map (utils/evaluate.py:37-150) and the three protocol
// passes of compute_map_and_print (utils/evaluate.py:163-185); duplicate copy at iris_evaluate.py:11-265.
//
// One CTA per query.  The ranked list is consumed in 256-position chunks: membership of each ranked id in the
// query's sorted id lists (binary search), then for every protocol a block-wide exclusive scan of the
// ok / junk flags gives each positive its ordinal i and its junk-adjusted rank r = j - #junk before j
// (the reference's while-loop, utils/evaluate.py:82-91).  The per-positive trapezoid terms are computed in
// parallel but ACCUMULATED by one thread in rank order, in fp64, with the reference's exact operation order
// ((p0 + p1) * recall_step / 2.0, utils/evaluate.py:32), so the results equal the Python floats bit for bit.
// A second single-block kernel forms the means over queries in query order (utils/evaluate.py:95,104-106).
#include "rir_common.cuh"

namespace rir {

constexpr int kMapThreads = 256;
constexpr int kMaxProto = 8;
constexpr int kMaxKappa = 16;

struct MapParams {
  const int32_t* ranks;
  const int32_t* positions;  // optional [nq, ld]: position of list entry j in the FULL ranking (ascending along j);
                             // nullptr: the entry's own index j (the list IS the ranking, possibly truncated)
  int nq;
  long long L, ld;
  const int32_t* ids[3];
  const int32_t* off[3];
  int proto[kMaxProto];
  int P;
  int kappas[kMaxKappa];
  int nk;
  double* map;     // [P]
  double* aps;     // [P, nq]
  double* mpr;     // [P, nk]
  double* prs;     // [P, nq, nk]
  int32_t* status; // [P, nq]
};

__device__ __forceinline__ bool in_sorted(const int32_t* a, int n, int32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int32_t x = a[mid];
    if (x < v) lo = mid + 1;
    else hi = mid;
  }
  return lo < n && a[lo] == v;
}

// exclusive prefix count of `flag` over the block; returns this thread's exclusive count, total in *total.
__device__ __forceinline__ int block_excl_count(bool flag, int* warp_tot /*[8]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  const int excl = __popc(m & ((1u << lane) - 1u));
  __syncthreads();  // protect warp_tot reuse
  if (lane == 0) warp_tot[warp] = __popc(m);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kMapThreads / 32; ++w) {
    const int c = warp_tot[w];
    if (w < warp) base += c;
    tot += c;
  }
  *total = tot;
  return base + excl;
}

__global__ void __launch_bounds__(kMapThreads) map_per_query_kernel(const MapParams p) {
  __shared__ int warp_tot[kMapThreads / 32];
  __shared__ double term[kMapThreads];
  __shared__ int rr[kMapThreads];
  __shared__ double s_ap[kMaxProto];
  __shared__ int s_npos[kMaxProto], s_njunk[kMaxProto], s_maxpos[kMaxProto];
  __shared__ int s_cntk[kMaxProto][kMaxKappa];

  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int32_t* lst[3];
  int len[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (p.ids[a] != nullptr && p.off[a] != nullptr) {
      lst[a] = p.ids[a] + p.off[a][q];
      len[a] = p.off[a][q + 1] - p.off[a][q];
    } else {
      lst[a] = nullptr;
      len[a] = 0;
    }
  }
  int nres[kMaxProto];
  double rstep[kMaxProto];
  for (int pr = 0; pr < p.P; ++pr) {
    int n = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (p.proto[pr] & (1 << a)) n += len[a];  // len(np.concatenate(ok lists)) — duplicates count, like the reference
    nres[pr] = n;
    rstep[pr] = n > 0 ? __ddiv_rn(1.0, (double)n) : 0.0;
  }
  if (tid < kMaxProto) {
    s_ap[tid] = 0.0;
    s_npos[tid] = 0;
    s_njunk[tid] = 0;
    s_maxpos[tid] = 0;
    for (int k = 0; k < kMaxKappa; ++k) s_cntk[tid][k] = 0;
  }
  __syncthreads();

  for (long long base = 0; base < p.L; base += kMapThreads) {
    const long long j = base + tid;
    const int32_t id = j < p.L ? p.ranks[(size_t)q * p.ld + j] : -1;
    int bits = 0;
    if (id >= 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
        if (len[a] > 0 && in_sorted(lst[a], len[a], id)) bits |= (1 << a);
    }
    if (__syncthreads_or(bits) == 0) continue;  // nothing relevant in this chunk (uniform)
    for (int pr = 0; pr < p.P; ++pr) {
      if (nres[pr] == 0) continue;  // uniform: empty ok list
      const bool isok = (bits & (p.proto[pr] & 7)) != 0;
      const bool isjunk = (bits & ((p.proto[pr] >> 4) & 7)) != 0;
      int tot_ok, tot_junk;
      const int ex_ok = block_excl_count(isok, warp_tot, &tot_ok);
      const int ex_junk = block_excl_count(isjunk, warp_tot, &tot_junk);
      if (isok) {
        const long long i = (long long)s_npos[pr] + ex_ok;           // ordinal of this positive
        const long long pj = p.positions ? (long long)p.positions[(size_t)q * p.ld + j] : j;
        const long long r = pj - ((long long)s_njunk[pr] + ex_junk);  // junk-adjusted 0-based rank
        const double p0 = (r == 0) ? 1.0 : __ddiv_rn((double)i, (double)r);
        const double p1 = __ddiv_rn((double)(i + 1), (double)(r + 1));
        term[ex_ok] = __ddiv_rn(__dmul_rn(__dadd_rn(p0, p1), rstep[pr]), 2.0);
        rr[ex_ok] = (int)r;
      }
      __syncthreads();
      if (tid == 0) {
        double ap = s_ap[pr];
        int mx = s_maxpos[pr];
        for (int t = 0; t < tot_ok; ++t) {
          ap = __dadd_rn(ap, term[t]);
          const int pos1 = rr[t] + 1;  // `pos += 1` (utils/evaluate.py:99)
          mx = pos1 > mx ? pos1 : mx;
          for (int k = 0; k < p.nk; ++k)
            if (pos1 <= p.kappas[k]) s_cntk[pr][k]++;
        }
        s_ap[pr] = ap;
        s_maxpos[pr] = mx;
        s_npos[pr] += tot_ok;
        s_njunk[pr] += tot_junk;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid < p.P) {
    const int pr = tid;
    const size_t o = (size_t)pr * p.nq + q;
    if (nres[pr] == 0) {
      p.aps[o] = INFINITY;
      p.status[o] = RIR_MAP_EMPTY_OK;
      for (int k = 0; k < p.nk; ++k) p.prs[o * p.nk + k] = INFINITY;
    } else {
      p.aps[o] = s_ap[pr];
      const int npos = s_npos[pr];
      p.status[o] = npos > 0 ? RIR_MAP_OK : RIR_MAP_NO_POS_RETRIEVED;
      for (int k = 0; k < p.nk; ++k) {
        double v = 0.0;
        if (npos > 0) {
          const int mx = s_maxpos[pr];
          const int kp = mx < p.kappas[k] ? mx : p.kappas[k];            // min(max(pos), keeps[k])
          const int c = (p.kappas[k] >= mx) ? npos : s_cntk[pr][k];      // (pos <= kp).sum()
          v = __ddiv_rn((double)c, (double)kp);
        }
        p.prs[o * p.nk + k] = v;
      }
    }
  }
}

__global__ void map_reduce_kernel(const MapParams p) {
  const int pr = threadIdx.x;
  if (pr >= p.P) return;
  double m = 0.0;
  int empty = 0;
  double prk[kMaxKappa];
  for (int k = 0; k < p.nk; ++k) prk[k] = 0.0;
  for (int q = 0; q < p.nq; ++q) {  // query order, like the reference's running sums
    const size_t o = (size_t)pr * p.nq + q;
    if (p.status[o] == RIR_MAP_EMPTY_OK) {
      ++empty;
      continue;
    }
    m = __dadd_rn(m, p.aps[o]);
    for (int k = 0; k < p.nk; ++k) prk[k] = __dadd_rn(prk[k], p.prs[o * p.nk + k]);
  }
  const double denom = (double)(p.nq - empty);
  p.map[pr] = __ddiv_rn(m, denom);  // nq == empty -> 0/0 = NaN; the host shim raises like Python's ZeroDivisionError
  for (int k = 0; k < p.nk; ++k) p.mpr[(size_t)pr * p.nk + k] = __ddiv_rn(prk[k], denom);
}

}  // namespace rir

using namespace rir;

static int compute_map_impl(const int32_t* ranks, const int32_t* positions, int nq, int64_t L, int64_t ld,
                            const int32_t* a_ids, const int32_t* a_off, const int32_t* b_ids, const int32_t* b_off,
                            const int32_t* c_ids, const int32_t* c_off, const int32_t* proto, int P, const int32_t* kappas,
                            int nk, double* map, double* aps, double* mpr, double* prs, int32_t* status, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(nq >= 1 && L >= 0 && ld >= L, "compute_map: bad shape nq=%d L=%lld ld=%lld", nq, (long long)L, (long long)ld);
  RIR_REQUIRE(ranks || L == 0, "compute_map: null ranks");
  RIR_REQUIRE(P >= 1 && P <= kMaxProto && proto, "compute_map: 1..%d protocols", kMaxProto);
  RIR_REQUIRE(nk >= 0 && nk <= kMaxKappa && (nk == 0 || kappas), "compute_map: 0..%d kappas", kMaxKappa);
  RIR_REQUIRE(map && aps && status && (nk == 0 || (mpr && prs)), "compute_map: null output");
  MapParams p;
  p.ranks = ranks; p.positions = positions; p.nq = nq; p.L = L; p.ld = ld;
  p.ids[0] = a_ids; p.off[0] = a_off; p.ids[1] = b_ids; p.off[1] = b_off; p.ids[2] = c_ids; p.off[2] = c_off;
  // proto / kappas are HOST arrays (tiny): they travel in the kernel parameter block
  for (int i = 0; i < kMaxProto; ++i) p.proto[i] = i < P ? proto[i] : 0;
  p.P = P;
  for (int i = 0; i < kMaxKappa; ++i) p.kappas[i] = i < nk ? kappas[i] : 0;
  p.nk = nk;
  p.map = map; p.aps = aps; p.mpr = mpr; p.prs = prs; p.status = status;
  cudaStream_t st = (cudaStream_t)stream;
  map_per_query_kernel<<<nq, kMapThreads, 0, st>>>(p);
  RIR_LAUNCH_OK();
  map_reduce_kernel<<<1, 32, 0, st>>>(p);
  RIR_LAUNCH_OK();
  return RIR_OK;
}

extern "C" int rir_compute_map(const int32_t* ranks, int nq, int64_t L, int64_t ld, const int32_t* a_ids,
                               const int32_t* a_off, const int32_t* b_ids, const int32_t* b_off, const int32_t* c_ids,
                               const int32_t* c_off, const int32_t* proto, int P, const int32_t* kappas, int nk,
                               double* map, double* aps, double* mpr, double* prs, int32_t* status, void* stream) {
  return compute_map_impl(ranks, nullptr, nq, L, ld, a_ids, a_off, b_ids, b_off, c_ids, c_off, proto, P, kappas, nk, map,
                          aps, mpr, prs, status, stream);
}

extern "C" int rir_compute_map_at(const int32_t* ranks, const int32_t* positions, int nq, int64_t L, int64_t ld,
                                  const int32_t* a_ids, const int32_t* a_off, const int32_t* b_ids, const int32_t* b_off,
                                  const int32_t* c_ids, const int32_t* c_off, const int32_t* proto, int P,
                                  const int32_t* kappas, int nk, double* map, double* aps, double* mpr, double* prs,
                                  int32_t* status, void* stream) {
  RIR_REQUIRE(positions != nullptr || L == 0, "compute_map_at: null positions");
  return compute_map_impl(ranks, positions, nq, L, ld, a_ids, a_off, b_ids, b_off, c_ids, c_off, proto, P, kappas, nk, map,
                          aps, mpr, prs, status, stream);
}

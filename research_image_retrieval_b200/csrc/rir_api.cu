// rir_api.cu — C-ABI glue: error state, device checks and the sim_topk orchestration
// (sample -> threshold -> scan -> select -> overflow fallback; see sim_topk.cuh).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "sim_topk.cuh"
#include "topk_select.cuh"

namespace rir {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_arch_ok[64];   // 0 unknown, 1 ok, -1 bad
static int g_sm_count[64];

int check_arch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available: librir.so has no CPU fallback");
    return RIR_E_ARCH;
  }
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_arch_ok[dev] == 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaGetDeviceProperties failed");
      return RIR_E_ARCH;
    }
    g_sm_count[dev] = prop.multiProcessorCount;
    g_arch_ok[dev] = (prop.major == 10) ? 1 : -1;
    if (g_arch_ok[dev] < 0)
      set_error("device %d is sm_%d%d; librir.so is built for sm_100a (B200) only", dev, prop.major, prop.minor);
  }
  if (g_arch_ok[dev] < 0) {
    set_error("current device is not sm_100 (B200); librir.so has no fallback path");
    return RIR_E_ARCH;
  }
  return RIR_OK;
}

// (kernel, device) -> largest dynamic shared memory size already granted with cudaFuncSetAttribute
int dyn_smem_granted(const void* kern, int dev, int bytes, bool record) {
  struct Entry { const void* k; int dev, bytes; };
  static Entry table[256];
  static int used = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (table[i].k == kern && table[i].dev == dev) {
      if (record && bytes > table[i].bytes) table[i].bytes = bytes;
      return table[i].bytes >= bytes ? 1 : 0;
    }
  if (record && used < 256) table[used++] = Entry{kern, dev, bytes};
  return 0;
}

bool pdl_enabled() {
  static const int on = getenv("RIR_PDL") ? atoi(getenv("RIR_PDL")) : 1;
  return on != 0;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_sm_count[dev] == 0) {  // workspace queries may come before the first launch
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) g_sm_count[dev] = n;
    else cudaGetLastError();
  }
  return g_sm_count[dev] > 0 ? g_sm_count[dev] : 148;
}

// ---------------------------------------------------------------------------------------------
// sim_topk planning (host): sample size, candidate capacity, workspace carve-up
// ---------------------------------------------------------------------------------------------
constexpr int kQueryGroup = 4096;      // queries processed per internal pass (bounds the workspace)
constexpr long long kScanAllMaxRows = 16384;
constexpr int kMaxKFilter = 8192;
constexpr int kMaxK = 16384;

// fixed workspace header (bytes): tau_score[4096] | tau_idx[4096] | cnt[4096] + barrier counter | ovf[4096]
constexpr size_t kHdrTauS = 0;
constexpr size_t kHdrTauI = kHdrTauS + (size_t)kQueryGroup * 4;
constexpr size_t kHdrCnt = kHdrTauI + (size_t)kQueryGroup * 4;
constexpr size_t kHdrGbar = kHdrCnt + (size_t)kQueryGroup * 4;  // 256 bytes: the grid-barrier counter
constexpr size_t kHdrOvf = kHdrGbar + 256;
constexpr size_t kHdrBytes = kHdrOvf + (size_t)kQueryGroup * 4;

struct SimPlan {
  bool scan_all;
  int nblk, sblk, cap;
  int group;  // queries per pass
  size_t off_tau_s, off_tau_i, off_cnt, off_ovf, off_sample, off_cand, total;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static bool make_plan(int nq, long long n, int k, SimPlan* pl) {
  if (nq < 0 || n < 1 || k < 1 || k > kMaxK || k > n) return false;
  pl->nblk = (int)((n + kSampleBlockRows - 1) / kSampleBlockRows);
  pl->group = nq < kQueryGroup ? (nq < 1 ? 1 : nq) : kQueryGroup;
  if (n <= kScanAllMaxRows) {
    pl->scan_all = true;
    pl->sblk = 0;
    pl->cap = (int)align_up((size_t)n, 32);
  } else {
    if (k > kMaxKFilter) return false;
    pl->scan_all = false;
    long long sblk = (long long)(0.02 * pl->nblk + 0.5);
    if (sblk < 32) sblk = 32;                        // >= 8192 sampled rows
    if (sblk > 148) sblk = 148;                      // one tile per SM is plenty: ~k*n/37888 survivors
    const long long need = (4ll * k + kSampleBlockRows - 1) / kSampleBlockRows;
    if (sblk < need) sblk = need;
    if (sblk > pl->nblk - 1) sblk = pl->nblk - 1;    // the last block may be partial: never sample it
    pl->sblk = (int)sblk;
    const double expected = (double)k * (double)n / ((double)sblk * kSampleBlockRows);
    long long cap = pow2_ceil_int((int)(3.0 * expected) + 1024 + sm_count() * kFusedTopT);
    if (cap < 2048) cap = 2048;
    pl->cap = (int)cap;
  }
  // The header (tau | cnt | grid-barrier counter | overflow flags) has a FIXED layout, sized for a full query group
  // whatever nq is: rir_sim_topk_workspace_init establishes "cnt = 0, barrier = 0, tau_score = sentinel" once and the
  // select kernel restores it at the end of every search, for every shape that re-uses the workspace.
  const size_t g = (size_t)align_up((size_t)pl->group, 128);
  pl->off_tau_s = kHdrTauS;
  pl->off_tau_i = kHdrTauI;
  pl->off_cnt = kHdrCnt;
  pl->off_ovf = kHdrOvf;
  size_t o = kHdrBytes;
  pl->off_sample = o; o = align_up(o + g * (size_t)pl->sblk * kSampleBlockRows * 4, 256);
  pl->off_cand = o;  o = align_up(o + g * (size_t)pl->cap * 8, 256);
  pl->total = o;
  return true;
}

}  // namespace rir

using namespace rir;

// optional profiling hook (bench.py's roofline): while armed, every full-scan launch of THIS host thread is bracketed
// by a pair of CUDA events from an internal pool — one pair per launch, so searches that run as several query groups
// (nq > 4096) are measured completely.
struct ScanEventPool {
  bool armed = false;
  bool paused = false;          // armed but not recording (sampled profiling: rir_profile_scan_pause)
  std::vector<cudaEvent_t> ev;  // pairs: ev[2i] start, ev[2i+1] stop
  size_t used = 0;              // events handed out since rir_profile_scan_begin
};
static thread_local ScanEventPool g_scan_ev;

static cudaEvent_t scan_event_next() {
  ScanEventPool& P = g_scan_ev;
  if (!P.armed || P.paused) return nullptr;
  if (P.used == P.ev.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    P.ev.push_back(e);
  }
  return P.ev[P.used++];
}

extern "C" int rir_profile_scan_begin(void) {
  g_scan_ev.armed = true;
  g_scan_ev.paused = false;
  g_scan_ev.used = 0;
  return RIR_OK;
}

extern "C" int rir_profile_scan_pause(int paused) {
  g_scan_ev.paused = paused != 0;
  return RIR_OK;
}

extern "C" int rir_profile_scan_end(float* ms_out, int cap, int* n_launches) {
  ScanEventPool& P = g_scan_ev;
  P.armed = false;
  const size_t pairs = P.used / 2;
  if (n_launches) *n_launches = (int)pairs;
  for (size_t i = 0; i < pairs; ++i) {
    RIR_CUDA_OK(cudaEventSynchronize(P.ev[2 * i + 1]));
    float ms = 0.f;
    RIR_CUDA_OK(cudaEventElapsedTime(&ms, P.ev[2 * i], P.ev[2 * i + 1]));
    if (ms_out && (int)i < cap) ms_out[i] = ms;
  }
  P.used = 0;
  return RIR_OK;
}

// optional development hook: event timeline of the tcgen05 scan kernels launched by this host thread
static thread_local unsigned long long* g_timeline = nullptr;
static thread_local int g_timeline_cap = 0;

extern "C" int rir_profile_timeline(void* dev_buf, int cap_events) {
  g_timeline = reinterpret_cast<unsigned long long*>(dev_buf);
  g_timeline_cap = dev_buf ? cap_events : 0;
  return RIR_OK;
}

extern "C" int rir_version(void) { return RIR_VERSION; }
extern "C" const char* rir_last_error(void) { return g_err; }
extern "C" int rir_device_check(void) { return check_arch(); }

static int elem_size(int dtype) { return dtype == RIR_BF16 ? 2 : (dtype == RIR_FP8E4M3 ? 1 : (dtype == RIR_F32 ? 4 : 0)); }

extern "C" size_t rir_sim_topk_workspace(int nq, int64_t n_local, int d, int k, int dtype) {
  SimPlan pl;
  if (elem_size(dtype) == 0 || d < 1) return 0;
  long long kk = k;
  if (kk > n_local) kk = n_local;  // the tail is padded with (-inf, -1)
  if (!make_plan(nq, n_local, (int)kk, &pl)) return 0;
  return pl.total;
}

extern "C" int rir_sim_topk_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "workspace_init: workspace must be non-null and 256-byte aligned");
  RIR_REQUIRE(workspace_bytes >= kHdrBytes, "workspace_init: workspace of %zu B is smaller than its %zu B header",
              workspace_bytes, kHdrBytes);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  RIR_CUDA_OK(cudaMemsetAsync(ws + kHdrTauS, 0xFF, (size_t)kQueryGroup * 4, st));              // kTauUnset
  RIR_CUDA_OK(cudaMemsetAsync(ws + kHdrCnt, 0, kHdrBytes - kHdrCnt, st));                      // cnt | barrier | ovf
  return RIR_OK;
}

// ---------------------------------------------------------------------------------------------
// asynchronous exchange (RIR_EXCHANGE_ASYNC): per inbox, a side stream + events so that the merge of search e runs
// next to the scan of search e+1 instead of in front of it
// ---------------------------------------------------------------------------------------------
struct AsyncExchange {
  const void* inbox;
  int dev;
  cudaStream_t side;
  cudaEvent_t sel_done[2], merge_done[2];
  uint32_t epoch[2];  // epoch whose merge merge_done[b] stands for (0 = none yet)
};
static std::mutex g_ax_mu;
static std::vector<AsyncExchange*> g_ax;

static AsyncExchange* ax_find(const void* inbox, bool create) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_ax_mu);
  for (AsyncExchange* a : g_ax)
    if (a->inbox == inbox && a->dev == dev) return a;
  if (!create) return nullptr;
  AsyncExchange* a = new AsyncExchange();
  a->inbox = inbox;
  a->dev = dev;
  a->epoch[0] = a->epoch[1] = 0u;
  bool ok = cudaStreamCreateWithFlags(&a->side, cudaStreamNonBlocking) == cudaSuccess;
  for (int b = 0; b < 2 && ok; ++b)
    ok = cudaEventCreateWithFlags(&a->sel_done[b], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&a->merge_done[b], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    delete a;
    return nullptr;
  }
  g_ax.push_back(a);
  return a;
}

static void ax_drop(const void* inbox) {
  std::lock_guard<std::mutex> lock(g_ax_mu);
  for (size_t i = 0; i < g_ax.size(); ++i)
    if (g_ax[i]->inbox == inbox) {
      AsyncExchange* a = g_ax[i];
      cudaStreamSynchronize(a->side);
      for (int b = 0; b < 2; ++b) {
        cudaEventDestroy(a->sel_done[b]);
        cudaEventDestroy(a->merge_done[b]);
      }
      cudaStreamDestroy(a->side);
      delete a;
      g_ax.erase(g_ax.begin() + i);
      return;
    }
}

// A rank may only PUBLISH epoch e (its select kernel stores into the peers' inboxes) once its own merge of epoch e-1 has
// read every list: a peer needs this rank's epoch-e lists before it can move on to epoch e+1 and overwrite inbox rows of
// parity (e-1) & 1.  With the merge in stream order that holds by itself; a merge on the side stream is waited for
// HERE, by the stream (between scan and select: the scan still overlaps it) — not by spinning CTAs, which would hold
// the SM resources the merge kernel needs (a 1,024-query select fills the register files: measured dead-lock).
static int wait_previous_async_merge(const Exchange& ex, cudaStream_t st) {
  AsyncExchange* a = ax_find(ex.inbox[ex.rank], false);
  if (a == nullptr) return RIR_OK;
  const int pb = (int)((ex.epoch - 1u) & 1u);
  if (a->epoch[pb] != 0u && a->epoch[pb] + 1u == ex.epoch) RIR_CUDA_OK(cudaStreamWaitEvent(st, a->merge_done[pb], 0));
  return RIR_OK;
}

static int sim_topk_impl(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale, int nq,
                         int64_t n_local, int d, int k, int64_t idx_offset, float* out_score, int32_t* out_idx,
                         void* workspace, size_t workspace_bytes, int path, void* stream, const Exchange* ex) {
  if (int e = check_arch()) return e;
  const bool ws_clean = (path & RIR_WS_CLEAN) != 0;  // the caller keeps the header invariant: no memset launches
  const bool ex_async = ex != nullptr && (path & RIR_EXCHANGE_ASYNC) != 0;
  path &= ~(RIR_WS_CLEAN | RIR_EXCHANGE_ASYNC);
  const int esz = elem_size(dtype);
  RIR_REQUIRE(esz != 0, "sim_topk: dtype must be RIR_F32, RIR_BF16 or RIR_FP8E4M3 (got %d)", dtype);
  RIR_REQUIRE(dtype != RIR_F32 || path != RIR_PATH_MMA, "sim_topk: fp32 descriptors run on the stream path only");
  RIR_REQUIRE(nq >= 0 && n_local >= 0 && d >= 1 && k >= 1, "sim_topk: bad shape nq=%d n=%lld d=%d k=%d", nq,
              (long long)n_local, d, k);
  RIR_REQUIRE(k <= kMaxK, "sim_topk: k=%d exceeds %d", k, kMaxK);
  RIR_REQUIRE(((size_t)d * esz) % 16 == 0, "sim_topk: row size %zu B must be a multiple of 16 (pad d with zeros)",
              (size_t)d * esz);
  if (nq == 0) return RIR_OK;
  RIR_REQUIRE(out_score && out_idx, "sim_topk: null output");
  RIR_REQUIRE(n_local + idx_offset < (1ll << 31) && idx_offset >= 0, "sim_topk: global row index exceeds int32");
  RIR_REQUIRE(path >= RIR_PATH_AUTO && path <= RIR_PATH_EXACT, "sim_topk: bad path %d", path);
  cudaStream_t st = (cudaStream_t)stream;
  if (nq == 0) return RIR_OK;
  RIR_REQUIRE(Q, "sim_topk: null Q");
  RIR_REQUIRE((reinterpret_cast<uintptr_t>(Q) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
              "sim_topk: Q and X must be 16-byte aligned");

  RIR_REQUIRE(X, "sim_topk: null X");
  RIR_REQUIRE(n_local >= 1, "sim_topk: empty shard (n_local == 0) - give every rank at least one row");
  const int k_req = k;
  if (ex && k > n_local) k = (int)n_local;  // a shard shorter than k publishes what it has, padded to k_req
  RIR_REQUIRE(k <= n_local, "sim_topk: k=%d exceeds the shard size %lld (clamp k on the host)", k, (long long)n_local);
  Exchange exl{};
  if (ex) {
    exl = *ex;
    exl.k_push = k_req;
    exl.fold = 0;
    ex = &exl;
  }

  SimPlan pl;
  if (!make_plan(nq, n_local, k, &pl)) {
    set_error("sim_topk: unsupported combination k=%d n=%lld (k <= %d, or n <= %lld for a full ranking)", k,
              (long long)n_local, kMaxKFilter, kScanAllMaxRows);
    return RIR_E_ARG;
  }
  if (path == RIR_PATH_EXACT) {
    SimParams p{};
    p.Q = Q; p.X = X; p.q_scale = q_scale; p.x_scale = x_scale;
    p.nq = nq; p.q0 = 0; p.n = n_local; p.d = d; p.row_bytes = d * esz;
    if (ex) p.ex = *ex;
    if (ex)
      if (int e = wait_previous_async_merge(*ex, st)) return e;
    if (int e = launch_exact_scan(p, dtype, nq, k, idx_offset, out_score, out_idx, nullptr, st)) return e;
    if (ex) return launch_merge_exchange(*ex, nq, k_req, out_score, out_idx, st);
    return RIR_OK;
  }
  RIR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "sim_topk: workspace must be non-null and 256-byte aligned");
  if (workspace_bytes < pl.total) {
    set_error("sim_topk: workspace of %zu B is smaller than the required %zu B", workspace_bytes, pl.total);
    return RIR_E_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  // one query group whose select CTAs are all co-resident: the select kernel merges the peers' lists itself
  if (ex && !ex_async && nq <= pl.group && select_can_fold_merge(nq, ex->G, k_req)) exl.fold = 1;

  for (int g0 = 0; g0 < nq; g0 += pl.group) {
    const int gq = (nq - g0) < pl.group ? (nq - g0) : pl.group;
    SimParams p{};
    p.Q = reinterpret_cast<const uint8_t*>(Q) + (size_t)g0 * d * esz;
    p.X = X;
    p.q_scale = q_scale ? q_scale + g0 : nullptr;
    p.x_scale = x_scale;
    p.n = n_local; p.d = d; p.row_bytes = d * esz;
    p.nblk = pl.nblk; p.sblk = pl.sblk;
    p.sample_scores = reinterpret_cast<float*>(ws + pl.off_sample);
    p.sample_keys = reinterpret_cast<unsigned long long*>(ws + pl.off_sample);
    p.topt = 0;
    p.sample_m = 0;
    p.tau_score = reinterpret_cast<float*>(ws + pl.off_tau_s);
    p.tau_idx = reinterpret_cast<uint32_t*>(ws + pl.off_tau_i);
    p.cnt = reinterpret_cast<uint32_t*>(ws + pl.off_cnt);
    p.cand = reinterpret_cast<unsigned long long*>(ws + pl.off_cand);
    p.cap = pl.cap;
    uint32_t* ovf = reinterpret_cast<uint32_t*>(ws + pl.off_ovf);
    const bool use_stream = (path == RIR_PATH_STREAM) || dtype == RIR_F32 || (path == RIR_PATH_AUTO && gq <= 2 && !mma_can_fuse(gq, n_local, k));
    // (measured, 1M x 2048 bf16, 1 query: fused tcgen05 scan 1571 q/s vs stream 1500 q/s — same scan time, the fused
    //  launch saves the separate sample / threshold passes)

    auto run_pass = [&](int mode) -> int {
      p.mode = mode;
      if (use_stream) {
        for (int s0 = 0; s0 < gq; s0 += 8) {  // the stream kernel holds <= 8 queries in shared memory
          p.q0 = s0;
          p.nq = (gq - s0) < 8 ? (gq - s0) : 8;
          if (int e = launch_sim_stream(p, dtype, st)) return e;
        }
        p.q0 = 0;
        p.nq = gq;
        return RIR_OK;
      }
      p.q0 = 0;
      p.nq = gq;
      return launch_sim_mma(p, dtype, st);
    };

    p.nq = gq;
    p.k = k;
    p.timeline = g_timeline;
    p.timeline_cap = g_timeline_cap;
    if (ex) {
      p.ex = *ex;
      p.ex.q_base = g0;
    }
    p.gbar = reinterpret_cast<uint32_t*>(ws + kHdrGbar);
    auto scan_mark = [&]() -> int {  // one event of the armed profiling pool (no-op otherwise)
      if (cudaEvent_t e = scan_event_next()) RIR_CUDA_OK(cudaEventRecord(e, st));
      return RIR_OK;
    };
    // The header invariant (cnt = 0, barrier = 0, tau_score = sentinel) is restored by the select kernel at the end
    // of every search; a caller that initialised the workspace once (rir_sim_topk_workspace_init) and says so with
    // RIR_WS_CLEAN skips these launches.  The first group of a plain call establishes it here.
    if (!ws_clean && g0 == 0 && !pl.scan_all) {
      RIR_CUDA_OK(cudaMemsetAsync(ws + kHdrCnt, 0, kHdrOvf - kHdrCnt, st));
      RIR_CUDA_OK(cudaMemsetAsync(ws + kHdrTauS, 0xFF, (size_t)pl.group * sizeof(float), st));
    }
    bool done = false;
    if (pl.scan_all) {
      // every row is a candidate: slot == row, cnt = n set by the select kernel's launch parameters
      if (int e = scan_mark()) return e;
      if (int e = run_pass(kModeScanAll)) return e;
      if (int e = scan_mark()) return e;
      p.mode = kModeScanAll;
      done = true;
    } else if (!use_stream && mma_can_fuse(gq, n_local, k)) {
      // ONE launch: the first round of the scan is the sample, tau is derived inside the kernel (sim_topk_mma.cu)
      if (int e = scan_mark()) return e;
      const int e = run_pass(kModeFused);
      if (e == RIR_OK) {
        if (int e2 = scan_mark()) return e2;
        p.mode = kModeFused;
        done = true;
      } else if (e != RIR_E_NOFUSE) {
        return e;
      } else if (g_scan_ev.armed && !g_scan_ev.paused && g_scan_ev.used > 0) {
        --g_scan_ev.used;  // the cooperative launch was refused: take the three-launch route below
      }
    }
    if (!done) {
      // what the sample pass keeps: best key per consumer warp (stream) / best T keys per sample tile (tcgen05);
      // >= 2k kept keys per query keep tau tight; very large k falls back to the dense sample
      const int sms = sm_count();
      p.topt = 0;
      p.sample_m = 0;
      if (use_stream) {
        if (2 * k <= sms * 8 && (size_t)sms * 8 * 8 <= (size_t)pl.sblk * kSampleBlockRows * 4) {
          p.topt = 1;
          p.sample_m = sms * 8;
          RIR_CUDA_OK(cudaMemsetAsync(p.sample_keys, 0, (size_t)gq * p.sample_m * sizeof(unsigned long long), st));
        }
      } else {
        int T = 1;
        while (T < 8 && (long long)T * pl.sblk < 2ll * k) T <<= 1;
        if ((long long)T * pl.sblk >= 2ll * k && (size_t)T * 8 <= (size_t)kSampleBlockRows * 4) {
          p.topt = T;
          p.sample_m = T * pl.sblk;
        }
      }
      if (int e = run_pass(kModeSample)) return e;
      if (int e = launch_sample_threshold(p, gq, k, st)) return e;
      if (int e = scan_mark()) return e;
      if (int e = run_pass(kModeScanFilter)) return e;
      if (int e = scan_mark()) return e;
      p.mode = kModeScanFilter;
    }
    float* os = out_score + (size_t)g0 * k;
    int32_t* oi = out_idx + (size_t)g0 * k;
    if (ex && g0 == 0)
      if (int e = wait_previous_async_merge(*ex, st)) return e;
    if (int e = launch_final_select(p, dtype, gq, k, idx_offset, os, oi, ovf, st)) return e;
    if (!pl.scan_all && !select_handles_overflow(k)) {
      // very large k: queries whose candidate list overflowed are redone by the separate exact kernel (no-op otherwise)
      if (int e = launch_exact_scan(p, dtype, gq, k, idx_offset, os, oi, ovf, st)) return e;
    }
  }
  // sharded: every rank's lists are on their way into the inboxes; wait for all G of them and merge
  if (ex_async) {
    // on the inbox's side stream, behind this call's select kernels: the caller's stream goes straight on to the next
    // search; rir_exchange_join / rir_exchange_sync order consumers behind the merge
    AsyncExchange* a = ax_find(ex->inbox[ex->rank], true);
    if (a == nullptr) {
      set_error("sim_topk_sharded: could not create the side stream of the asynchronous exchange");
      return RIR_E_CUDA;
    }
    const int b = (int)(ex->epoch & 1u);
    RIR_CUDA_OK(cudaEventRecord(a->sel_done[b], st));
    RIR_CUDA_OK(cudaStreamWaitEvent(a->side, a->sel_done[b], 0));
    if (int e = launch_merge_exchange(*ex, nq, k_req, out_score, out_idx, a->side, true)) return e;
    RIR_CUDA_OK(cudaEventRecord(a->merge_done[b], a->side));
    a->epoch[b] = ex->epoch;
    return RIR_OK;
  }
  if (ex && !ex->fold) return launch_merge_exchange(*ex, nq, k_req, out_score, out_idx, st);
  return RIR_OK;
}

extern "C" int rir_sim_topk(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale,
                            int nq, int64_t n_local, int d, int k, int64_t idx_offset, float* out_score,
                            int32_t* out_idx, void* workspace, size_t workspace_bytes, int path, void* stream) {
  return sim_topk_impl(Q, X, dtype, q_scale, x_scale, nq, n_local, d, k, idx_offset, out_score, out_idx, workspace,
                       workspace_bytes, path, stream, nullptr);
}

// ---------------------------------------------------------------------------------------------
// sharded search over NVLink peer memory
// ---------------------------------------------------------------------------------------------
extern "C" size_t rir_exchange_bytes(int G, int nq_max, int k_max) {
  if (G < 1 || G > kMaxPeers || nq_max < 1 || k_max < 1) return 0;
  return exchange_bytes(G, nq_max, k_max);
}

extern "C" int rir_peer_alloc(size_t bytes, void** ptr) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: bad arguments");
  RIR_CUDA_OK(cudaMalloc(ptr, bytes));  // a plain cudaMalloc allocation: exportable with cudaIpcGetMemHandle
  RIR_CUDA_OK(cudaMemset(*ptr, 0, bytes));
  RIR_CUDA_OK(cudaDeviceSynchronize());
  return RIR_OK;
}

extern "C" int rir_peer_free(void* ptr) {
  if (ptr) {
    ax_drop(ptr);
    RIR_CUDA_OK(cudaFree(ptr));
  }
  return RIR_OK;
}

extern "C" int rir_exchange_join(const void* own_inbox, uint32_t epoch, void* stream) {
  AsyncExchange* a = ax_find(own_inbox, false);
  if (a == nullptr) return RIR_OK;  // no asynchronous search was issued on this inbox
  const int b = (int)(epoch & 1u);
  RIR_REQUIRE(a->epoch[b] == epoch, "exchange_join: epoch %u is not outstanding (parity holds %u)", epoch, a->epoch[b]);
  RIR_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, a->merge_done[b], 0));
  return RIR_OK;
}

extern "C" int rir_exchange_sync(const void* own_inbox, uint32_t epoch) {
  AsyncExchange* a = ax_find(own_inbox, false);
  if (a == nullptr) return RIR_OK;
  const int b = (int)(epoch & 1u);
  RIR_REQUIRE(a->epoch[b] == epoch, "exchange_sync: epoch %u is not outstanding (parity holds %u)", epoch, a->epoch[b]);
  RIR_CUDA_OK(cudaEventSynchronize(a->merge_done[b]));
  return RIR_OK;
}

extern "C" int rir_peer_export(void* ptr, void* handle64) {
  RIR_REQUIRE(ptr && handle64, "peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  RIR_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return RIR_OK;
}

extern "C" int rir_peer_open(const void* handle64, void** ptr) {
  RIR_REQUIRE(ptr && handle64, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  RIR_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RIR_OK;
}

extern "C" int rir_peer_close(void* ptr) {
  if (ptr) RIR_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return RIR_OK;
}

extern "C" int rir_sim_topk_sharded(const void* Q, const void* X, int dtype, const float* q_scale, const float* x_scale,
                                    int nq, int64_t n_local, int d, int k, int64_t idx_offset, float* out_score,
                                    int32_t* out_idx, void* workspace, size_t workspace_bytes, int path, void* stream,
                                    int G, int rank, uint32_t epoch, int nq_max, int k_max, void* const* inbox) {
  RIR_REQUIRE(G >= 1 && G <= kMaxPeers && rank >= 0 && rank < G, "sim_topk_sharded: bad rank %d of %d (<= %d ranks)", rank,
              G, kMaxPeers);
  RIR_REQUIRE(epoch != 0u, "sim_topk_sharded: epoch starts at 1");
  RIR_REQUIRE(nq >= 0 && nq <= nq_max && k >= 1 && k <= k_max, "sim_topk_sharded: nq=%d k=%d exceed the inbox (%d, %d)",
              nq, k, nq_max, k_max);
  RIR_REQUIRE(inbox != nullptr, "sim_topk_sharded: null inbox table");
  Exchange ex{};
  ex.G = G; ex.rank = rank; ex.epoch = epoch; ex.nq_max = nq_max; ex.k_max = k_max; ex.q_base = 0;
  for (int g = 0; g < G; ++g) {
    RIR_REQUIRE(inbox[g] != nullptr, "sim_topk_sharded: inbox of rank %d is null", g);
    ex.inbox[g] = reinterpret_cast<unsigned long long*>(inbox[g]);
  }
  if (nq == 0) return RIR_OK;
  return sim_topk_impl(Q, X, dtype, q_scale, x_scale, nq, n_local, d, k, idx_offset, out_score, out_idx, workspace,
                       workspace_bytes, path, stream, &ex);
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry point: the whole query path as ONE call (the reference call site hands over CPU tensors:
// iris_evaluate.py:378-386 works on `query_features` that come out of extract_vectors as CPU fp32)
// ---------------------------------------------------------------------------------------------
struct HostPlan {
  size_t off_qf, off_qp, off_qs, off_sc, off_ix, off_sim, total;
};
static bool make_host_plan(int nq, int64_t n_local, int d, int k, int dtype, HostPlan* hp) {
  const int esz = elem_size(dtype);
  if (esz == 0 || nq < 1 || d < 1 || k < 1) return false;
  long long kk = k;
  if (kk > n_local) kk = n_local;
  const size_t sim = rir_sim_topk_workspace(nq, n_local, d, (int)kk, dtype);
  if (sim == 0) return false;
  size_t o = 0;
  hp->off_sim = o; o = align_up(o + sim, 256);
  hp->off_qf = o;  o = align_up(o + (size_t)nq * d * 4, 256);
  hp->off_qp = o;  o = align_up(o + (size_t)nq * d * esz, 256);
  hp->off_qs = o;  o = align_up(o + (size_t)nq * 4, 256);
  hp->off_sc = o;  o = align_up(o + (size_t)nq * k * 4, 256);
  hp->off_ix = o;  o = align_up(o + (size_t)nq * k * 4, 256);
  hp->total = o;
  return true;
}

extern "C" size_t rir_search_host_workspace(int nq, int64_t n_local, int d, int k, int dtype) {
  HostPlan hp;
  return make_host_plan(nq, n_local, d, k, dtype, &hp) ? hp.total : 0;
}

extern "C" int rir_search_host(const float* q_host, const void* X, int dtype, const float* x_scale, int nq,
                               int64_t n_local, int d, int k, int64_t idx_offset, float* out_score_host,
                               int32_t* out_idx_host, void* workspace, size_t workspace_bytes, int path, void* stream,
                               int G, int rank, uint32_t epoch, int nq_max, int k_max, void* const* inbox) {
  if (int e = check_arch()) return e;
  RIR_REQUIRE(q_host && out_score_host && out_idx_host, "search_host: null host buffer");
  RIR_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "search_host: workspace must be 256-byte aligned");
  if (nq == 0) return RIR_OK;
  HostPlan hp;
  if (!make_host_plan(nq, n_local, d, k, dtype, &hp)) {
    set_error("search_host: unsupported shape nq=%d n=%lld d=%d k=%d dtype=%d", nq, (long long)n_local, d, k, dtype);
    return RIR_E_ARG;
  }
  if (workspace_bytes < hp.total) {
    set_error("search_host: workspace of %zu B is smaller than the required %zu B", workspace_bytes, hp.total);
    return RIR_E_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  float* qf = reinterpret_cast<float*>(ws + hp.off_qf);
  void* qp = ws + hp.off_qp;
  float* qs = reinterpret_cast<float*>(ws + hp.off_qs);
  float* sc = reinterpret_cast<float*>(ws + hp.off_sc);
  int32_t* ix = reinterpret_cast<int32_t*>(ws + hp.off_ix);
  // Pinned (page-locked) host buffers are device-accessible under UVA: the pack kernel then reads the queries
  // straight from host memory and the select / merge kernel writes the top-k straight into the caller's buffers —
  // no copy launches on the critical path.  Pageable buffers take the cudaMemcpyAsync route.
  auto device_view = [](const void* host) -> void* {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
  };
  // bit 0: queries read in place, bit 1: results written in place (development override RIR_HOST_ZERO_COPY)
  static const int zero_copy = getenv("RIR_HOST_ZERO_COPY") ? atoi(getenv("RIR_HOST_ZERO_COPY")) : 2;
  const float* q_dev_view =
      ((zero_copy & 1) && dtype != RIR_F32) ? reinterpret_cast<const float*>(device_view(q_host)) : nullptr;
  float* sc_dev_view = (zero_copy & 2) ? reinterpret_cast<float*>(device_view(out_score_host)) : nullptr;
  int32_t* ix_dev_view = (zero_copy & 2) ? reinterpret_cast<int32_t*>(device_view(out_idx_host)) : nullptr;
  const bool direct_out = sc_dev_view != nullptr && ix_dev_view != nullptr;
  if (direct_out) {
    sc = sc_dev_view;
    ix = ix_dev_view;
  }
  const float* q_src = q_dev_view;
  if (q_src == nullptr) {
    RIR_CUDA_OK(cudaMemcpyAsync(qf, q_host, (size_t)nq * d * 4, cudaMemcpyHostToDevice, st));
    q_src = qf;
  }
  const void* Q = q_src;
  const float* q_scale = nullptr;
  if (dtype != RIR_F32) {
    if (int e = rir_pack_descriptors(q_src, nq, d, dtype, qp, qs, stream)) return e;
    Q = qp;
    if (dtype == RIR_FP8E4M3) q_scale = qs;
  }
  const size_t sim_bytes = hp.off_qf;  // the region in front of the staging buffers
  int rc;
  if (G > 1)
    rc = rir_sim_topk_sharded(Q, X, dtype, q_scale, x_scale, nq, n_local, d, k, idx_offset, sc, ix, ws + hp.off_sim,
                              sim_bytes, path, stream, G, rank, epoch, nq_max, k_max, inbox);
  else
    rc = rir_sim_topk(Q, X, dtype, q_scale, x_scale, nq, n_local, d, k, idx_offset, sc, ix, ws + hp.off_sim, sim_bytes,
                      path, stream);
  if (rc) return rc;
  if (!direct_out) {
    RIR_CUDA_OK(cudaMemcpyAsync(out_score_host, sc, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    RIR_CUDA_OK(cudaMemcpyAsync(out_idx_host, ix, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
  }
  return RIR_OK;
}

"""Similarity search: exact Q·Xᵀ top-k, sharded merge and alpha query expansion on librir.so kernels.

Drop-in for the reference's similarity + ranking call site

    similarity = torch.mm(query_features, gallery_features.t()).cpu().numpy()      iris_evaluate.py:383
    ranks = np.argsort(-similarity, axis=1)                                         iris_evaluate.py:386

and its top-k precedent `compute_similarity` + `torch.topk(similarity, k)`
(reference/manus/7_AdaptiveHybridModel/modified/adaptive_hybrid_retrieval_complete.py:11-16, 428).

  rank(q, g, k=None)        -> int64 ndarray [k_or_N, nq]   (column per query — the layout compute_map documents,
                                                             utils/evaluate.py:49)
  Database / sim_topk       -> resident bf16 / fp8 / fp32 descriptor shard + exact top-k
  ShardedDatabase           -> row-sharded over the ranks of a torch.distributed group, allgather + merge_topk
  alpha_query_expansion     -> q' = L2(q + sum_j max(s_j,0)^alpha x_j), then search again

Order rule everywhere: descending score, ties -> ascending database index.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import PATHS, RIR_BF16, RIR_EXCHANGE_ASYNC, RIR_F32, RIR_FP8E4M3, RIR_WS_CLEAN

_DTYPES = {"bf16": RIR_BF16, "fp8": RIR_FP8E4M3, "fp32": RIR_F32}
_TORCH_DT = {"bf16": torch.bfloat16, "fp32": torch.float32}
if hasattr(torch, "float8_e4m3fn"):
    _TORCH_DT["fp8"] = torch.float8_e4m3fn

MAX_K_FILTER = 8192          # k limit for shards larger than MAX_FULL_RANK rows (include/rir.h)
MAX_FULL_RANK = 16384        # shards up to this many rows can be ranked completely


# ----------------------------------------------------------------------------------------------
# host-side helpers (pure logic; unit-tested on CPU)
# ----------------------------------------------------------------------------------------------
def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Row range [lo, hi) of shard `rank`: ceil(n / world) rows each, the last one shorter."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad shard {rank}/{world}")
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def pad_dim(d: int, dtype: str) -> int:
    """Descriptor rows must be a multiple of 16 bytes (TMA / 128-bit loads); zero padding leaves dot products unchanged."""
    esz = {"bf16": 2, "fp8": 1, "fp32": 4}[dtype]
    per = 16 // esz
    return -(-d // per) * per


def clamp_k(k: Optional[int], n: int) -> int:
    if k is None:
        k = n
    k = int(k)
    if k < 1:
        raise ValueError("k must be >= 1")
    return min(k, n)


def check_k_supported(k: int, n_local: int) -> None:
    if n_local > MAX_FULL_RANK and k > MAX_K_FILTER:
        raise ValueError(f"k={k} is not supported for a shard of {n_local} rows: k <= {MAX_K_FILTER}, "
                         f"or a full ranking for shards of at most {MAX_FULL_RANK} rows")


def merge_topk_host(scores: np.ndarray, idx: np.ndarray, k: int):
    """Reference semantics of the cross-shard merge on [G, nq, k] arrays (numpy; used by the CPU tests of the
    multi-rank plumbing — the product path is rir_merge_topk)."""
    G, nq, kk = scores.shape
    s = np.transpose(scores, (1, 0, 2)).reshape(nq, G * kk)
    i = np.transpose(idx, (1, 0, 2)).reshape(nq, G * kk)
    s = np.where(i < 0, -np.inf, s)
    big = np.iinfo(np.int64).max
    order = np.lexsort((np.where(i < 0, big, i.astype(np.int64)), -s.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(s, order, 1), np.take_along_axis(i, order, 1)


# ----------------------------------------------------------------------------------------------
# packed descriptor shards
# ----------------------------------------------------------------------------------------------
def pack_descriptors(v: torch.Tensor, dtype: str = "bf16"):
    """fp32 [n, d] CUDA descriptors -> (rows in the search dtype [n, d_pad], per-row scale or None)."""
    if dtype not in _DTYPES:
        raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
    if not v.is_cuda:
        raise TypeError("descriptors must be on the GPU (no CPU path)")
    v = v.float().contiguous()
    n, d = v.shape
    dp = pad_dim(d, dtype)
    if dp != d:
        v = torch.nn.functional.pad(v, (0, dp - d))
    if dtype == "fp32":
        return v, None
    lib = _lib.load()
    if dtype == "bf16":
        out = torch.empty((n, dp), dtype=torch.bfloat16, device=v.device)
        scale = None
    else:
        out = torch.empty((n, dp), dtype=torch.uint8, device=v.device)
        scale = torch.empty(n, dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(lib.rir_pack_descriptors(v.data_ptr(), n, dp, _DTYPES[dtype], out.data_ptr(),
                                            None if scale is None else scale.data_ptr(), _lib.stream_ptr()))
    return out, scale


class PendingQuery:
    """Handle of an enqueued search whose results are not yet valid (Database.query_host_async,
    ShardedDatabase.search_async, HostQueryPipeline.submit).  `.result()` blocks the HOST until they are and returns
    (scores, idx); `.wait()` makes the current CUDA stream wait instead (device-side consumers)."""

    def __init__(self, event, sc, ix, keep=None, join=None, sync=None):
        self._event, self._sc, self._ix = event, sc, ix
        self._keep = keep              # keeps input buffers alive until the work that reads them has run
        self._join, self._sync = join, sync   # asynchronous exchange: the merge lives on the library's side stream

    def done(self) -> bool:
        return self._event.query() if self._sync is None else False

    def wait(self):
        if self._join is not None:
            self._join()
        else:
            torch.cuda.current_stream().wait_event(self._event)
        return self._sc, self._ix

    def result(self):
        if self._sync is not None:
            self._sync()
        else:
            self._event.synchronize()
        return self._sc, self._ix


class Database:
    """One resident shard of database descriptors in the layout the search kernels read.

    rows: [n_local, d] bf16 / uint8(fp8 e4m3) / fp32, row-major, 16-byte-multiple rows; scale: per-row fp32 or None;
    idx_offset: global index of local row 0."""

    def __init__(self, rows: torch.Tensor, scale: Optional[torch.Tensor], dtype: str, idx_offset: int = 0,
                 d_logical: Optional[int] = None, rescore_rows: Optional[torch.Tensor] = None):
        self.rows, self.scale, self.dtype, self.idx_offset = rows, scale, dtype, int(idx_offset)
        self.n, self.d = rows.shape
        self.d_logical = d_logical or self.d
        self.rescore_rows = rescore_rows  # bf16 copy of the shard used to re-score fp8 candidates
        self._ws = {}   # (kind, stream) -> initialised workspace

    @classmethod
    def from_descriptors(cls, v: torch.Tensor, dtype: str = "bf16", idx_offset: int = 0, normalize: bool = False,
                         rescore: bool = False):
        """rescore=True (fp8 only) also keeps a bf16 copy: the fp8 scan returns 2k+16 candidates which are re-scored
        against it, so the final top-k meets the 5e-3 bar against an fp32 rescore."""
        if normalize:
            from .pooling import l2n
            v = l2n(v.float().contiguous())
        rows, scale = pack_descriptors(v, dtype)
        rs = None
        if rescore and dtype == "fp8":
            rs, _ = pack_descriptors(v, "bf16")
            if rs.shape[1] != rows.shape[1]:  # fp8 rows pad d to 16, bf16 to 8: use the common padded width
                rs = torch.nn.functional.pad(rs, (0, rows.shape[1] - rs.shape[1]))
        return cls(rows, scale, dtype, idx_offset, d_logical=v.shape[1], rescore_rows=rs)

    def pack_queries(self, q: torch.Tensor):
        """fp32 [nq, d_logical] -> the shard's dtype (padded like the rows)."""
        if q.shape[1] != self.d_logical:
            raise ValueError(f"query dimension {q.shape[1]} != database dimension {self.d_logical}")
        return pack_descriptors(q.to(self.rows.device), self.dtype)

    def _workspace(self, kind: str, need: int) -> torch.Tensor:
        """Scratch for the current CUDA stream, initialised once (rir_sim_topk_workspace_init): searches then pass
        RIR_WS_CLEAN and issue no memset launches — their select kernel leaves the header ready for the next search.
        One workspace per (kind, stream), so searches on different streams never share scratch."""
        key = (kind, torch.cuda.current_stream(self.rows.device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=self.rows.device)
            with torch.cuda.device(self.rows.device):
                _lib.check(_lib.load().rir_sim_topk_workspace_init(ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
            self._ws[key] = ws
        return ws

    def _drop_workspaces(self):
        """After a failed call the header state is unknown: re-initialise on next use."""
        self._ws.clear()

    def workspace(self, nq: int, k: int) -> torch.Tensor:
        need = _lib.load().rir_sim_topk_workspace(nq, self.n, self.d, k, _DTYPES[self.dtype])
        if need == 0 and not (nq == 0):
            raise ValueError(f"unsupported search shape nq={nq} n={self.n} d={self.d} k={k}")
        return self._workspace("sim", need)

    def search(self, q_rows: torch.Tensor, q_scale: Optional[torch.Tensor], k: int, path: str = "auto", out=None):
        """Exact local top-k.  Returns (scores [nq, k] fp32, idx [nq, k] int32 GLOBAL indices) on the GPU
        (`out` = (scores, idx) tensors to fill instead of allocating)."""
        try:
            return sim_topk(q_rows, self.rows, k, dtype=self.dtype, q_scale=q_scale, x_scale=self.scale,
                            idx_offset=self.idx_offset, path=path, workspace=self.workspace(q_rows.shape[0], k),
                            ws_clean=True, out=out)
        except _lib.RirError:
            self._drop_workspaces()
            raise

    def query_host_async(self, q_host: torch.Tensor, k: int, out=None, path: str = "auto", _exchange=None):
        """Enqueue the host-facing call and return at once: fp32 CPU queries [nq, d] (pinned for an asynchronous copy)
        -> top-k written into host buffers.  One C-ABI call (rir_search_host) enqueues H2D -> pack -> search -> D2H on
        the current stream.  Returns a PendingQuery; `.result()` waits for THIS call only and returns
        (scores [nq, k] fp32, idx [nq, k] int32) CPU tensors.  A serving loop keeps two batches in flight (distinct
        `out` buffers) so the device never idles between them."""
        if q_host.is_cuda or q_host.dtype != torch.float32 or q_host.dim() != 2 or not q_host.is_contiguous():
            raise TypeError("query_host expects a contiguous float32 CPU tensor [nq, d]")
        nq, d = q_host.shape
        if d != self.d or self.rescore_rows is not None:
            raise ValueError("query_host needs d == the packed row width and no rescoring copy; use query()")
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32).pin_memory(), torch.empty((nq, k), dtype=torch.int32).pin_memory())
        sc, ix = out
        lib = _lib.load()
        dt = _DTYPES[self.dtype]
        need = lib.rir_search_host_workspace(nq, self.n, d, k, dt)
        if need == 0:
            raise ValueError(f"unsupported search shape nq={nq} n={self.n} d={d} k={k}")
        ws = self._workspace("host", need)
        ex = _exchange or (1, 0, 0, 0, 0, None)
        with torch.cuda.device(self.rows.device):
            try:
                _lib.check(lib.rir_search_host(q_host.data_ptr(), self.rows.data_ptr(), dt,
                                               None if self.scale is None else self.scale.data_ptr(), nq, self.n, d, k,
                                               self.idx_offset, sc.data_ptr(), ix.data_ptr(), ws.data_ptr(),
                                               ws.numel(), PATHS[path] | RIR_WS_CLEAN, _lib.stream_ptr(), *ex))
            except _lib.RirError:
                self._drop_workspaces()
                raise
            done = torch.cuda.Event()
            done.record()
        return PendingQuery(done, sc, ix, keep=q_host)

    def query_host(self, q_host: torch.Tensor, k: int, out=None, path: str = "auto", _exchange=None):
        """query_host_async(...).result(): the synchronous host-facing call (`out` to re-use pinned result buffers)."""
        return self.query_host_async(q_host, k, out=out, path=path, _exchange=_exchange).result()

    def query(self, q: torch.Tensor, k: int, path: str = "auto"):
        """fp32 queries [nq, d_logical] -> top-k.  With a rescoring copy: fp8 scan for 2k+16 candidates, then a bf16
        re-score of those rows (rir_rescore_topk) decides the final k."""
        q = q.to(self.rows.device).float()
        qr, qs = self.pack_queries(q)
        if self.rescore_rows is None:
            return self.search(qr, qs, k, path=path)
        k_in = min(self.n, 2 * k + 16, 8192)
        _, cand = self.search(qr, qs, k_in, path=path)
        qb, _ = pack_descriptors(q, "bf16")
        if qb.shape[1] != self.rescore_rows.shape[1]:
            qb = torch.nn.functional.pad(qb, (0, self.rescore_rows.shape[1] - qb.shape[1]))
        sc = torch.empty((q.shape[0], k), dtype=torch.float32, device=q.device)
        ix = torch.empty((q.shape[0], k), dtype=torch.int32, device=q.device)
        with torch.cuda.device(q.device):
            _lib.check(_lib.load().rir_rescore_topk(qb.data_ptr(), self.rescore_rows.data_ptr(), RIR_BF16, None, None,
                                                    q.shape[0], self.n, self.idx_offset, self.rescore_rows.shape[1],
                                                    cand.data_ptr(), k_in, k, sc.data_ptr(), ix.data_ptr(),
                                                    _lib.stream_ptr()))
        return sc, ix


def sim_topk(Q: torch.Tensor, X: torch.Tensor, k: int, dtype: str = "bf16", q_scale=None, x_scale=None,
             idx_offset: int = 0, path: str = "auto", workspace: Optional[torch.Tensor] = None, out=None,
             ws_clean: bool = False):
    """Thin wrapper over rir_sim_topk (include/rir.h).  Q [nq, d], X [n, d] already in the search dtype.
    ws_clean: `workspace` was initialised with rir_sim_topk_workspace_init (RIR_WS_CLEAN, see include/rir.h)."""
    lib = _lib.load()
    if dtype not in _DTYPES:
        raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
    if path not in PATHS:
        raise ValueError(f"path must be one of {sorted(PATHS)}")
    if not (Q.is_cuda and X.is_cuda):
        raise TypeError("Q and X must be CUDA tensors (no CPU path)")
    if Q.dim() != 2 or X.dim() != 2 or Q.shape[1] != X.shape[1]:
        raise ValueError(f"shape mismatch: Q {tuple(Q.shape)} X {tuple(X.shape)}")
    if not (Q.is_contiguous() and X.is_contiguous()):
        raise ValueError("Q and X must be contiguous row-major")
    nq, d = Q.shape
    n = X.shape[0]
    if not (1 <= k <= n):
        raise ValueError(f"k={k} must be in [1, n={n}]")
    check_k_supported(k, n)
    if out is None:
        sc = torch.empty((nq, k), dtype=torch.float32, device=X.device)
        ix = torch.empty((nq, k), dtype=torch.int32, device=X.device)
    else:
        sc, ix = out
    if workspace is None and path != "exact":
        need = lib.rir_sim_topk_workspace(nq, n, d, k, _DTYPES[dtype])
        workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=X.device)
    with torch.cuda.device(X.device):
        _lib.check(lib.rir_sim_topk(Q.data_ptr(), X.data_ptr(), _DTYPES[dtype],
                                    None if q_scale is None else q_scale.data_ptr(),
                                    None if x_scale is None else x_scale.data_ptr(), nq, n, d, k, int(idx_offset),
                                    sc.data_ptr(), ix.data_ptr(),
                                    None if workspace is None else workspace.data_ptr(),
                                    0 if workspace is None else workspace.numel(),
                                    PATHS[path] | (RIR_WS_CLEAN if ws_clean and workspace is not None else 0),
                                    _lib.stream_ptr()))
    return sc, ix


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k: Optional[int] = None):
    """[G, nq, k] per-shard results -> global (scores [nq, k], idx [nq, k]) via rir_merge_topk."""
    G, nq, kk = scores.shape
    k = kk if k is None else k
    if k != kk:
        raise ValueError("merge_topk keeps k == the per-shard k")
    scores = scores.contiguous()
    idx = idx.contiguous()
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(_lib.load().rir_merge_topk(scores.data_ptr(), idx.data_ptr(), G, nq, k, out_s.data_ptr(),
                                              out_i.data_ptr(), None, 0, _lib.stream_ptr()))
    return out_s, out_i


# ----------------------------------------------------------------------------------------------
# reference call-site drop-in
# ----------------------------------------------------------------------------------------------
def rank(query_features, gallery_features, k: Optional[int] = None, dtype: str = "fp32", normalize: bool = False,
         return_scores: bool = False, path: str = "auto", device=None):
    """`np.argsort(-torch.mm(q, g.t()), axis=1)` without the score matrix (iris_evaluate.py:383-386).

    Returns an int64 ndarray [k_or_N, nq] — column per query, the layout compute_map consumes
    (utils/evaluate.py:49).  dtype='fp32' keeps the reference arithmetic type; 'bf16'/'fp8' use the tensor-core path.
    normalize=True applies F.normalize to both sides first (iris_evaluate.py:379-380).

    Limits (include/rir.h): a FULL ranking (k=None) is produced for galleries of at most 16,384 rows (ROxford / RParis
    size); larger galleries (e.g. +1M distractors) take a top-k with k <= 8192 — for the revisited protocol over such a
    gallery use `revisited_map_full`, which needs no ranked list at all.  fp32 on a large gallery: candidates
    (2k + 64 per query) come from ONE bf16 tensor-core scan, their scores and final order from an fp32 re-score of
    those rows (rir_rescore_topk) — fp32 scores, exact up to bf16 ties at the over-fetch boundary — instead of
    ceil(nq / 8) CUDA-core passes over the fp32 rows."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    q = torch.as_tensor(query_features).to(device=device, dtype=torch.float32)
    g = torch.as_tensor(gallery_features).to(device=device, dtype=torch.float32)
    if normalize:
        from .pooling import l2n
        q, g = l2n(q.contiguous()), l2n(g.contiguous())
    n = g.shape[0]
    if n > MAX_FULL_RANK and (k is None or int(k) > MAX_K_FILTER):
        raise ValueError(f"a gallery of {n} rows is ranked to a top-k with k <= {MAX_K_FILTER} (a full ranking only up to "
                         f"{MAX_FULL_RANK} rows); for revisited mAP over the whole gallery use revisited_map_full()")
    if dtype == "fp32" and n > MAX_FULL_RANK and 2 * int(k) + 64 <= MAX_K_FILTER and path == "auto":
        k = clamp_k(k, n)
        k_in = 2 * k + 64
        dbb = Database.from_descriptors(g, "bf16")
        qr, _ = dbb.pack_queries(q)
        _, cand = dbb.search(qr, None, k_in)
        g32, q32 = pack_descriptors(g, "fp32")[0], pack_descriptors(q, "fp32")[0]
        sc = torch.empty((q.shape[0], k), dtype=torch.float32, device=device)
        ix = torch.empty((q.shape[0], k), dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.load().rir_rescore_topk(q32.data_ptr(), g32.data_ptr(), RIR_F32, None, None, q.shape[0], n, 0,
                                                    g32.shape[1], cand.data_ptr(), k_in, k, sc.data_ptr(), ix.data_ptr(),
                                                    _lib.stream_ptr()))
    else:
        db = Database.from_descriptors(g, dtype)
        k = clamp_k(k, db.n)
        qr, qs = db.pack_queries(q)
        sc, ix = db.search(qr, qs, k, path=path)
    ranks = ix.t().contiguous().to(torch.int64).cpu().numpy()
    if return_scores:
        return ranks, sc.t().contiguous().cpu().numpy()
    return ranks


# ----------------------------------------------------------------------------------------------
# row-sharded database over a torch.distributed group
# ----------------------------------------------------------------------------------------------
class ShardedDatabase:
    """Database rows sharded across the ranks of a process group (one process per GPU).

    Each rank searches its shard (global indices via idx_offset).  The one exchange step runs either
      * over NVLink peer memory (`enable_peer_exchange`): the select kernel stores the local top-k straight into every
        rank's inbox and a merge kernel waits for the G lists — rir_sim_topk_sharded, no collective call; or
      * as an all-gather of the packed [nq, k] candidates (NCCL on GPUs, gloo in the CPU plumbing tests) followed by
        rir_merge_topk on every rank."""

    def __init__(self, local: Database, group=None):
        import torch.distributed as dist
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._inbox = None       # this rank's inbox (device pointer, owned)
        self._peers = None       # ctypes array [G] of inbox pointers as mapped here
        self._opened = []        # IPC mappings to close
        self._epoch = 0
        self._nq_max = self._k_max = 0
        # rows over all shards (k is clamped to it: a list can never hold more real entries)
        self.n_global = int(local.n)
        if self.world > 1:
            t = torch.tensor([int(local.n)], dtype=torch.int64, device=local.rows.device)
            dist.all_reduce(t, group=group)
            self.n_global = int(t.item())

    # ---- NVLink peer-memory exchange -----------------------------------------------------------
    def enable_peer_exchange(self, nq_max: int, k_max: int) -> bool:
        """Collective.  Allocates this rank's inbox, exchanges CUDA IPC handles through the process group and maps the
        peers' inboxes.  Returns False — on EVERY rank, leaving the all-gather path in place — for a single rank or
        when any rank cannot allocate / export / map (no peer access, IPC unavailable); never raises half-way through a
        collective."""
        import ctypes

        import torch.distributed as dist
        if self.world == 1:
            return False
        if self.world > 16:
            raise ValueError("peer exchange supports at most 16 ranks")
        lib = _lib.load()
        dev = self.local.rows.device
        ptr, handle, err = ctypes.c_void_p(), ctypes.create_string_buffer(64), None
        with torch.cuda.device(dev):
            try:
                nbytes = lib.rir_exchange_bytes(self.world, int(nq_max), int(k_max))
                if nbytes == 0:
                    raise ValueError("bad exchange shape")
                _lib.check(lib.rir_peer_alloc(nbytes, ctypes.byref(ptr)))
                _lib.check(lib.rir_peer_export(ptr, handle))
            except Exception as e:  # reported to every rank below
                err = str(e)
            infos = [None] * self.world
            dist.all_gather_object(infos, (err, bytes(handle.raw)), group=self.group)
            opened, peers = [], (ctypes.c_void_p * self.world)()
            if all(i[0] is None for i in infos):
                try:
                    for g, (_, h) in enumerate(infos):
                        if g == self.rank:
                            peers[g] = ptr.value
                        else:
                            q = ctypes.c_void_p()
                            _lib.check(lib.rir_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(q)))
                            peers[g] = q.value
                            opened.append(q)
                except Exception as e:
                    err = str(e)
            else:
                err = err or "a peer could not allocate its inbox"
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=self.group)  # also: every inbox is mapped and zeroed
            if not all(oks):
                for q in opened:
                    lib.rir_peer_close(q)
                dist.barrier(group=self.group)
                if ptr.value:
                    lib.rir_peer_free(ptr)
                self.peer_exchange_error = err or "a peer could not map the inboxes"
                return False
        self._inbox, self._peers, self._opened = ptr, peers, opened
        self._nq_max, self._k_max, self._epoch = int(nq_max), int(k_max), 0
        return True

    def close(self):
        """Collective.  Unmaps the peers' inboxes and frees this rank's."""
        import torch.distributed as dist
        if self._inbox is None:
            return
        lib = _lib.load()
        torch.cuda.synchronize(self.local.rows.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.local.rows.device):
            for q in self._opened:
                lib.rir_peer_close(q)
            dist.barrier(group=self.group)
            lib.rir_peer_free(self._inbox)
        self._inbox, self._peers, self._opened = None, None, []

    def _exchange_waiters(self, epoch: int, dev):
        lib, inbox = _lib.load(), self._inbox

        def join():
            with torch.cuda.device(dev):
                _lib.check(lib.rir_exchange_join(inbox, epoch, _lib.stream_ptr()))

        def sync():
            with torch.cuda.device(dev):
                _lib.check(lib.rir_exchange_sync(inbox, epoch))
        return join, sync

    def search_async(self, q_rows, q_scale, k: int, path: str = "auto", out=None) -> PendingQuery:
        """Sharded search with the exchange OFF the critical path (RIR_EXCHANGE_ASYNC): this rank's scan + select run
        on the current stream; the merge that waits for the peers' lists runs on a side stream, next to the scan of
        the FOLLOWING search.  Returns a PendingQuery: call `.wait()` (stream order) or `.result()` (host) before
        touching the outputs.  Collective; at most two searches outstanding, with distinct `out` buffers."""
        k = min(int(k), self.n_global)
        q_rows = q_rows.contiguous()
        if q_scale is not None:
            q_scale = q_scale.contiguous()
        if self.world == 1 or self._inbox is None or not (0 < q_rows.shape[0] <= self._nq_max and k <= self._k_max):
            sc, ix = self.search(q_rows, q_scale, k, path=path, out=out)
            ev = torch.cuda.Event()
            ev.record()
            return PendingQuery(ev, sc, ix)
        sc, ix = self._search_peer(q_rows, q_scale, k, path, out=out, flags=RIR_EXCHANGE_ASYNC)
        join, sync = self._exchange_waiters(self._epoch, self.local.rows.device)
        return PendingQuery(None, sc, ix, keep=(q_rows, q_scale), join=join, sync=sync)

    def _search_peer(self, q_rows, q_scale, k: int, path: str, out=None, flags: int = 0):
        lib = _lib.load()
        loc = self.local
        nq = q_rows.shape[0]
        k_local = min(k, loc.n)
        check_k_supported(k_local, loc.n)
        if out is None:
            sc = torch.empty((nq, k), dtype=torch.float32, device=loc.rows.device)
            ix = torch.empty((nq, k), dtype=torch.int32, device=loc.rows.device)
        else:
            sc, ix = out
        ws = loc.workspace(nq, k_local)
        self._epoch += 1
        with torch.cuda.device(loc.rows.device):
            try:
                _lib.check(lib.rir_sim_topk_sharded(
                    q_rows.data_ptr(), loc.rows.data_ptr(), _DTYPES[loc.dtype],
                    None if q_scale is None else q_scale.data_ptr(), None if loc.scale is None else loc.scale.data_ptr(),
                    nq, loc.n, loc.d, k, loc.idx_offset, sc.data_ptr(), ix.data_ptr(), ws.data_ptr(), ws.numel(),
                    PATHS[path] | RIR_WS_CLEAN | flags, _lib.stream_ptr(), self.world, self.rank, self._epoch,
                    self._nq_max, self._k_max, self._peers))
            except _lib.RirError:
                loc._drop_workspaces()
                raise
        return sc, ix

    def query_host_async(self, q_host: torch.Tensor, k: int, out=None, path: str = "auto"):
        """Database.query_host_async over the sharded database (peer exchange must be enabled for world > 1).  Calls
        are collective and must be issued in the same order on every rank; at most two may be in flight (the inbox
        has two parities)."""
        if self.world == 1:
            return self.local.query_host_async(q_host, k, out=out, path=path)
        if self._inbox is None or q_host.shape[0] > self._nq_max or k > self._k_max:
            raise ValueError("enable_peer_exchange(nq_max, k_max) first (and keep nq, k within it)")
        self._epoch += 1
        return self.local.query_host_async(q_host, k, out=out, path=path,
                                           _exchange=(self.world, self.rank, self._epoch, self._nq_max, self._k_max,
                                                      self._peers))

    def query_host(self, q_host: torch.Tensor, k: int, out=None, path: str = "auto"):
        return self.query_host_async(q_host, k, out=out, path=path).result()

    def search(self, q_rows, q_scale, k: int, path: str = "auto", exchange: str = "auto", out=None):
        """exchange: "auto" = peer memory when enabled and the batch fits the inbox, else all-gather; "nccl" forces the
        all-gather + merge path.  Collective: every rank must take the same branch, so the choice depends only on
        arguments that are identical on all ranks (nq, k, the inbox shape) — never on rank-local tensor state."""
        k = min(int(k), self.n_global)
        q_rows = q_rows.contiguous()
        if q_scale is not None:
            q_scale = q_scale.contiguous()
        if exchange != "nccl" and self._inbox is not None and 0 < q_rows.shape[0] <= self._nq_max and k <= self._k_max:
            return self._search_peer(q_rows, q_scale, k, path, out=out)
        k_local = min(k, self.local.n)
        sc, ix = self.local.search(q_rows, q_scale, k_local, path=path,
                                   out=out if (self.world == 1 and k_local == k) else None)
        sc, ix = pad_topk(sc, ix, k)
        if self.world == 1:
            return sc, ix
        all_s, all_i = gather_topk(sc, ix, self.world, self.group)
        return merge_topk(all_s, all_i)


class HostQueryPipeline:
    """Serving loop for host-resident query batches: `submit(q_host)` enqueues one batch and returns a PendingQuery.

    What the reference call site does per batch (iris_evaluate.py:378-386: CPU fp32 query features -> similarity ->
    ranking on the host) as a two-stage device pipeline:
      copy stream : pinned fp32 queries -> H2D -> pack to the shard's dtype  (double-buffered staging)
      main stream : wait for that copy -> search (local or sharded with peer exchange) -> the top-k is written by the
                    select / merge kernel straight into pinned host buffers -> completion event
    With two batches in flight the copy + pack of batch i+1 runs under the scan of batch i and the device never waits
    for the host's turnaround.  `db` is a Database or a ShardedDatabase (peer exchange enabled when world > 1)."""

    def __init__(self, db, nq: int, k: int, depth: int = 2, path: str = "auto"):
        self.db = db
        self.local = db.local if isinstance(db, ShardedDatabase) else db
        loc = self.local
        if loc.rescore_rows is not None or loc.dtype == "fp32":
            raise ValueError("HostQueryPipeline serves bf16 / fp8 shards without a rescoring copy")
        if isinstance(db, ShardedDatabase) and db.world > 1 and (db._inbox is None or nq > db._nq_max or k > db._k_max):
            raise ValueError("enable_peer_exchange(nq_max, k_max) first (and keep nq, k within it)")
        if not (isinstance(db, ShardedDatabase) and db.world > 1) and k > loc.n:
            raise ValueError(f"k={k} exceeds the database size {loc.n}")
        self.nq, self.k, self.depth, self.path = int(nq), int(k), int(depth), path
        dev = loc.rows.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(self.depth):
            self.slots.append({
                "q32": torch.empty((nq, loc.d), dtype=torch.float32, device=dev),
                "qr": torch.empty((nq, loc.d), dtype=loc.rows.dtype, device=dev),
                "qs": torch.empty(nq, dtype=torch.float32, device=dev) if loc.dtype == "fp8" else None,
                "sc": torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                "ix": torch.empty((nq, k), dtype=torch.int32).pin_memory(),
                "copied": torch.cuda.Event(), "done": None})
        self._i = 0

    def submit(self, q_host: torch.Tensor) -> PendingQuery:
        loc, lib = self.local, _lib.load()
        if q_host.is_cuda or q_host.dtype != torch.float32 or tuple(q_host.shape) != (self.nq, loc.d):
            raise TypeError(f"expected a float32 CPU tensor [{self.nq}, {loc.d}] (pinned for an asynchronous copy)")
        slot = self.slots[self._i % self.depth]
        self._i += 1
        main = torch.cuda.current_stream(loc.rows.device)
        dt = _DTYPES[loc.dtype]
        with torch.cuda.device(loc.rows.device):
            if slot["done"] is not None:
                self.copy_stream.wait_event(slot["done"])    # the search that last read this staging slot has finished
            with torch.cuda.stream(self.copy_stream):
                slot["q32"].copy_(q_host, non_blocking=True)
                _lib.check(lib.rir_pack_descriptors(slot["q32"].data_ptr(), self.nq, loc.d, dt, slot["qr"].data_ptr(),
                                                    None if slot["qs"] is None else slot["qs"].data_ptr(),
                                                    self.copy_stream.cuda_stream))
                slot["copied"].record(self.copy_stream)
            main.wait_event(slot["copied"])
            k_local = min(self.k, loc.n)
            check_k_supported(k_local, loc.n)
            ws = loc.workspace(self.nq, k_local)
            qs_ptr = None if slot["qs"] is None else slot["qs"].data_ptr()
            xs_ptr = None if loc.scale is None else loc.scale.data_ptr()
            join = sync = None
            try:
                # pinned host memory is device-addressable (UVA): the kernels store the top-k straight into it
                if isinstance(self.db, ShardedDatabase) and self.db.world > 1:
                    sdb = self.db
                    sdb._epoch += 1
                    _lib.check(lib.rir_sim_topk_sharded(
                        slot["qr"].data_ptr(), loc.rows.data_ptr(), dt, qs_ptr, xs_ptr, self.nq, loc.n, loc.d, self.k,
                        loc.idx_offset, slot["sc"].data_ptr(), slot["ix"].data_ptr(), ws.data_ptr(), ws.numel(),
                        PATHS[self.path] | RIR_WS_CLEAN | RIR_EXCHANGE_ASYNC, main.cuda_stream, sdb.world, sdb.rank,
                        sdb._epoch, sdb._nq_max, sdb._k_max, sdb._peers))
                    join, sync = sdb._exchange_waiters(sdb._epoch, loc.rows.device)   # the merge is on the side stream
                else:
                    _lib.check(lib.rir_sim_topk(
                        slot["qr"].data_ptr(), loc.rows.data_ptr(), dt, qs_ptr, xs_ptr, self.nq, loc.n, loc.d, k_local,
                        loc.idx_offset, slot["sc"].data_ptr(), slot["ix"].data_ptr(), ws.data_ptr(), ws.numel(),
                        PATHS[self.path] | RIR_WS_CLEAN, main.cuda_stream))
            except _lib.RirError:
                loc._drop_workspaces()
                raise
            slot["done"] = torch.cuda.Event()   # scan + select have read the staging slot (and, unsharded, written the top-k)
            slot["done"].record(main)
        return PendingQuery(slot["done"], slot["sc"], slot["ix"], keep=q_host, join=join, sync=sync)


def pad_topk(sc: torch.Tensor, ix: torch.Tensor, k: int):
    """A shard shorter than k pads its list with (-inf, -1) so every rank contributes [nq, k]."""
    have = sc.shape[1]
    if have >= k:
        return sc, ix
    pad_s = torch.full((sc.shape[0], k - have), -math.inf, dtype=sc.dtype, device=sc.device)
    pad_i = torch.full((ix.shape[0], k - have), -1, dtype=ix.dtype, device=ix.device)
    return torch.cat([sc, pad_s], 1), torch.cat([ix, pad_i], 1)


def gather_topk(sc: torch.Tensor, ix: torch.Tensor, world: int, group=None):
    """The one exchange step of the sharded search: all-gather of the packed [nq, k] candidates
    (nq*k*8 bytes per rank; NCCL over NVLink on GPUs, gloo in the CPU tests) -> ([G, nq, k], [G, nq, k])."""
    import torch.distributed as dist
    nq, k = sc.shape
    all_s = torch.empty((world * nq, k), dtype=sc.dtype, device=sc.device)  # concatenated along dim 0 == [G, nq, k]
    all_i = torch.empty((world * nq, k), dtype=ix.dtype, device=ix.device)
    dist.all_gather_into_tensor(all_s, sc.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, ix.contiguous(), group=group)
    return all_s.view(world, nq, k), all_i.view(world, nq, k)


# ----------------------------------------------------------------------------------------------
# alpha query expansion
# ----------------------------------------------------------------------------------------------
def alpha_query_expansion(db, q_rows, q_scale, scores, idx, kq: int = 10, alpha: float = 3.0, group=None):
    """q' = L2(q + sum_{j<kq} max(s_j, 0)^alpha * x_{idx_j}); returns (q'_rows, q'_scale, q'_fp32).

    `db` is a Database or ShardedDatabase; scores / idx are the (merged, global-index) top-k of the first pass."""
    import torch.distributed as dist
    local = db.local if isinstance(db, ShardedDatabase) else db
    lib = _lib.load()
    nq, d = q_rows.shape
    kq = min(kq, scores.shape[1])
    acc = torch.zeros((nq, d), dtype=torch.float32, device=q_rows.device)
    scores = scores.contiguous()
    idx = idx.contiguous()
    dt = _DTYPES[local.dtype]
    with torch.cuda.device(q_rows.device):
        _lib.check(lib.rir_aqe_accumulate(local.rows.data_ptr(), dt, None if local.scale is None else local.scale.data_ptr(),
                                          local.n, local.idx_offset, d, scores.data_ptr(), idx.data_ptr(), nq,
                                          scores.shape[1], kq, float(alpha), acc.data_ptr(), _lib.stream_ptr()))
    if isinstance(db, ShardedDatabase) and db.world > 1:
        dist.all_reduce(acc, group=db.group)
    out32 = torch.empty((nq, d), dtype=torch.float32, device=q_rows.device)
    if local.dtype == "fp32":
        out_q, out_scale = None, None
    else:
        out_q = torch.empty_like(q_rows)
        out_scale = torch.empty(nq, dtype=torch.float32, device=q_rows.device)
    with torch.cuda.device(q_rows.device):
        _lib.check(lib.rir_aqe_finalize(q_rows.data_ptr(), dt, None if q_scale is None else q_scale.data_ptr(),
                                        acc.data_ptr(), nq, d, out32.data_ptr(),
                                        None if out_q is None else out_q.data_ptr(),
                                        None if out_scale is None else out_scale.data_ptr(), _lib.stream_ptr()))
    if local.dtype == "fp32":
        return out32, None, out32
    if local.dtype == "bf16":
        out_scale = None
    return out_q, out_scale, out32


def search_with_aqe(db, q_rows, q_scale, k: int = 100, kq: int = 10, alpha: float = 3.0, path: str = "auto"):
    """First-pass top-k -> alpha-QE -> second-pass top-k (BASELINE cfg-3).  Returns (scores, idx, q'_fp32)."""
    sc, ix = db.search(q_rows, q_scale, k, path=path)
    q2, q2s, q2f = alpha_query_expansion(db, q_rows, q_scale, sc, ix, kq=kq, alpha=alpha)
    sc2, ix2 = db.search(q2, q2s, k, path=path)
    return sc2, ix2, q2f


# ----------------------------------------------------------------------------------------------
# re-ranking hook
# ----------------------------------------------------------------------------------------------
def rerank_topk(scores: torch.Tensor, idx: torch.Tensor, pair_score_fn, top_r: Optional[int] = None):
    """Hook for a pairwise re-ranker over the top of each list (CVNet-Rerank / SuperGlobal style: the reference only has
    the stub `CVNet_Rerank.forward(query_img, key_img) -> score`, models/cvnet_modules/CVNet_Rerank_model.py:49-74, and
    lists both methods on its roadmap, memo.md:46-55).

    scores / idx: [nq, k] from a search (GPU).  `pair_score_fn(q_ids [nq], cand_idx [nq, r]) -> new scores [nq, r]`
    (fp32 CUDA tensor; -1 candidates may get any score) is the caller's model.  The first r = top_r (default k)
    candidates of every query are re-sorted by the new score (descending, ties -> ascending index, by
    rir_merge_topk); the tail keeps its order behind them.  Returns (scores, idx) [nq, k]: the head carries the new
    scores, the tail the old ones."""
    nq, k = idx.shape
    r = k if top_r is None else max(0, min(int(top_r), k))
    if r == 0 or nq == 0:
        return scores, idx
    head_i = idx[:, :r].contiguous()
    q_ids = torch.arange(nq, device=idx.device)
    new = pair_score_fn(q_ids, head_i)
    if not (isinstance(new, torch.Tensor) and new.is_cuda and tuple(new.shape) == (nq, r)):
        raise ValueError(f"pair_score_fn must return a CUDA tensor of shape {(nq, r)}")
    hs, hi = merge_topk(new.float().contiguous()[None], head_i[None])
    if r == k:
        return hs, hi
    return torch.cat([hs, scores[:, r:]], 1), torch.cat([hi, idx[:, r:]], 1)

"""Data formats either side of the hot path.

  RoxfordAndRparis(dataset, dir_main)   same name / arguments / returned cfg keys as dataset/configdataset.py:27-57
                                        (reads `gnd_{dataset}.pkl`; the pickle holds `imlist`, `qimlist`, `gnd`)
  gnd_to_csr(gnd)                       the three sorted CSR id lists rir_compute_map consumes
  DescriptorStore                       a sharded, memory-mappable descriptor store for the R1M distractors: the step
                                        the reference runs per `Distractor_lmdb(partition=...)` range
                                        (dataset/configdataset.py:307-357) writes one shard file; a search rank maps
                                        the shards that cover its row range straight into its HBM layout.
Pure host code (file formats only — no arithmetic), unit-tested on the CPU.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import List, Tuple

import numpy as np
import torch

DATASETS = ['oxford5k', 'paris6k', 'roxford5k', 'rparis6k']


def RoxfordAndRparis(dataset, dir_main):
    """cfg dict for a test set: gnd, imlist, qimlist, n, nq and the file-name lists (dataset/configdataset.py:27-57)."""
    dataset = dataset.lower()
    if dataset not in DATASETS:
        raise ValueError('Unknown dataset: {}!'.format(dataset))
    gnd_fname = os.path.join(dir_main, dataset, 'gnd_{}.pkl'.format(dataset))
    with open(gnd_fname, 'rb') as f:
        cfg = pickle.load(f)
    cfg['gnd_fname'] = gnd_fname
    cfg['ext'] = cfg['qext'] = '.jpg'
    cfg['dir_data'] = os.path.join(dir_main, dataset)
    cfg['dir_images'] = os.path.join(cfg['dir_data'], 'jpg')
    cfg['n'], cfg['nq'] = len(cfg['imlist']), len(cfg['qimlist'])
    cfg['im_fname'] = [os.path.join(cfg['dir_images'], name + '.jpg') for name in cfg['imlist']]
    cfg['qim_fname'] = [os.path.join(cfg['dir_images'], name + '.jpg') for name in cfg['qimlist']]
    cfg['dataset'] = dataset
    return cfg


def gnd_to_csr(gnd, keys=('easy', 'hard', 'junk')):
    """list of per-query dicts -> {key: (ids int32 sorted per query, off int32 [nq+1])}; a missing key is an empty list
    (the reference tolerates a missing 'junk', utils/evaluate.py:70-73)."""
    from .evaluate import ids_to_csr
    return {k: ids_to_csr([g.get(k, []) for g in gnd]) for k in keys}


class DescriptorStore:
    """Directory of raw row-major shard files + `manifest.json`.

    manifest: {"dim": d, "dtype": "bf16"|"fp8"|"fp32", "shards": [{"file", "rows", "row0"}...]}; fp8 shards have a
    sibling `<file>.scale` (fp32 per row).  Files are plain little-endian arrays, so `np.memmap` maps them."""

    _NP = {"bf16": np.uint16, "fp8": np.uint8, "fp32": np.float32}

    def __init__(self, path: str):
        self.path = path
        with open(os.path.join(path, "manifest.json")) as f:
            self.manifest = json.load(f)
        self.dim, self.dtype = int(self.manifest["dim"]), self.manifest["dtype"]
        self.shards = self.manifest["shards"]
        self.n = sum(s["rows"] for s in self.shards)

    # ---- writing ----
    @staticmethod
    def create(path: str, dim: int, dtype: str = "bf16"):
        if dtype not in DescriptorStore._NP:
            raise ValueError(f"dtype must be one of {sorted(DescriptorStore._NP)}")
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "manifest.json"), "w") as f:
            json.dump({"dim": int(dim), "dtype": dtype, "shards": []}, f)
        return DescriptorStore(path)

    def append(self, rows: torch.Tensor, scale: torch.Tensor | None = None):
        """rows [m, dim] already packed (`pack_descriptors`): bf16 / uint8(fp8) / fp32 tensor, any device."""
        r = rows.detach().cpu().contiguous()
        if r.dim() != 2 or r.shape[1] != self.dim:
            raise ValueError(f"expected [m, {self.dim}] rows, got {tuple(r.shape)}")
        want = {"bf16": torch.bfloat16, "fp8": torch.uint8, "fp32": torch.float32}[self.dtype]
        if r.dtype != want:
            raise TypeError(f"store holds {self.dtype} rows, got {r.dtype}")
        if (self.dtype == "fp8") != (scale is not None):
            raise ValueError("fp8 shards need their per-row scales (and only they do)")
        name = f"shard_{len(self.shards):05d}.bin"
        raw = r.view(torch.uint16) if self.dtype == "bf16" else r
        raw.numpy().tofile(os.path.join(self.path, name))
        if scale is not None:
            scale.detach().float().cpu().contiguous().numpy().tofile(os.path.join(self.path, name + ".scale"))
        self.shards.append({"file": name, "rows": int(r.shape[0]), "row0": int(self.n)})
        self.n += int(r.shape[0])
        self.manifest["shards"] = self.shards
        tmp = os.path.join(self.path, "manifest.json.tmp")
        with open(tmp, "w") as f:
            json.dump(self.manifest, f)
        os.replace(tmp, os.path.join(self.path, "manifest.json"))

    # ---- reading ----
    def pieces(self, lo: int, hi: int) -> List[Tuple[dict, int, int]]:
        """(shard, first row inside it, rows) triples covering global rows [lo, hi)."""
        out = []
        for s in self.shards:
            a, b = max(lo, s["row0"]), min(hi, s["row0"] + s["rows"])
            if a < b:
                out.append((s, a - s["row0"], b - a))
        return out

    def load_rows(self, lo: int, hi: int, device="cpu"):
        """Rows [lo, hi) -> (rows tensor in the store dtype, per-row scale or None), memory-mapped then copied once."""
        if not (0 <= lo <= hi <= self.n):
            raise ValueError(f"row range [{lo}, {hi}) outside the store (n = {self.n})")
        tdt = {"bf16": torch.bfloat16, "fp8": torch.uint8, "fp32": torch.float32}[self.dtype]
        rows = torch.empty((hi - lo, self.dim), dtype=tdt, device=device)
        scale = torch.empty(hi - lo, dtype=torch.float32, device=device) if self.dtype == "fp8" else None
        pos = 0
        for s, first, cnt in self.pieces(lo, hi):
            mm = np.memmap(os.path.join(self.path, s["file"]), dtype=self._NP[self.dtype], mode="r",
                           shape=(s["rows"], self.dim))
            blk = torch.from_numpy(np.array(mm[first:first + cnt]))  # one host copy out of the page cache
            if self.dtype == "bf16":
                blk = blk.view(torch.bfloat16)
            rows[pos:pos + cnt] = blk.to(device)
            if scale is not None:
                sm = np.memmap(os.path.join(self.path, s["file"] + ".scale"), dtype=np.float32, mode="r", shape=(s["rows"],))
                scale[pos:pos + cnt] = torch.from_numpy(np.array(sm[first:first + cnt])).to(device)
            pos += cnt
        return rows, scale

    def load_database(self, world: int = 1, rank: int = 0, device=None):
        """The search shard of `rank` (row range of search.shard_bounds) as a resident Database."""
        from .search import Database, shard_bounds
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        lo, hi = shard_bounds(self.n, world, rank)
        rows, scale = self.load_rows(lo, hi, device)
        return Database(rows, scale, self.dtype, idx_offset=lo)

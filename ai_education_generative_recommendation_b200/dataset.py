"""Item-embedding input contract of the encode driver (reference RQ-VAE/vision_data.py:9-30):
float32 [N, dim] `item_embs` (+ JSON `meta`).  Reads the reference's HDF5 file with h5py when it is
importable, otherwise with the package's own reader for that layout (`h5lite`; h5py is not part of this image),
or a `.npy` with the same array."""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.utils.data as data


class EmbDataset(data.Dataset):
    """`EmbDataset(path)` is the reference's dataset (vision_data.py:9-30).  `EmbDataset(path, rank=r, world=G)` opens the
    same file for one worker of a sharded job: only the contiguous item range [r·N/G, (r+1)·N/G) is read (for the chunked +
    gzip layout the reference writes, only the chunks that hold it are inflated), `first_row` / `n_total` say where the
    shard sits, and `pinned()` / `to_device()` give the rows as page-locked memory / start the asynchronous upload."""

    def __init__(self, path, rank: int = 0, world: int = 1):
        self.h5_path = path
        self.rank, self.world = int(rank), int(world)
        self.embeddings, self.meta, self.n_total, self.first_row = self._load_data(path, self.rank, self.world)
        self.dim = self.embeddings.shape[-1]
        what = f"rows {self.first_row}..{self.first_row + len(self.embeddings)} of {self.n_total}" if self.world > 1 else f"{len(self.embeddings)}"
        print(f"[RQ-VAE] Loaded {what} embeddings from {path}, dim={self.dim}")

    @staticmethod
    def _load_data(path, rank=0, world=1):
        if path.endswith(".npy"):
            emb = np.load(path, mmap_mode="r")                 # a shard touches only its pages of the file
            n = emb.shape[0]
            lo, hi = (rank * n) // world, ((rank + 1) * n) // world
            meta_path = os.path.splitext(path)[0] + "_meta.json"
            meta = json.load(open(meta_path)) if os.path.exists(meta_path) else {}
            return np.ascontiguousarray(emb[lo:hi], dtype=np.float32), meta, n, lo
        try:
            import h5py
            lite = False
        except ImportError:
            from . import h5lite as h5py          # same two accesses (vision_data.py:18-21), read-only subset of HDF5
            lite = True
        with h5py.File(path, "r") as f:
            ds = f["item_embs"]
            n = int(ds.shape[0])
            lo, hi = (rank * n) // world, ((rank + 1) * n) // world
            if world == 1:
                emb = ds[:]
            else:
                emb = ds.read_rows(lo, hi) if lite else ds[lo:hi]
            meta = json.loads(f["meta"][()].decode("utf-8")) if "meta" in f else {}
        return np.ascontiguousarray(emb, dtype=np.float32), meta, n, lo

    def pinned(self) -> torch.Tensor:
        """This dataset's rows in page-locked host memory (allocated once): the source of asynchronous H2D copies and of
        `rqb200_generate_codes_host`."""
        if getattr(self, "_pinned", None) is None:
            t = torch.empty(self.embeddings.shape, dtype=torch.float32).pin_memory()
            t.copy_(torch.from_numpy(self.embeddings))
            self._pinned = t
        return self._pinned

    def to_device(self, device, stream=None) -> torch.Tensor:
        """Asynchronous upload of the rows from pinned memory on `stream` (default: the current stream); the returned CUDA
        tensor is ordered after the copy on that stream."""
        src = self.pinned()
        if stream is None:
            return src.to(device, non_blocking=True)
        with torch.cuda.stream(stream):
            return src.to(device, non_blocking=True)

    def shard(self, rank: int, world: int) -> np.ndarray:
        """This rank's contiguous row range [r*N/G, (r+1)*N/G) of the catalogue — what a torchrun worker
        feeds to `sharding.generate_codes_sharded` / `rqb200_generate_codes_host`."""
        if self.world > 1:
            raise ValueError("this dataset already is one shard (opened with rank / world)")
        n = len(self.embeddings)
        lo, hi = (rank * n) // world, ((rank + 1) * n) // world
        return self.embeddings[lo:hi]

    def __getitem__(self, index):
        return torch.from_numpy(np.asarray(self.embeddings[index], dtype=np.float32))

    def __len__(self):
        return len(self.embeddings)

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
    def __init__(self, path):
        self.h5_path = path
        self.embeddings, self.meta = self._load_data(path)
        self.dim = self.embeddings.shape[-1]
        print(f"[RQ-VAE] Loaded {len(self.embeddings)} embeddings from {path}, dim={self.dim}")

    @staticmethod
    def _load_data(path):
        if path.endswith(".npy"):
            emb = np.load(path, mmap_mode="r")
            meta_path = os.path.splitext(path)[0] + "_meta.json"
            meta = json.load(open(meta_path)) if os.path.exists(meta_path) else {}
            return np.ascontiguousarray(emb, dtype=np.float32), meta
        try:
            import h5py
        except ImportError:
            from . import h5lite as h5py          # same two accesses (vision_data.py:18-21), read-only subset of HDF5
        with h5py.File(path, "r") as f:
            emb = f["item_embs"][:]
            meta = json.loads(f["meta"][()].decode("utf-8")) if "meta" in f else {}
        return np.ascontiguousarray(emb, dtype=np.float32), meta

    def shard(self, rank: int, world: int) -> np.ndarray:
        """This rank's contiguous row range [r*N/G, (r+1)*N/G) of the catalogue — what a torchrun worker
        feeds to `sharding.generate_codes_sharded` / `rqb200_generate_codes_host`."""
        n = len(self.embeddings)
        lo, hi = (rank * n) // world, ((rank + 1) * n) // world
        return self.embeddings[lo:hi]

    def __getitem__(self, index):
        return torch.from_numpy(np.asarray(self.embeddings[index], dtype=np.float32))

    def __len__(self):
        return len(self.embeddings)

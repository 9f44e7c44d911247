"""Consumers of the semantic-id file: the token offsetting and the TIGER sequence builder of
reference RQVAE-T5/data_read.ipynb (cells 2-3), on the device.

    tokens = offset_codes(ids, codebook_size)              # [N, L+1] int32, token = id + column * K + 1
    splits = build_tiger_splits(user_ids, item_lists, ids, codebook_size)
    splits.train.history(i) / .target(i)                   # int32 token rows, same values as the notebook's dicts

The reference's own consumers disagree about the row offset (`data[item_id - 1]` in the notebook, `codes[1:]` in
visualize_semantic_id_clusters.py:71); this module follows the notebook (the TIGER training input) and keeps the id
file row-aligned with the embedding rows.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr


def offset_codes(ids: torch.Tensor, codebook_size: int) -> torch.Tensor:
    """item_to_offset_code for every item at once: [N, C] int64 CUDA ids → [N, C] int32 tokens."""
    if not ids.is_cuda or ids.dtype != torch.int64:
        raise RuntimeError("offset_codes needs an int64 CUDA tensor (there is no CPU fallback)")
    ids = ids.contiguous()
    out = torch.empty(ids.shape, dtype=torch.int32, device=ids.device)
    check(_cabi.lib().rqb200_offset_tokens(ptr(ids), ids.shape[0], ids.shape[1], int(codebook_size), ptr(out),
                                           stream_ptr(ids.device)))
    return out


def gather_item_tokens(tokens: torch.Tensor, item_ids: torch.Tensor) -> torch.Tensor:
    """tokens[item_id - 1] for a flat list of 1-indexed item ids → [len, C] int32."""
    item_ids = item_ids.to(device=tokens.device, dtype=torch.int64).contiguous()
    out = torch.empty((item_ids.numel(), tokens.shape[1]), dtype=torch.int32, device=tokens.device)
    bad = torch.zeros((1,), dtype=torch.int32, device=tokens.device)
    check(_cabi.lib().rqb200_gather_item_tokens(ptr(tokens.contiguous()), tokens.shape[0], tokens.shape[1], ptr(item_ids),
                                                item_ids.numel(), ptr(out), ptr(bad), stream_ptr(tokens.device)))
    if int(bad.item()):
        raise IndexError(f"item id outside [1, {tokens.shape[0]}] in the interaction lists")
    return out


@dataclass
class TigerSplit:
    """One of the two datasets of data_read.ipynb cell 2.  Sample s: history = rows [h0, h1) and target = rows [t0, t1)
    of `seq_tokens` (the token rows of all interaction lists back to back) — views, nothing is copied per sample."""
    user_id: np.ndarray            # [S] int32
    hist: np.ndarray               # [S, 2] int64 row ranges into seq_tokens
    tgt: np.ndarray                # [S, 2] int64
    seq_tokens: torch.Tensor       # [total items, C] int32 (CUDA)

    def __len__(self):
        return len(self.user_id)

    def history(self, s: int) -> torch.Tensor:
        return self.seq_tokens[self.hist[s, 0]:self.hist[s, 1]]

    def target(self, s: int) -> torch.Tensor:
        return self.seq_tokens[self.tgt[s, 0]:self.tgt[s, 1]]

    def to_dicts(self) -> List[dict]:
        """The notebook's list of {'user_id', 'history', 'target'} (nested Python ints)."""
        host = self.seq_tokens.cpu().numpy()
        return [{"user_id": int(u), "history": host[h0:h1].tolist(), "target": host[t0:t1].tolist()}
                for u, (h0, h1), (t0, t1) in zip(self.user_id, self.hist, self.tgt)]

    def flat_arrays(self):
        """(user_id int32 [S], history, target) with history/target as lists of flattened int32 arrays — what
        save_single_h5 (cell 3) writes as variable-length rows."""
        host = self.seq_tokens.cpu().numpy()
        return (self.user_id.astype(np.int32), [host[a:b].reshape(-1) for a, b in self.hist],
                [host[a:b].reshape(-1) for a, b in self.tgt])


@dataclass
class TigerSplits:
    train: TigerSplit
    test: TigerSplit


def build_tiger_splits(user_ids: Sequence[int], item_lists: Sequence[Sequence[int]], ids: torch.Tensor,
                       codebook_size: int) -> TigerSplits:
    """Leave-one-out + teacher-forcing split of data_read.ipynb cell 2:
       len < 2 skipped; len == 2 → train (history seq[:1], target seq[1:]);
       len >= 3 → test (seq[:-1] → seq[-1:]) and train (seq[:-2] → seq[1:-1])."""
    tokens = offset_codes(ids, codebook_size)
    lens = np.asarray([len(s) for s in item_lists], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)])[:-1]
    flat = np.concatenate([np.asarray(s, dtype=np.int64) for s in item_lists]) if len(item_lists) else np.zeros(0, np.int64)
    seq_tokens = gather_item_tokens(tokens, torch.from_numpy(flat))
    uid = np.asarray(user_ids)
    two, long_ = lens == 2, lens >= 3
    keep = two | long_                                              # train samples keep the user order
    s, n = starts[keep], lens[keep]
    is_two = two[keep]
    tr_hist = np.stack([s, np.where(is_two, s + 1, s + n - 2)], 1)
    tr_tgt = np.stack([s + 1, np.where(is_two, s + 2, s + n - 1)], 1)
    s3, n3 = starts[long_], lens[long_]
    te_hist = np.stack([s3, s3 + n3 - 1], 1)
    te_tgt = np.stack([s3 + n3 - 1, s3 + n3], 1)
    return TigerSplits(train=TigerSplit(uid[keep].astype(np.int32), tr_hist, tr_tgt, seq_tokens),
                       test=TigerSplit(uid[long_].astype(np.int32), te_hist, te_tgt, seq_tokens))

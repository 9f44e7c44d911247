"""Consumers of the semantic ids: token offsetting and the TIGER split of
RQVAE-T5/data_read.ipynb cells 2-3.  CPU: the oracle restatement against a hand-computed case and the token-range
rules of check_data_alignment.py:105-156; GPU: the CUDA path against the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD


def _ragged(rng, n_users, n_items):
    lens = rng.integers(0, 9, size=n_users)
    lens[:4] = [0, 1, 2, 3]
    return np.arange(100, 100 + n_users), [rng.integers(1, n_items + 1, size=int(k)).tolist() for k in lens]


def test_oracle_tokens_and_split_known_answer(oracle):
    data = np.array([[0, 1, 2, 0], [7, 7, 7, 1], [3, 0, 5, 0]], dtype=np.int64)      # 3 items, K = 8
    assert oracle.item_to_offset_code(data, 1, 8) == [1, 10, 19, 25]
    assert oracle.item_to_offset_code(data, 2, 8) == [8, 16, 24, 26]
    train, test = oracle.tiger_splits([5, 6, 7, 8], [[1], [1, 2], [1, 2, 3], [3, 3, 1, 2]], data, 8)
    a, b, c = ([1, 10, 19, 25], [8, 16, 24, 26], [4, 9, 22, 25])
    assert train == [{"user_id": 6, "history": [a], "target": [b]},
                     {"user_id": 7, "history": [a], "target": [b]},
                     {"user_id": 8, "history": [c, c], "target": [c, a]}]
    assert test == [{"user_id": 7, "history": [a, b], "target": [c]},
                    {"user_id": 8, "history": [c, c, a], "target": [b]}]


def test_token_ranges_of_the_shipped_artifact(oracle):
    """check_data_alignment.py:105-156: position p only emits tokens in [p*K + 1, p*K + K]; PAD (0) never appears."""
    ids = np.load(os.path.join(GOLD, "course_semantic_ids.npy")).astype(np.int64)
    K = 8
    toks = np.array([oracle.item_to_offset_code(ids, i + 1, K) for i in range(len(ids))])
    for p in range(3):                                               # the code columns; the suffix column may exceed K
        assert toks[:, p].min() >= p * K + 1 and toks[:, p].max() <= p * K + K
    assert (toks != 0).all()


@pytest.mark.gpu
def test_cuda_consumers_match_oracle(oracle):
    import ai_education_generative_recommendation_b200 as rq
    rng = np.random.default_rng(9)
    n_items, K = 5000, 256
    ids = np.concatenate([rng.integers(0, K, size=(n_items, 3)), rng.integers(0, 4, size=(n_items, 1))], 1).astype(np.int64)
    idt = torch.from_numpy(ids).cuda()
    toks = rq.offset_codes(idt, K)
    assert toks.dtype == torch.int32
    ref = np.array([oracle.item_to_offset_code(ids, i + 1, K) for i in range(n_items)])
    assert np.array_equal(toks.cpu().numpy(), ref)
    users, lists = _ragged(rng, 700, n_items)
    splits = rq.build_tiger_splits(users, lists, idt, K)
    tr, te = oracle.tiger_splits(users, lists, ids, K)
    assert splits.train.to_dicts() == tr and splits.test.to_dicts() == te
    uid, hist, tgt = splits.test.flat_arrays()
    assert uid.dtype == np.int32 and all(h.dtype == np.int32 for h in hist)
    assert np.array_equal(hist[0], np.array(te[0]["history"], dtype=np.int32).flatten())      # cell 3's vlen rows
    assert torch.equal(splits.train.history(2), torch.tensor(tr[2]["history"], dtype=torch.int32).cuda().reshape(-1, 4))
    with pytest.raises(IndexError):
        rq.build_tiger_splits([1], [[1, n_items + 1]], idt, K)
    empty = rq.build_tiger_splits([], [], idt, K)
    assert len(empty.train) == 0 and len(empty.test) == 0

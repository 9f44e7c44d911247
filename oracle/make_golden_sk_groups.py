"""Generates tests/golden/sinkhorn_groups_c2.npz: REAL collision groups of the C2 synthetic catalogue re-encoded by the
UNMODIFIED reference's Sinkhorn branch (VectorQuantizer.forward(use_sk=True), reference RQ-VAE/models/vq.py:63-83,
layers.py:85-108), and checks the oracle against it.

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists):  python oracle/make_golden_sk_groups.py
The GPU box never runs this; it only reads the committed fixture.

Why a second Sinkhorn fixture: the random groups of sinkhorn_cases.npz rarely produce ties, real groups do.  A row that
dominates several codebook columns ends up with several entries equal to 1/K up to the last bits (1.9 % of the rows of
these groups have a top-2 relative gap below 1e-13, 1.4 % an exact tie), so the arg-max is decided by the exact sequence
of fp64 divisions — the case the CUDA kernels keep every division of the reference for.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/RQ-VAE"
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)

from conftest import load_golden, synth_weights                            # noqa: E402
from oracle import oracle as O                                             # noqa: E402
from ai_education_generative_recommendation_b200 import synth              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_ITEMS, N_GROUPS, EPS, ITERS = 60000, 300, 0.003, 50


def main():
    from models.vq import VectorQuantizer
    O.build()
    g, cfg, cbs = load_golden("c2_slice")
    _, (ew, eb), _ = synth_weights(cfg)
    x = synth.synth_items(2024, 0, N_ITEMS, cfg["in_dim"], 1_000_000)
    z = O.mlp(x, ew, eb)
    codes = O.quantize(z, cbs, want_xq=False)[0]
    r = z.copy()                                             # residual entering the last level (rq.py:47)
    for l in range(len(cbs) - 1):
        q = np.ascontiguousarray(cbs[l], dtype=np.float32)[codes[:, l]]
        r = r - (r + (q - r))
    groups = O.collision_groups(codes)
    groups.sort(key=len, reverse=True)
    picked = groups[:40] + groups[40::max(1, (len(groups) - 40) // (N_GROUPS - 40))][:N_GROUPS - 40]   # the largest + a spread
    cb = np.ascontiguousarray(cbs[-1], dtype=np.float32)
    vq = VectorQuantizer(cb.shape[0], cb.shape[1], sk_epsilon=EPS, sk_iters=ITERS)
    vq.embedding.weight.data.copy_(torch.from_numpy(cb))
    rows, offsets, idx = [], [0], []
    tiny = ties = 0
    for gi in picked:
        rr = np.ascontiguousarray(r[gi])
        with torch.no_grad():
            _, _, ind = vq(torch.from_numpy(rr), use_sk=True)
        d = O.quantize(rr, [cb], want_xq=False, dist_level=0, threads=1)[3]
        mine = O.sinkhorn_assign(d, EPS, ITERS)
        assert np.array_equal(mine, ind.numpy()), "oracle Sinkhorn differs from the reference on a real group"
        Q = np.sort(O.sinkhorn(O.center_distance(d).astype(np.float64), EPS, ITERS), axis=1)
        gap = (Q[:, -1] - Q[:, -2]) / Q[:, -1]
        ties += int((gap == 0).sum())
        tiny += int(((gap > 0) & (gap < 1e-13)).sum())
        rows.append(rr)
        offsets.append(offsets[-1] + len(gi))
        idx.append(ind.numpy())
    rows = np.concatenate(rows)
    np.savez_compressed(os.path.join(GOLD, "sinkhorn_groups_c2.npz"), residual=rows, offsets=np.array(offsets, dtype=np.int64),
                        idx=np.concatenate(idx).astype(np.int16), codebook=cb, eps=np.float64(EPS), iters=np.int64(ITERS))
    print(f"{len(picked)} groups, {len(rows)} rows (largest {max(len(p) for p in picked)}): oracle == reference on all; "
          f"{ties} exact top-2 ties, {tiny} rows with 0 < gap < 1e-13")


if __name__ == "__main__":
    main()

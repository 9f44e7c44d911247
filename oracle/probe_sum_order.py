"""Reveals the fp32 summation ORDER of the reference's CPU GEMM (ATen addmm / matmul → MKL sgemm) for a given
(rows M, inner dimension K, outputs N), on the live stack of this container.  TEST INFRASTRUCTURE ONLY.

Why: the reference re-encodes every collision group as its own batch of 2 … a few dozen rows (reference
RQ-VAE/infer.py:120-122 → rqvae.py:67-71 → layers.py:42-43, vq.py:73) and its GEMM picks another kernel — another
summation order — for such small M than for the catalogue pass.  The oracle (rqvae_oracle.c: dot_lane16,
oracle.py: small_batch_plan) and the product (csrc/small_batch.cu) restate what this script found.

Method (mask probe).  All products are 1.0 except leaf i = +2^40 and leaf j = -2^40 (the bias is leaf K).  Whatever is
added to the running value between the two big terms is absorbed; what is added after they cancelled survives, so
    y = (K + 1) - 2 - #{leaves inside the smallest subtree of the summation tree containing both i and j}.
All pairs (one pair per output column, so a call probes N pairs) give the subtree sizes of all lowest common ancestors,
from which the tree is rebuilt top-down.  The recipe read off the tree is then checked bit for bit on random data.

Usage:  python oracle/probe_sum_order.py                 # the table below for the BASELINE layer shapes
        python oracle/probe_sum_order.py tree 768 256 8   # print the tree for K=768, N=256, M=8
Findings on torch 2.11 + MKL (1 and 8 threads agree unless noted):
  * 2 <= M <= 15 and 24·M <= K  →  "lane16": 16 interleaved fma chains (lane l: k ≡ l mod 16, ascending), folded
    ((p0+p1)+p2)+p3 over the four 128-bit quarters, then (s0+s1)+(s2+s3), bias last.
    Checked for K in {32,48,64,96,104,128,192,256,384,512,640,768,1024,2048} with N = 256 and N = 1024, and for the
    BASELINE shapes (768→256, 256→128, 128→32/64, 1024→256, e 32/64 → K codes 8/256/1024).  Known exceptions (not on
    any BASELINE path, NOT restated): N <= 128 with K in {384, 640} switches back to chains earlier.
  * otherwise the catalogue order: one fma chain per K-block, blocks closed by separate adds, bias first
    (oracle.mkl_kblocks) — except 1024 → 256 with 16 <= M < 176, where the result depends on the THREAD COUNT of the
    reference process (8 threads: four 256-blocks folded pairwise up to M = 128, a third order up to 175; 1 thread: the
    catalogue order from M = 120 at the latest).  The reference is not self-consistent there; not restated.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)
sys.setrecursionlimit(100000)

BIG = 2.0 ** 40


def lca_sizes(M, K, N, matmul=False):
    """S[r, i, j] = number of OTHER leaves inside the smallest subtree holding leaves i and j, for output row r."""
    nl = K + 1
    pairs = [(i, j) for i in range(nl) for j in range(i + 1, nl)]
    S = np.zeros((M, nl, nl), dtype=np.int32)
    x = torch.ones(M, K)
    for c0 in range(0, len(pairs), N):
        chunk = pairs[c0:c0 + N]
        W, b = torch.ones(N, K), torch.ones(N)
        ii, jj, cols = np.array([p[0] for p in chunk]), np.array([p[1] for p in chunk]), np.arange(len(chunk))
        Wn, bn = W.numpy(), b.numpy()
        m = ii < K
        Wn[cols[m], ii[m]] = BIG
        bn[cols[~m]] = BIG
        m = jj < K
        Wn[cols[m], jj[m]] = -BIG
        bn[cols[~m]] = -BIG
        y = (torch.matmul(x, W.t()) + b if matmul else torch.nn.functional.linear(x, W, b)).numpy()
        S[:, ii, jj] = nl - 2 - y[:, :len(chunk)].astype(np.int32)
        S[:, jj, ii] = S[:, ii, jj]
    return S


def build(S, leaves):
    """Summation tree (nested tuples of leaf ids) of one output from its LCA-size matrix."""
    if len(leaves) == 1:
        return int(leaves[0])
    n = len(leaves)
    rem, children = np.asarray(leaves), []
    while len(rem):
        m = S[rem[0], rem] + 2 < n           # same child of this node ⇔ their own subtree is smaller than the node
        m[0] = True
        children.append(rem[m].tolist())
        rem = rem[~m]
    if len(children) == 1:
        raise RuntimeError("inconsistent probe (non-associative pattern?)")
    return tuple(build(S, c) for c in children)


def fmt(t, K):
    """Compact rendering: [head >> a..b step s] = a left-deep chain that adds leaves a, a+s, … to `head`."""
    if isinstance(t, int):
        return "b" if t == K else str(t)
    chain, cur = [], t
    while isinstance(cur, tuple) and len(cur) == 2 and isinstance(cur[1], int):
        chain.append(cur[1])
        cur = cur[0]
    if len(chain) >= 2:
        chain = chain[::-1]
        d = set(np.diff(chain).tolist())
        body = f"{chain[0]}..{chain[-1]} step {d.pop()}" if len(d) == 1 else str(chain)
        return f"[{fmt(cur, K)} >> {body}]"
    return "(" + " + ".join(fmt(c, K) for c in t) + ")"


def is_lane16(M, K, N, matmul=False):
    """Cheap classifier: in lane16 leaves 0 and 1 sit in different lanes (their LCA is near the root)."""
    x, W, b = torch.ones(M, K), torch.ones(N, K), torch.zeros(N)
    W[:, 0], W[:, 1] = BIG, -BIG
    y = (torch.matmul(x, W.t()) if matmul else torch.nn.functional.linear(x, W, b)).numpy()
    return (K - 2.0) not in set(y.ravel().tolist())


def check_recipe():
    """The oracle's restatement against the live stack, bit for bit, on random data."""
    from oracle import oracle as O
    O.build()
    torch.manual_seed(1)
    bad = 0
    for (K, N) in [(768, 256), (256, 128), (128, 32), (128, 64), (1024, 256)]:
        for M in list(range(2, 40)) + [100, 300]:
            if K == 1024 and 16 <= M < 176:
                continue
            x, w, b = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
            ref = torch.nn.functional.linear(x, w, b).numpy()
            mine = O.linear_group(x.numpy(), w.numpy(), b.numpy(), relu=False)
            if not np.array_equal(ref.view(np.int32), mine.view(np.int32)):
                bad += 1
                print("MISMATCH linear", K, N, M)
    for (e, Kc) in [(32, 8), (32, 256), (64, 256), (64, 1024)]:
        for M in list(range(2, 20)) + [40]:
            r, E = torch.randn(M, e), torch.randn(Kc, e)
            d_ref = (torch.sum(r ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1, keepdim=True).t()
                     - 2 * torch.matmul(r, E.t())).numpy()
            d = O.quantize(r.numpy(), [E.numpy()], want_xq=False, dist_level=0, threads=1,
                           dot_kind=O.small_batch_plan(M, e, Kc))[3]
            if not np.array_equal(d.view(np.int32), d_ref.view(np.int32)):
                bad += 1
                print("MISMATCH distance", e, Kc, M)
    print("recipe check: mismatching (shape, M) cases:", bad)
    return bad


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("NT", os.cpu_count() or 1)))
    if sys.argv[1:2] == ["tree"]:
        K, N, M = (int(a) for a in sys.argv[2:5])
        S = lca_sizes(M, K, N, matmul="matmul" in sys.argv)
        seen = {}
        for r in range(M):
            seen.setdefault(fmt(build(S[r], list(range(K + 1))), K), []).append(r)
        for f, rows in seen.items():
            print(f"rows {rows}: {f}")
        sys.exit(0)
    print("torch", torch.__version__, "threads", torch.get_num_threads())
    for matmul, K, N in [(False, 768, 256), (False, 256, 128), (False, 128, 32), (False, 128, 64), (False, 1024, 256),
                         (True, 32, 8), (True, 32, 256), (True, 64, 256), (True, 64, 1024)]:
        ms = [M for M in range(2, 40) if is_lane16(M, K, N, matmul)]
        rule = [M for M in range(2, 40) if 2 <= M <= 15 and 24 * M <= K]
        print(f"{'matmul' if matmul else 'linear'} K={K} N={N}: lane16 for M in {ms[:1]}..{ms[-1:]}  rule says {rule[:1]}..{rule[-1:]}"
              f"  {'OK' if ms == rule else 'DIFFERENT'}")
    sys.exit(1 if check_recipe() else 0)

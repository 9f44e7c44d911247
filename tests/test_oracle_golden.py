"""CPU: the oracle against the committed golden vectors (made by oracle/make_golden.py from the reference)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import infer_fixture, load_golden, synth_weights
from ai_education_generative_recommendation_b200 import synth


@pytest.mark.parametrize("name", ["c1_slice", "c2_slice", "c3_slice", "c5_slice", "odd1_slice", "odd2_slice", "odd3_slice"])
def test_oracle_matches_reference_slices(oracle, name):
    g, cfg, cbs = load_golden(name)
    _, (ew, eb), (dw, db) = synth_weights(cfg)
    n = int(g["n_rows"])
    x = synth.synth_items(int(g["seed"]), 0, n, cfg["in_dim"], int(g["n_total"]))
    z = oracle.mlp(x, ew, eb)
    keep = g["z_head"].shape[0]
    assert np.array_equal(z[:keep].view(np.int32), g["z_head"].view(np.int32))          # bit-exact latent
    idx, xq, ssq, _ = oracle.quantize(z, cbs)
    assert np.array_equal(idx, g["codes"].astype(np.int64))                             # bit-exact codes
    assert np.array_equal(xq[:keep].view(np.int32), g["xq_head"].view(np.int32))
    out = oracle.mlp(xq[:8], dw, db) if n >= 16 else None
    loss = oracle.rq_loss(ssq, n, cfg["e_dim"], 0.25)
    assert abs(loss - float(g["rq_loss"])) <= 1e-5 * abs(float(g["rq_loss"]))
    full = oracle.mlp(xq, dw, db)
    assert np.array_equal(full[:8].view(np.int32), g["out_head"].view(np.int32))
    recon = float(np.mean((full.astype(np.float64) - x.astype(np.float64)) ** 2))
    assert abs(recon - float(g["recon_loss"])) <= 1e-5 * float(g["recon_loss"])


def test_oracle_suffix_reproduces_shipped_artifact(oracle):
    import os
    arr = np.load(os.path.join(os.path.dirname(__file__), "golden", "course_semantic_ids.npy")).astype(np.int64)
    assert arr.shape == (707, 4)
    got = oracle.suffix_dedup(arr[:, :3])
    assert np.array_equal(got, arr)
    assert len(np.unique(got, axis=0)) == 707


def test_oracle_sinkhorn_cases(oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sinkhorn_cases.npz"))
    for i in range(int(g["n_cases"])):
        B, K, iters = [int(v) for v in g[f"meta{i}"]]
        d = oracle.quantize(g[f"r{i}"], [g[f"cb{i}"]], want_xq=False, dist_level=0, threads=1)[3]
        got = oracle.sinkhorn_assign(d, float(g[f"eps{i}"]), iters)
        assert np.array_equal(got, g[f"idx{i}"].astype(np.int64)), i


def test_oracle_driver_reproduces_reference_infer_exactly(oracle):
    """BASELINE config 1: the reference's verbatim infer() on 707 items.  With every collision group re-encoded as its
    own small batch (the reference's literal computation, in its small-batch summation order) the oracle driver gives the
    reference's ids on ALL rows and the reference's codes after EVERY round."""
    f = infer_fixture("c1_infer")
    cfg, cbs, x, golden, trace = f["cfg"], f["codebooks"], f["x"], f["semantic_ids"], f["trace"]
    _, (ew, eb), _ = synth_weights(cfg)
    assert golden.shape == (707, 4) and len(np.unique(golden, axis=0)) == 707
    mine = []
    got, stats = oracle.generate_codes(x, ew, eb, cbs, cfg["sk_epsilons"], cfg["sk_iters"], group_order=True, trace=mine,
                                       batch_size=64)          # the reference ran with batch 64: a last batch of 3 rows
    assert np.array_equal(got, golden)
    assert stats["rounds"] == f["rounds"] == len(mine) - 1
    for a, b in zip(mine, trace):
        assert np.array_equal(a, b)
    assert np.array_equal(oracle.suffix_dedup(trace[-1]), golden)
    # the round-1 shortcut (re-quantise from the catalogue-pass latent) is NOT what the reference computes
    old, _ = oracle.generate_codes(x, ew, eb, cbs, cfg["sk_epsilons"], cfg["sk_iters"], group_order=False)
    assert (old != golden).any(1).sum() == 6


@pytest.mark.parametrize("name,rounds", [("c2_infer", (0, 29)), ("c3_infer", (0, 1, 29))])
def test_oracle_rounds_against_reference_infer_at_catalogue_shapes(oracle, name, rounds):
    """BASELINE config 2 / 3 shapes, 60 000 / 20 000 items through the unmodified reference infer().  Every round is a
    pure function of the codes before it (infer.py:117-129): re-encode the reference's groups of round t and compare with
    the reference's codes after round t.  Latents, prefix codes and distances are restated bit for bit; the only rows that
    may differ are those whose two best codes TIE in the reference's own fp64 Sinkhorn matrix (relative gap <= 1e-10,
    recorded in the fixture by oracle/make_golden.py) — there the reference's result hangs on the last bit of its exp()."""
    f = infer_fixture(name)
    cfg, cbs, x, trace, ties = f["cfg"], f["codebooks"], f["x"], f["trace"], f["ties"]
    _, (ew, eb), _ = synth_weights(cfg)
    Lv = len(cbs)
    eps = [0.0] * (Lv - 1) + [cfg["sk_epsilons"][-1]]
    assert np.array_equal(oracle.quantize(oracle.mlp(x, ew, eb), cbs, want_xq=False)[0], trace[0])     # pass 1: exact
    assert np.array_equal(oracle.suffix_dedup(trace[-1]), f["semantic_ids"])
    checked = 0
    for t in rounds:
        tie = np.zeros(f["n"], dtype=bool)
        tie[ties[t]] = True
        for grp in oracle.collision_groups(trace[t]):
            mine = oracle.reencode_group(x[grp], ew, eb, cbs, eps, cfg["sk_iters"])
            ref = trace[t + 1][grp]
            assert np.array_equal(mine[:, :-1], ref[:, :-1])                   # arg-min levels: always exact
            diff = (mine != ref).any(1)
            assert not (diff & ~tie[grp]).any()                                # differing rows ⊂ fp64-tie set, exactly
            checked += len(grp)
    assert checked > 500


def brute_suffix(codes):
    out = []
    for i in range(len(codes)):
        out.append(sum(1 for j in range(i) if (codes[j] == codes[i]).all()))
    return np.array(out, dtype=np.int64)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 60), st.integers(1, 5), st.integers(1, 4), st.integers(0, 2 ** 31 - 1))
def test_oracle_suffix_property(oracle, n, L, K, seed):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, K, size=(n, L)).astype(np.int64)
    got = oracle.suffix_dedup(codes) if n else np.zeros((0, L + 1), dtype=np.int64)
    assert np.array_equal(got[:, :L], codes)
    assert np.array_equal(got[:, L], brute_suffix(codes)) if n else True


def test_sumsq_order_is_not_plain_sum(oracle):
    rng = np.random.default_rng(0)
    v = rng.standard_normal((2000, 64)).astype(np.float32)
    got = oracle.sumsq(v)
    seq = np.zeros(2000, dtype=np.float32)
    for k in range(64):
        seq = seq + v[:, k] * v[:, k]
    assert np.allclose(got, seq, rtol=1e-5)
    assert (got != seq).any()          # the ATen order differs from a left-to-right sum in the last bit


def test_synth_generator_known_answers():
    x = synth.synth_items(2024, 0, 2048, 768, 1_000_000)
    assert x.dtype == np.float32 and x.shape == (2048, 768)
    assert not x[0].any()                                   # padding row
    assert np.array_equal(x[5:9], synth.synth_items(2024, 5, 4, 768, 1_000_000))   # row ranges compose
    assert abs(float(x[1:].std()) - 0.51) < 0.02
    big = synth.synth_items(2024, 0, 20000, 8, 20000)
    uniq = len(np.unique(big, axis=0))
    assert 20000 - 60 < uniq < 20000                        # ~0.1 % exact duplicates


def test_kblock_rule(oracle):
    assert oracle.mkl_kblocks(768) == [384, 384]
    assert oracle.mkl_kblocks(256) == [256]
    assert oracle.mkl_kblocks(1024) == [384, 384, 256]
    assert oracle.mkl_kblocks(400) == [200, 200]


def test_oracle_sinkhorn_real_collision_groups(oracle):
    """300 real collision groups of the C2 catalogue re-encoded by the reference's own Sinkhorn branch
    (oracle/make_golden_sk_groups.py): rows with exact and last-bit ties in Q, where only the reference's sequence of fp64
    divisions gives the reference's arg-max."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sinkhorn_groups_c2.npz"))
    off, cb = g["offsets"], g["codebook"]
    eps, iters = float(g["eps"]), int(g["iters"])
    for a, b in zip(off[:-1], off[1:]):
        r = np.ascontiguousarray(g["residual"][a:b])
        d = oracle.quantize(r, [cb], want_xq=False, dist_level=0, threads=1)[3]
        assert np.array_equal(oracle.sinkhorn_assign(d, eps, iters), g["idx"][a:b].astype(np.int64)), int(a)


def _kmeans_case(g, i):
    n, e, K, iters, n_total = (int(v) for v in g[f"meta{i}"])
    x = synth.synth_items(2024, 1, n, e, n_total)
    init = np.ascontiguousarray(x[(np.arange(K) * (n // K)) % n])
    if n == 3000:
        init[1::32] = init[0::32]          # duplicated initial centres → empty clusters (as in make_golden_kmeans.py)
    return x, init, K, iters, g[f"centers{i}"]


def _sorted_rows(a):
    return a[np.lexsort(a.T[::-1])]


def test_oracle_lloyd_matches_scikit_learn_golden(oracle):
    """scikit-learn `KMeans(init=<array>, n_init=1, algorithm="lloyd")` — what layers.py:77 runs after seeding — is
    deterministic; its centres (oracle/make_golden_kmeans.py) pin the oracle's Lloyd loop incl. the relocation of empty
    clusters (case 3 loses 29 clusters after the first iteration).  Tolerance: fp32 rounding of the centres."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmeans_sklearn.npz"))
    for i in range(int(g["n_cases"])):
        x, init, K, iters, want = _kmeans_case(g, i)
        got = oracle.kmeans_lloyd(x, init, iters)
        rel = np.abs(_sorted_rows(got) - _sorted_rows(want)).max() / np.abs(want).max()
        assert rel <= 1e-6, (i, rel)


def test_host_kmeans_logic_matches_scikit_learn_golden(oracle):
    """The product's Lloyd driver (kmeans_gpu.kmeans_fit: statistics → relocation of empty clusters → update) with the
    per-rank kernels replaced by a numpy twin — the host logic the GPU path shares — against the same golden centres."""
    import os
    import torch
    from test_sharding_gloo import NumpyKMeansOps
    from ai_education_generative_recommendation_b200.kmeans_gpu import kmeans_fit
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmeans_sklearn.npz"))
    for i in (2, 4):
        x, init, K, iters, want = _kmeans_case(g, i)
        got = kmeans_fit(torch.from_numpy(x), K, iters, init=torch.from_numpy(init), ops=NumpyKMeansOps()).numpy()
        rel = np.abs(_sorted_rows(got) - _sorted_rows(want)).max() / np.abs(want).max()
        assert rel <= 1e-5, (i, rel)

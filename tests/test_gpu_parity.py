"""GPU: the CUDA path (through the C ABI) against the golden vectors and the CPU oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import build_model, infer_fixture, load_golden, synth_weights
from ai_education_generative_recommendation_b200 import _cabi, synth
import ai_education_generative_recommendation_b200 as rq

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_synth(seed, first, n, dim, n_total):
    x = torch.empty((n, dim), dtype=torch.float32, device=DEV)
    _cabi.check(_cabi.lib().rqb200_synth_items(seed, first, n, dim, n_total, x.data_ptr(), _cabi.stream_ptr()))
    return x


def test_native_library_is_loaded_and_sees_the_gpu():
    assert _cabi.lib().rqb200_device_count() >= 1
    assert os.path.basename(_cabi.LIB_PATH) == "librqvae_b200.so"


def test_synth_kernel_matches_numpy_twin():
    for (first, n, dim, tot) in [(0, 4096, 768, 1_000_000), (999_000, 1000, 768, 1_000_000), (12345, 777, 1024, 10 ** 8)]:
        got = gpu_synth(2024, first, n, dim, tot).cpu().numpy()
        ref = synth.synth_items(2024, first, n, dim, tot)
        assert np.array_equal(got.view(np.int32), ref.view(np.int32))


@pytest.mark.parametrize("name", ["c1_slice", "c2_slice", "c3_slice", "c5_slice"])
def test_get_indices_and_forward_match_reference_golden(oracle, name):
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    n = int(g["n_rows"])
    x = synth.synth_items(int(g["seed"]), 0, n, cfg["in_dim"], int(g["n_total"]))
    xt = torch.from_numpy(x).to(DEV)
    codes = m.get_indices(xt, use_sk=False)
    assert codes.dtype == torch.int64 and tuple(codes.shape) == (n, len(cbs))
    assert np.array_equal(codes.cpu().numpy(), g["codes"].astype(np.int64))              # bit-exact codes
    z = m.encoder(xt).cpu().numpy()
    keep = g["z_head"].shape[0]
    assert np.array_equal(z[:keep].view(np.int32), g["z_head"].view(np.int32))           # bit-exact latent
    out, rq_loss, idx = m(xt, use_sk=False)
    assert np.array_equal(idx.cpu().numpy(), g["codes"].astype(np.int64))
    assert np.array_equal(out[:8].cpu().numpy().view(np.int32), g["out_head"].view(np.int32))
    assert abs(float(rq_loss) - float(g["rq_loss"])) <= 1e-4 * abs(float(g["rq_loss"]))  # tolerance: 1e-4 rel
    total, recon = m.compute_loss(out, rq_loss, xs=xt)
    assert abs(float(recon) - float(g["recon_loss"])) <= 1e-4 * float(g["recon_loss"])
    assert abs(float(total) - float(g["total_loss"])) <= 1e-4 * float(g["total_loss"])
    out2, rq2, idx2, total2, recon2 = m.forward_losses(xt)
    assert torch.equal(out2, out) and torch.equal(idx2, idx)
    assert abs(float(recon2) - float(g["recon_loss"])) <= 1e-4 * float(g["recon_loss"])
    x_q, _, _ = m.rq(m.encoder(xt), use_sk=False)
    assert np.array_equal(x_q[:keep].cpu().numpy().view(np.int32), g["xq_head"].view(np.int32))
    # whole-slice check against the oracle as well (latent, x_q, decoded rows)
    _, (ew, eb), (dw, db) = synth_weights(cfg)
    zo = oracle.mlp(x, ew, eb)
    assert np.array_equal(z.view(np.int32), zo.view(np.int32))
    _, xqo, _, _ = oracle.quantize(zo, cbs)
    assert np.array_equal(x_q.cpu().numpy().view(np.int32), xqo.view(np.int32))
    assert np.array_equal(out.cpu().numpy().view(np.int32), oracle.mlp(xqo, dw, db).view(np.int32))


def test_registered_torch_ops_equal_the_ctypes_binding():
    """torch.ops.rqvae_b200.* (csrc/torch_ops.cpp) and the ctypes binding call the same C ABI: same tensors out."""
    from ai_education_generative_recommendation_b200 import torch_ops
    torch_ops.load()
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    n = 3000
    xt = torch.from_numpy(synth.synth_items(2024, 0, n, cfg["in_dim"], 1_000_000)).to(DEV)
    h = torch_ops.handle(m)
    want = g["codes"][:n].astype(np.int64)
    for mode in (_cabi.ENCODE_EXACT, _cabi.ENCODE_FAST):
        assert np.array_equal(torch.ops.rqvae_b200.encode_indices(h, xt, mode).cpu().numpy(), want)
    z = torch.ops.rqvae_b200.encode_latents(h, xt)
    assert torch.equal(z, m.encoder(xt))
    idx, xq, sumsq = torch.ops.rqvae_b200.quantize(h, z)
    x_q, _, codes = m.rq(z, use_sk=False)
    assert np.array_equal(idx.cpu().numpy(), want) and torch.equal(idx, codes) and torch.equal(xq, x_q)
    ids = torch.ops.rqvae_b200.resolve_collisions(h, idx, list(cfg["num_emb_list"]))
    ref, stats = rq.suffix_dedup(m, idx)
    assert torch.equal(ids, ref)
    assert tuple(torch.ops.rqvae_b200.collision_rate(h, idx, list(cfg["num_emb_list"]))) == (stats["distinct"], stats["max_conflicts"])
    d = torch.rand(24, 256, device=DEV)
    a = torch.ops.rqvae_b200.sinkhorn_assign(d, 0.003, 50)
    scratch = torch.empty((24, 256), dtype=torch.float64, device=DEV)
    b = torch.empty((24,), dtype=torch.int64, device=DEV)
    _cabi.check(_cabi.lib().rqb200_sinkhorn_assign(d.data_ptr(), 24, 256, 0.003, 50, scratch.data_ptr(), b.data_ptr(), _cabi.stream_ptr()))
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor|CPU"):
        torch.ops.rqvae_b200.encode_indices(h, xt.cpu(), 0)


def test_ragged_and_empty_batches(oracle):
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    _, (ew, eb), _ = synth_weights(cfg)
    for n in (0, 1, 2, 3, 15, 16, 127, 128, 129, 1000):
        x = synth.synth_items(7, 1, n, cfg["in_dim"], 5000) if n else np.zeros((0, cfg["in_dim"]), np.float32)
        got = m.get_indices(torch.from_numpy(x).to(DEV)).cpu().numpy()
        assert got.shape == (n, 3)
        if n:
            assert np.array_equal(got, oracle.get_indices(x, ew, eb, cbs))
    x3 = torch.from_numpy(synth.synth_items(7, 1, 24, cfg["in_dim"], 5000)).to(DEV).view(2, 3, 4, -1)
    assert tuple(m.get_indices(x3).shape) == (2, 3, 4, 3)                                 # leading dims kept


def test_exact_integer_lane_with_ties(oracle):
    """Small-integer inputs/weights: every partial sum is exact in fp32, ties are abundant, so any
    summation order gives the same bits and only first-index tie breaking decides."""
    rng = np.random.default_rng(5)
    cfg = dict(in_dim=768, num_emb_list=[256, 256, 256], e_dim=32, layers=[256, 128], sk_epsilons=[0.0] * 3, sk_iters=5)
    n = 4096
    x = rng.integers(-2, 3, size=(n, 768)).astype(np.float32)
    def sparse(o, i, dens):
        return (rng.integers(-1, 2, size=(o, i)) * (rng.random((o, i)) < dens)).astype(np.float32)
    ew = [sparse(256, 768, 0.10), sparse(128, 256, 0.10), sparse(32, 128, 0.25)]
    eb = [rng.integers(-1, 2, size=o).astype(np.float32) for o in (256, 128, 32)]
    z = oracle.mlp(x, ew, eb)
    cbs = []
    r = z.copy()
    for lvl in range(3):
        cb = r[rng.integers(0, n, size=256)].copy()
        cb[128:] = cb[:128]                                   # duplicated codebook rows force exact ties
        cbs.append(cb)
        idx, _, _, _ = oracle.quantize(r, [cb], want_xq=False)
        r = r - (r + (cb[idx[:, 0]] - r))
    ref = oracle.quantize(z, cbs, want_xq=False)[0]
    assert (ref < 128).all()                                  # first index always wins
    m = rq.RQVAE(in_dim=768, num_emb_list=[256] * 3, e_dim=32, layers=[256, 128], sk_epsilons=[0.0] * 3).to(DEV).eval()
    with torch.no_grad():
        for i, lin in enumerate([m.encoder.mlp_layers[1], m.encoder.mlp_layers[4], m.encoder.mlp_layers[7]]):
            lin.weight.copy_(torch.from_numpy(ew[i])); lin.bias.copy_(torch.from_numpy(eb[i]))
        for lvl in range(3):
            m.rq.vq_layers[lvl].embedding.weight.copy_(torch.from_numpy(cbs[lvl]))
    got = m.get_indices(torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", [5000, 20011])
@pytest.mark.parametrize("name", ["c1_slice", "c2_slice", "c5_slice"])
def test_sliced_tiled_and_row_per_thread_quantizers_agree(oracle, name, n):
    """The exact quantizer has three thread mappings: a row per thread (any outputs); codes only — a row shared by
    8 lanes when few rows are in flight (n = 5000) and a register-tiled 64-row CTA beyond that (n = 20011).  Same codes
    bit for bit, including NaN rows ("NaN is the minimum", torch.argmin) and exact ties."""
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    m._sync()
    e, Lv = cfg["e_dim"], len(cbs)
    rng = np.random.default_rng(21)
    z = (rng.standard_normal((n, e)) * 0.3).astype(np.float32)
    z[7, 3] = np.nan                                              # a NaN latent: every distance is NaN, code 0 wins
    z[n - 1, 0] = np.inf
    z[100:200] = cbs[0][rng.integers(0, cbs[0].shape[0], size=100)]   # rows sitting exactly on codes
    km = min(100, cbs[0].shape[0] - 1)
    z[300:300 + km] = 0.5 * (cbs[0][:km] + cbs[0][1:km + 1])      # midpoints between neighbouring codes
    zt = torch.from_numpy(z).to(DEV)
    lib = _cabi.lib()
    sliced = torch.empty((n, Lv), dtype=torch.int64, device=DEV)
    plain = torch.empty((n, Lv), dtype=torch.int64, device=DEV)
    xq = torch.empty((n, e), dtype=torch.float32, device=DEV)    # asking for x_q selects the row-per-thread kernel
    _cabi.check(lib.rqb200_quantize(m._handle, zt.data_ptr(), n, sliced.data_ptr(), 0, 0, 0, 0, _cabi.stream_ptr()))
    _cabi.check(lib.rqb200_quantize(m._handle, zt.data_ptr(), n, plain.data_ptr(), 0, xq.data_ptr(), 0, 0, _cabi.stream_ptr()))
    assert torch.equal(sliced, plain)
    ok = np.isfinite(z).all(1)
    ref = oracle.quantize(z[ok], cbs, want_xq=False)[0]
    assert np.array_equal(sliced.cpu().numpy()[ok], ref)


def test_distances_match_oracle(oracle):
    g, cfg, cbs = load_golden("c3_slice")
    m = build_model(cfg, cbs)
    m._sync()
    rng = np.random.default_rng(1)
    r = (rng.standard_normal((300, cfg["e_dim"])) * 0.4).astype(np.float32)
    for lvl in (0, 3):
        d = m._distances(lvl, torch.from_numpy(r).to(DEV)).cpu().numpy()
        ref = oracle.quantize(r, [cbs[lvl]], want_xq=False, dist_level=0, threads=1)[3]
        assert np.array_equal(d.view(np.int32), ref.view(np.int32))


@pytest.mark.parametrize("n,L,K", [(0, 3, 8), (1, 3, 8), (5000, 3, 8), (100_000, 3, 256), (70_001, 4, 1024), (3000, 1, 2),
                                   (200_000, 5, 16)])
def test_suffix_dedup_bit_exact(oracle, n, L, K):
    rng = np.random.default_rng(n + L)
    codes = rng.integers(0, K, size=(n, L)).astype(np.int64)
    if n > 10:
        codes[rng.integers(0, n, size=n // 3)] = codes[rng.integers(0, n, size=n // 3)]      # many duplicates
    ct = torch.from_numpy(codes).to(DEV)
    out, stats = rq.suffix_dedup(None, ct, [K] * L)
    ref = oracle.suffix_dedup(codes) if n else np.zeros((0, L + 1), np.int64)
    assert np.array_equal(out.cpu().numpy(), ref)
    if n:
        assert stats["distinct"] == len(np.unique(codes, axis=0))
        assert stats["max_conflicts"] == int(ref[:, -1].max()) + 1
        out2, _ = rq.suffix_dedup(None, ct, None)               # column ranges scanned on the device
        assert np.array_equal(out2.cpu().numpy(), ref)
        out3, none = rq.suffix_dedup(None, ct, [K] * L, want_stats=False)      # enqueue-only form: no statistics, same ids
        assert none is None and np.array_equal(out3.cpu().numpy(), ref)


@pytest.mark.parametrize("case", ["one_run", "two_runs_and_tail", "wide_keys_8_passes", "eight_levels", "tile_edges"])
def test_suffix_dedup_long_runs_and_wide_keys(oracle, case):
    """The one-launch digit passes (look-back over many tiles) and the rank kernel's backward search for the start of a
    run that reaches into a tile: runs spanning hundreds of tiles, keys of 60 / 64 bits (eight digit passes), runs that
    begin and end exactly on tile boundaries."""
    rng = np.random.default_rng(5)
    n = 300_007
    if case == "one_run":
        codes, K = np.full((n, 3), 7, dtype=np.int64), [256] * 3
    elif case == "two_runs_and_tail":
        codes, K = rng.integers(0, 256, size=(n, 3)).astype(np.int64), [256] * 3
        codes[rng.permutation(n)[:200_000]] = (3, 1, 4)
        codes[rng.permutation(n)[:60_000]] = (3, 1, 5)
    elif case == "wide_keys_8_passes":
        codes, K = rng.integers(0, 4096, size=(n, 5)).astype(np.int64), [4096] * 5
        codes[rng.integers(0, n, size=n // 2)] = codes[rng.integers(0, n, size=n // 2)]
    elif case == "eight_levels":
        codes, K = rng.integers(0, 256, size=(n, 8)).astype(np.int64), [256] * 8
        codes[:, :6] = codes[:, :1] % 3                         # few distinct high digits, random low ones
        codes[rng.integers(0, n, size=n // 2)] = codes[rng.integers(0, n, size=n // 2)]
    else:
        n = 4 * 4096 + 2048                                     # sorted positions: runs of exactly 2048 / 4096 keys
        codes, K = (np.arange(n)[:, None] // np.array([[4096, 2048, 1 << 30]])).astype(np.int64), [8, 16, 2]
        codes = codes[rng.permutation(n)]
    ct = torch.from_numpy(codes).to(DEV)
    ref = oracle.suffix_dedup(codes)
    for Ks in (K, None):
        out, stats = rq.suffix_dedup(None, ct, Ks)
        assert np.array_equal(out.cpu().numpy(), ref)
        assert stats["max_conflicts"] == int(ref[:, -1].max()) + 1
        assert stats["distinct"] == len(np.unique(codes, axis=0))


def test_peer_memory_dedup_single_rank_equals_plain_dedup(oracle):
    """rqb200_shard_* with world = 1 (no peers): the owner is the rank itself, same ids as rqb200_suffix_dedup."""
    from ai_education_generative_recommendation_b200 import sharding
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    rng = np.random.default_rng(11)
    for n, spread in ((0, 4), (1, 4), (5000, 4), (70_001, 40)):
        codes = rng.integers(0, spread, size=(n, 3)).astype(np.int64)
        pd = sharding.PeerShardDedup(m, None, max_local_items=max(n, 1))
        got = pd(torch.from_numpy(codes).to(DEV), [256, 256, 256]).cpu().numpy()
        pd.close()
        assert np.array_equal(got, oracle.suffix_dedup(codes))


def test_suffix_dedup_shipped_artifact():
    arr = np.load(os.path.join(os.path.dirname(__file__), "golden", "course_semantic_ids.npy")).astype(np.int64)
    out, stats = rq.suffix_dedup(None, torch.from_numpy(arr[:, :3].copy()).to(DEV), [8, 8, 8])
    assert np.array_equal(out.cpu().numpy(), arr)
    assert stats["distinct"] == 167 and stats["max_conflicts"] == 18


def test_collision_groups_match_oracle(oracle):
    g, cfg, cbs = load_golden("c1_slice")
    m = build_model(cfg, cbs)
    m._sync()
    codes = g["codes"].astype(np.int64)
    items, offsets, max_group = rq.collision_groups(m, torch.from_numpy(codes).to(DEV))
    items, offsets = items.cpu().numpy(), offsets.cpu().numpy()
    mine = sorted(tuple(items[offsets[i]:offsets[i + 1]]) for i in range(len(offsets) - 1))
    ref = sorted(tuple(int(v) for v in grp) for grp in oracle.collision_groups(codes))
    assert mine == ref
    assert max_group == max(len(grp) for grp in ref)


def test_sinkhorn_cases_match_reference_golden(oracle):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sinkhorn_cases.npz"))
    lib = _cabi.lib()
    for i in range(int(g["n_cases"])):
        B, K, iters = [int(v) for v in g[f"meta{i}"]]
        eps = float(g[f"eps{i}"])
        d = oracle.quantize(g[f"r{i}"], [g[f"cb{i}"]], want_xq=False, dist_level=0, threads=1)[3]
        dt = torch.from_numpy(d).to(DEV)
        scratch = torch.empty((B, K), dtype=torch.float64, device=DEV)
        idx = torch.empty((B,), dtype=torch.int64, device=DEV)
        _cabi.check(lib.rqb200_sinkhorn_assign(dt.data_ptr(), B, K, eps, iters, scratch.data_ptr(), idx.data_ptr(),
                                               _cabi.stream_ptr()))
        assert np.array_equal(idx.cpu().numpy(), g[f"idx{i}"].astype(np.int64)), i
        # sinkhorn_algorithm drop-in: Q close to the fp64 oracle
        c = oracle.center_distance(d).astype(np.float64)
        Q = rq.sinkhorn_algorithm(torch.from_numpy(c).to(DEV), eps, iters).cpu().numpy()
        Qo = oracle.sinkhorn(c, eps, iters)
        assert np.allclose(Q, Qo, rtol=1e-9, atol=1e-300)


def test_sinkhorn_division_is_the_ieee_quotient():
    """The Sinkhorn kernels divide through a correctly rounded reciprocal + two fma corrections (csrc/sinkhorn.cu: div_by);
    the quotient must have the bits of `a / b`: 2^32 operand pairs per exponent spread, hard denominators included."""
    bad = ctypes.c_uint64(0)
    for seed, span in ((1, 2), (2, 40), (3, 300), (4, 900)):
        _cabi.check(_cabi.lib().rqb200_debug_check_division(seed, span, 1 << 32, ctypes.byref(bad), _cabi.stream_ptr()))
        assert bad.value == 0, (span, bad.value)


def test_generate_codes_reproduces_reference_infer_exactly(oracle):
    """BASELINE config 1 end to end: 707 items, K=8, 30 Sinkhorn rounds, suffix column — the reference's verbatim
    infer() output on ALL rows, and the reference's codes after EVERY round (no tolerance: integer output)."""
    f = infer_fixture("c1_infer")
    cfg, x, golden, trace = f["cfg"], f["x"], f["semantic_ids"], f["trace"]
    n = f["n"]
    for fast in (False, True):
        m = build_model(cfg, f["codebooks"])
        # batch_size = 64 as in the reference run: its last DataLoader batch has 707 % 64 = 3 rows (small-batch arithmetic)
        out, stats = rq.generate_codes(m, x, fast=fast, batch_size=64)
        out = out.cpu().numpy()
        assert out.shape == (n, 4) and out.dtype == np.int64
        assert np.array_equal(out, golden)
        assert stats["rounds"] == f["rounds"]
        assert np.array_equal(rq.generate_codes(build_model(cfg, f["codebooks"]), x, fast=fast)[0].cpu().numpy(), golden)
    m = build_model(cfg, f["codebooks"])
    for vq in m.rq.vq_layers[:-1]:
        vq.sk_epsilon = 0.0                                                    # infer.py:109-110
    xt = torch.from_numpy(x).to(DEV)
    codes = torch.from_numpy(trace[0].copy()).to(DEV)
    for t in range(f["rounds"]):
        assert rq.generate_code.reencode_round(m, codes, xt)[0] > 0
        assert np.array_equal(codes.cpu().numpy(), trace[t + 1]), t


@pytest.mark.parametrize("name", ["c2_infer", "c3_infer"])
def test_reencode_rounds_against_reference_infer_at_catalogue_shapes(oracle, name):
    """BASELINE config 2 / 3 shapes (60 000 / 20 000 items, groups of 2..80 rows) through the unmodified reference:
    every round of the product, started from the reference's codes before that round, gives the reference's codes
    after it — except on rows whose two best codes tie in the reference's own fp64 Sinkhorn matrix (the fixture's
    tie set; asserted as a subset, exactly), and never on the arg-min levels."""
    f = infer_fixture(name)
    cfg, x, trace, ties = f["cfg"], f["x"], f["trace"], f["ties"]
    m = build_model(cfg, f["codebooks"])
    xt = torch.from_numpy(x).to(DEV)
    pass1 = m.get_indices(xt, use_sk=False).cpu().numpy()
    assert np.array_equal(pass1, trace[0])
    for vq in m.rq.vq_layers[:-1]:
        vq.sk_epsilon = 0.0
    differing = 0
    for t in range(f["rounds"]):
        codes = torch.from_numpy(trace[t].copy()).to(DEV)
        rq.generate_code.reencode_round(m, codes, xt)
        got = codes.cpu().numpy()
        assert np.array_equal(got[:, :-1], trace[t + 1][:, :-1]), t
        diff = np.nonzero(got[:, -1] != trace[t + 1][:, -1])[0]
        assert np.isin(diff, ties[t]).all(), (t, diff[:10])
        differing += len(diff)
    assert differing <= sum(len(v) for v in ties)
    # host-resident catalogue (members gathered on the host per round): same arithmetic
    codes_d = torch.from_numpy(trace[0].copy()).to(DEV)
    codes_h = codes_d.clone()
    rq.generate_code.reencode_round(m, codes_d, xt)
    rq.generate_code.reencode_round(m, codes_h, x)
    assert torch.equal(codes_d, codes_h)
    # the memo of the per-item part (rqb200_reencode_groups_memo) changes nothing: a chain of rounds with and without it
    memo = rq.generate_code._ReencodeMemo(m, xt.shape[0], xt.device)
    codes_m = torch.from_numpy(trace[0].copy()).to(DEV)
    codes_p = codes_m.clone()
    for t in range(min(6, f["rounds"])):
        rq.generate_code.reencode_round(m, codes_m, xt, memo=memo)
        rq.generate_code.reencode_round(m, codes_p, xt)
        assert torch.equal(codes_m, codes_p), t
    assert int((memo.have != 0).sum()) > 0


def test_generate_codes_matches_oracle_driver_at_scale(oracle):
    g, cfg, cbs = load_golden("c2_slice")
    cfg = dict(cfg, sk_epsilons=[0.0, 0.0, 0.003])
    m = build_model(cfg, cbs)
    _, (ew, eb), _ = synth_weights(cfg)
    n = 6000
    x = synth.synth_items(2024, 0, n, 768, 1_000_000)
    x[3000:3040] = x[100:140]                                                  # exact duplicates → real collision groups
    out, stats = rq.generate_codes(m, x)
    ref, rstats = oracle.generate_codes(x, ew, eb, cbs, cfg["sk_epsilons"], cfg["sk_iters"], group_order=True)
    assert stats["rounds"] == rstats["rounds"]
    assert np.array_equal(out.cpu().numpy(), ref)
    assert len(np.unique(ref, axis=0)) == n


def test_generate_codes_fast_and_exact_pass1_agree():
    """The driver's default (tensor-core pass 1) gives the same ids as the all-exact driver, Sinkhorn rounds included."""
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    n = 40_000
    x = synth.synth_items(2024, 0, n, 768, 1_000_000)
    x[20_000:20_300] = x[500:800]                                              # exact duplicates → collision groups
    xt = torch.from_numpy(x).to(DEV)
    fast, fs = rq.generate_codes(m, xt, fast=True)
    m2 = build_model(cfg, cbs)
    exact, es = rq.generate_codes(m2, xt, fast=False)
    assert fs["pass1_route"] == "tensor-core" and es["pass1_route"] == "exact"
    assert fs["rounds"] == es["rounds"] and fs["rounds"] >= 1
    assert torch.equal(fast, exact)
    host, _ = rq.generate_codes(build_model(cfg, cbs), x, fast=True)           # host (numpy) catalogue: same result
    assert torch.equal(host, exact)


def test_kmeans_lloyd_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    centres = rng.standard_normal((16, 32)).astype(np.float32)
    x = (centres[rng.integers(0, 16, size=20000)] + 0.1 * rng.standard_normal((20000, 32))).astype(np.float32)
    init = x[rng.choice(20000, size=16, replace=False)].copy()
    got = rq.kmeans(torch.from_numpy(x).to(DEV), 16, 10, init=torch.from_numpy(init), tol=0.0).cpu().numpy()
    ref = oracle.kmeans_lloyd(x, init, 10, tol=0.0)
    # Statistical parity only (SURVEY.md §7 hard part 6): the GPU assigns with the quantizer's fp32 distances, the
    # oracle with fp64, so a handful of boundary points may switch cluster.  Tolerance: 2e-2 abs on centres, 1e-3 rel inertia.
    assert np.abs(got - ref).max() <= 2e-2, np.abs(got - ref).max()
    def inertia(c):
        return ((x[:, None, :].astype(np.float64) - c[None].astype(np.float64)) ** 2).sum(-1).min(1).sum()
    assert abs(inertia(got) - inertia(ref)) <= 1e-3 * inertia(ref)
    seeded = rq.kmeans(torch.from_numpy(x).to(DEV), 16, 20, seed=1)
    assert tuple(seeded.shape) == (16, 32) and seeded.is_cuda
    d = ((x[:, None, :] - seeded.cpu().numpy()[None]) ** 2).sum(-1).min(1).mean()
    assert d < 0.01 * 32 * 2                                                    # found the planted clusters
    with pytest.raises(ValueError, match="should be >= n_clusters"):
        rq.kmeans(torch.from_numpy(x[:10]).to(DEV), 16, 5)


def test_gpu_lloyd_matches_scikit_learn_golden():
    """The GPU Lloyd loop from a given init against scikit-learn's own centres (tests/golden/kmeans_sklearn.npz,
    oracle/make_golden_kmeans.py): codebook shapes of BASELINE configs 2, 3, 1 and 5; case 3 exercises the relocation of
    empty clusters.  Centres compared as sets, tolerance = fp32 rounding."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmeans_sklearn.npz"))
    srt = lambda a: a[np.lexsort(a.T[::-1])]
    for i in range(int(g["n_cases"])):
        n, e, K, iters, n_total = (int(v) for v in g[f"meta{i}"])
        x = synth.synth_items(2024, 1, n, e, n_total)
        init = np.ascontiguousarray(x[(np.arange(K) * (n // K)) % n])
        if n == 3000:
            init[1::32] = init[0::32]      # duplicated initial centres → empty clusters (as in make_golden_kmeans.py)
        got = rq.kmeans(torch.from_numpy(x).to(DEV), K, iters, init=torch.from_numpy(init)).cpu().numpy()
        want = g[f"centers{i}"]
        rel = np.abs(srt(got) - srt(want)).max() / np.abs(want).max()
        assert rel <= 1e-5, (i, rel)


def test_kmeans_init_path_of_the_model():
    cfg = dict(in_dim=768, num_emb_list=[64, 64], e_dim=32, layers=[256, 128], sk_epsilons=[0.0, 0.0], sk_iters=5)
    m = rq.RQVAE(in_dim=768, num_emb_list=[64, 64], e_dim=32, layers=[256, 128], kmeans_init=True, kmeans_iters=5,
                 sk_epsilons=[0.0, 0.0]).to(DEV)
    m.train()
    x = torch.from_numpy(synth.synth_items(1, 1, 2048, 768, 100000)).to(DEV)
    out, loss, idx = m(x, use_sk=False)
    assert all(q.initted for q in m.rq.vq_layers)
    assert float(m.rq.vq_layers[0].embedding.weight.abs().sum()) > 0
    m.eval()
    assert tuple(m.get_indices(x).shape) == (2048, 2)


def test_host_buffer_end_to_end_call(oracle):
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    m._sync()
    _, (ew, eb), _ = synth_weights(cfg)
    n = 5000
    x = synth.synth_items(2024, 0, n, 768, 1_000_000)
    xh = torch.from_numpy(x).pin_memory()
    out = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    stats = (ctypes.c_int64 * 4)()
    _cabi.check(_cabi.lib().rqb200_generate_codes_host(m._handle, _cabi.ENCODE_EXACT, xh.data_ptr(), n, 1024,
                                                       out.data_ptr(), stats, _cabi.stream_ptr()))
    ref = oracle.suffix_dedup(oracle.get_indices(x, ew, eb, cbs))
    assert np.array_equal(out.numpy(), ref)
    side = torch.cuda.Stream()                                  # the call runs on the stream it is given
    out2 = torch.empty_like(out).pin_memory()
    _cabi.check(_cabi.lib().rqb200_generate_codes_host(m._handle, _cabi.ENCODE_FAST, xh.data_ptr(), n, 1000,
                                                       out2.data_ptr(), stats, ctypes.c_void_p(side.cuda_stream)))
    assert np.array_equal(out2.numpy(), ref)
    assert stats[1] == len(np.unique(ref[:, :3], axis=0))


def test_full_size_properties_c2(oracle):
    """BASELINE config 2 at full size (1M x 768, 3 x 256, e 32): sampled parity + size-independent properties."""
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    _, (ew, eb), _ = synth_weights(cfg)
    n = 1_000_000
    x = gpu_synth(2024, 0, n, 768, n)
    codes = m.get_indices(x)
    out, stats = rq.suffix_dedup(m, codes)
    o = out.cpu().numpy()
    assert np.array_equal(o[:8192, :3], g["codes"].astype(np.int64))             # golden head of the same catalogue
    sel = np.random.default_rng(0).choice(n, size=20000, replace=False)
    sel.sort()
    xs = x[torch.from_numpy(sel).to(DEV)].cpu().numpy()
    assert np.array_equal(o[sel, :3], oracle.get_indices(xs, ew, eb, cbs))       # sampled bit-exact parity
    # uniqueness, suffix rule via checksums: per-key count == max suffix + 1
    key = (o[:, 0] * 256 + o[:, 1]) * 256 + o[:, 2]
    full = key * (int(o[:, 3].max()) + 1) + o[:, 3]
    assert len(np.unique(full)) == n
    uk, cnt = np.unique(key, return_counts=True)
    assert stats["distinct"] == len(uk) and stats["max_conflicts"] == cnt.max()
    order = np.argsort(key, kind="stable")
    starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    expect = np.arange(n) - np.repeat(starts, cnt)
    assert np.array_equal(o[order, 3], expect)
    # idempotence: re-encoding a shuffled copy gives the permuted codes
    perm = torch.randperm(n, device=DEV)[:200_000]
    assert torch.equal(m.get_indices(x[perm]), codes[perm])

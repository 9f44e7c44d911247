"""GPU: the tcgen05 encoder and the margin-gated fast route give the same codes as the exact route."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import build_model, load_golden, synth_weights
from ai_education_generative_recommendation_b200 import _cabi, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_synth(seed, first, n, dim, n_total):
    x = torch.empty((n, dim), dtype=torch.float32, device=DEV)
    _cabi.check(_cabi.lib().rqb200_synth_items(seed, first, n, dim, n_total, x.data_ptr(), _cabi.stream_ptr()))
    return x


@pytest.mark.parametrize("name", ["c2_slice", "c3_slice", "c5_slice"])
def test_tensor_core_encoder_is_fp32_class(name):
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    n = 50_000 + 77                                   # ragged last tile
    x = gpu_synth(2024, 0, n, cfg["in_dim"], int(g["n_total"]))
    z = m.encoder(x)                                  # exact SIMT route (bit-equal to the reference)
    zt = m.encode_tc(x)
    err = (zt - z).norm(dim=1) / (z.norm(dim=1) + 1e-3)
    assert torch.isfinite(zt).all()
    # tolerance: split-fp16 GEMM with truncating fp32 accumulation; must stay well inside the gate's 2^-15 bound
    assert float(err.max()) < 2.0 ** -16, float(err.max())


@pytest.mark.parametrize("name", ["c1_slice", "c2_slice", "c3_slice", "c5_slice"])
def test_fast_route_codes_equal_exact_route(name):
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    n_gold = int(g["n_rows"])
    xg = torch.from_numpy(synth.synth_items(int(g["seed"]), 0, n_gold, cfg["in_dim"], int(g["n_total"]))).to(DEV)
    m.encode_mode = _cabi.ENCODE_FAST
    got = m.get_indices(xg)
    assert np.array_equal(got.cpu().numpy(), g["codes"].astype(np.int64))       # reference golden, bit-exact codes
    n = 300_000
    x = gpu_synth(2024, 0, n, cfg["in_dim"], int(g["n_total"]))
    fast = m.get_indices(x)
    rescued = m.last_stats["rescued_rows"]
    m.encode_mode = _cabi.ENCODE_EXACT
    exact = m.get_indices(x)
    assert torch.equal(fast, exact)
    if name != "c1_slice":                      # K=8 codebooks on 707-item clusters: many near-ties by construction
        assert rescued <= 0.05 * n, rescued


@pytest.mark.parametrize("e_dim", [32, 64])
def test_fast_route_with_ragged_codebooks(e_dim):
    """Codebook sizes that are no multiple of the 128-code chunk of the tensor-core quantizer (its last chunk then scans
    32 / 64 / 96 columns through the rolled path), five levels, both e_dim the tensor-core encoder ends in: fast route ==
    exact route."""
    Ks = [100, 7, 300, 5, 200]
    cfg = {"in_dim": 768, "layers": [256, 128], "e_dim": e_dim, "num_emb_list": Ks, "sk_epsilons": [0.0] * len(Ks), "sk_iters": 50}
    rng = np.random.default_rng(e_dim)
    probe = build_model(cfg, [np.zeros((k, e_dim), np.float32) for k in Ks])
    n = 60_000 + 19
    x = gpu_synth(2024, 0, n, cfg["in_dim"], 1_000_000)
    z = probe.encoder(x[:4096]).cpu().numpy()
    scale = float(z.std())
    cbs = [z[rng.permutation(4096)[:Ks[0]]].copy()]                       # level 0: latents themselves (rows ON codes)
    for l in range(1, len(Ks)):
        cbs.append((rng.standard_normal((Ks[l], e_dim)) * scale * 0.6 ** l).astype(np.float32))
    m = build_model(cfg, cbs)
    m.encode_mode = _cabi.ENCODE_FAST
    fast = m.get_indices(x)
    stats = m.last_stats
    m.encode_mode = _cabi.ENCODE_EXACT
    exact = m.get_indices(x)
    assert torch.equal(fast, exact)
    assert int(fast.max()) < max(Ks) and all(int(fast[:, l].max()) < Ks[l] for l in range(len(Ks)))
    assert stats["rescued_rows"] < n                                      # the gate certified rows: the fast kernels did run


@pytest.mark.parametrize("name", ["c2_slice", "c3_slice", "c5_slice"])
def test_screening_tier_keeps_codes_bit_exact(name):
    """Opt-in tier 1 (one fp16 pass, looser gate) → three-pass re-run of the gated rows → exact rescue: same codes."""
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    n = 200_000 + 13
    x = gpu_synth(2024, 0, n, cfg["in_dim"], int(g["n_total"]))
    exact = m.get_indices(x)
    m.encode_mode = _cabi.ENCODE_FAST
    m.set_screen(True, 2.0 ** -11)
    fast = m.get_indices(x)
    st = dict(m.last_stats)
    m.set_screen(False)
    m.encode_mode = _cabi.ENCODE_EXACT
    assert torch.equal(fast, exact)
    assert 0 < st["three_pass_rows"] < 0.5 * n, st          # the screening pass certifies most rows by itself
    assert st["rescued_rows"] <= st["three_pass_rows"]


def test_fast_route_handles_overflowing_rows():
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    x = gpu_synth(2024, 0, 4096, 768, 1_000_000)
    x[5] *= 1.0e6                                   # beyond fp16 range → must be rescued, not mis-coded
    x[77, 3] = float("inf")
    m.encode_mode = _cabi.ENCODE_FAST
    fast = m.get_indices(x)
    m.encode_mode = _cabi.ENCODE_EXACT
    assert torch.equal(fast, m.get_indices(x))


def test_one_cta_and_two_cta_kernels_agree():
    """The 2-CTA (cta_group::2) first-layer kernel and the 1-CTA kernel implement the same arithmetic."""
    import os, subprocess, sys
    code = ("import sys, torch; sys.path.insert(0, 'tests'); from conftest import build_model, load_golden;"
            "from ai_education_generative_recommendation_b200 import _cabi;"
            "g, cfg, cbs = load_golden('c2_slice'); m = build_model(cfg, cbs);"
            "x = torch.empty((70001, 768), device='cuda:0');"
            "_cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, 70001, 768, 1000000, x.data_ptr(), _cabi.stream_ptr()));"
            "z = m.encode_tc(x); torch.cuda.synchronize(); torch.save(z.cpu(), sys.argv[1])")
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as tmp:
        for flag in ("0", "1"):
            env = dict(os.environ, RQB200_TC2=flag)
            path = os.path.join(tmp, f"z{flag}.pt")
            r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300,
                               cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(torch.load(path))
    assert torch.equal(outs[0], outs[1])          # same K order, same MMA sequence ⇒ identical bits


@pytest.mark.parametrize("name", ["c2_slice", "c3_slice"])
def test_fast_route_enqueues_without_waiting_and_replays_from_a_cuda_graph(name):
    """rqb200_get_indices(FAST, stats = NULL) keeps the row counts of its re-run tiers on the device (the exact tier runs
    persistent grids over device-counted rows), so the whole call can be captured once and replayed: same codes as the
    exact route, also on new data written into the captured input buffer."""
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    n = 200_000 + 13
    lib = _cabi.lib()
    x = gpu_synth(2024, 0, n, cfg["in_dim"], int(g["n_total"]))
    m.encode_mode = _cabi.ENCODE_EXACT
    want = m.get_indices(x)
    m._sync()
    codes = torch.empty((n, len(cbs)), dtype=torch.int64, device=DEV)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):                             # warm-up: workspaces, function attributes, weight images
            _cabi.check(lib.rqb200_get_indices(m._handle, _cabi.ENCODE_FAST, x.data_ptr(), n, codes.data_ptr(), 0, None,
                                               _cabi.stream_ptr()))
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert torch.equal(codes, want)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        _cabi.check(lib.rqb200_get_indices(m._handle, _cabi.ENCODE_FAST, x.data_ptr(), n, codes.data_ptr(), 0, None,
                                           _cabi.stream_ptr()))
    codes.fill_(-1)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(codes, want)
    x.copy_(gpu_synth(7, 5, n, cfg["in_dim"], int(g["n_total"])))          # other rows, another number of gated rows
    m.encode_mode = _cabi.ENCODE_EXACT
    want2 = m.get_indices(x)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(codes, want2)
    tiers = (ctypes.c_int64 * 2)()
    _cabi.check(lib.rqb200_model_last_tier_rows(m._handle, tiers))
    assert 0 <= tiers[1] <= n and 0 <= tiers[0] <= n

"""GPU: the gate of the tensor-core route at catalogue scale, and the small-batch arithmetic of the module-level API.

The fast route (TF32 screening tier → split-fp16 three-pass tier → exact SIMT tier, each certifying the rows it keeps
with the quantizer's margin gate) must give the SAME integer codes as the exact route.  The gate bounds are calibrated
(tools/calibrate_gate.py), not derived, so they are checked here where it is cheap: 10 M rows per BASELINE shape, rows
pushed onto code boundaries, and rows of small magnitude.  Every comparison is exact (integer output).
"""
import os

import numpy as np
import pytest
import torch

from conftest import build_model, load_golden
from ai_education_generative_recommendation_b200 import _cabi
import ai_education_generative_recommendation_b200 as rq

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_synth(seed, first, n, dim, n_total):
    x = torch.empty((n, dim), dtype=torch.float32, device=DEV)
    _cabi.check(_cabi.lib().rqb200_synth_items(seed, first, n, dim, n_total, x.data_ptr(), _cabi.stream_ptr()))
    return x


def both_routes(m, x):
    m.encode_mode = _cabi.ENCODE_EXACT
    exact = m.get_indices(x, use_sk=False)
    m.encode_mode = _cabi.ENCODE_FAST
    fast = m.get_indices(x, use_sk=False)
    return exact, fast, dict(m.last_stats)


@pytest.mark.parametrize("name,chunks", [("c2_slice", 10), ("c3_slice", 10), ("c5_slice", 10)])
@pytest.mark.parametrize("screen", ["default", "tf32", "off"])
def test_fast_route_equals_exact_route_on_10m_rows(name, chunks, screen):
    """10 x 1 M rows cut from different places of the synthetic catalogue (and from catalogues of other seeds): codes of
    the tensor-core route == codes of the exact route on every row."""
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    if not m.fast_route_supported():
        pytest.skip("shapes without a tensor-core route")
    if screen == "tf32":
        m.set_screen("tf32")
    elif screen == "off":
        m.set_screen(0)
    n_total = int(g["n_total"])
    rows = 1_000_000
    rescued = rerun = 0
    for c in range(chunks):
        seed = 2024 + (c % 3)                      # other seeds = other cluster centres and noise
        first = (c * 7_919_993) % max(n_total - rows, 1)
        x = gpu_synth(seed, first, rows, cfg["in_dim"], n_total)
        exact, fast, st = both_routes(m, x)
        bad = int((exact != fast).any(1).sum())
        assert bad == 0, f"{name} chunk {c} (seed {seed}, first row {first}): {bad} rows differ between the routes"
        rescued += st["rescued_rows"]
        rerun += st["three_pass_rows"]
    print(f"{name} screen={screen}: {chunks * rows} rows, {rerun} re-run by the three-pass tier, {rescued} by the exact tier")
    assert 0 < rescued < 0.1 * chunks * rows


@pytest.mark.parametrize("name", ["c2_slice", "c3_slice", "c5_slice"])
def test_rows_on_code_boundaries(name):
    """Adversarial lane: pairs of items with different codes are joined by a segment and the point where the EXACT code
    changes is located by bisection to the last representable step; the fast route is then asked for the codes of the
    points at the boundary and a few ulps to millionths on either side of it.  These are the rows a wrong gate loses."""
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    if not m.fast_route_supported():
        pytest.skip("shapes without a tensor-core route")
    m.set_screen("tf32")
    pairs = 4096
    xa = gpu_synth(2024, 1, pairs, cfg["in_dim"], int(g["n_total"]))
    xb = gpu_synth(2024, 500_001, pairs, cfg["in_dim"], int(g["n_total"]))
    m.encode_mode = _cabi.ENCODE_EXACT
    ca = m.get_indices(xa, use_sk=False)
    lo = torch.zeros((pairs, 1), dtype=torch.float32, device=DEV)
    hi = torch.ones((pairs, 1), dtype=torch.float32, device=DEV)
    for _ in range(30):                               # code(lo) == code(a) stays invariant
        mid = 0.5 * (lo + hi)
        cm = m.get_indices(xa + mid * (xb - xa), use_sk=False)
        same = (cm == ca).all(1, keepdim=True)
        lo = torch.where(same, mid, lo)
        hi = torch.where(same, hi, mid)
    deltas = torch.tensor([0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4, 1e-3, -1e-3], device=DEV)
    t = torch.cat([(lo + d).clamp(0, 1) for d in deltas] + [hi.clone()], 0)
    x = xa.repeat(len(deltas) + 1, 1) + t * (xb - xa).repeat(len(deltas) + 1, 1)
    exact, fast, st = both_routes(m, x.contiguous())
    bad = int((exact != fast).any(1).sum())
    assert bad == 0, f"{bad} of {x.shape[0]} boundary rows differ between the routes"
    assert len(torch.unique(exact, dim=0)) > pairs // 4              # the lane really straddles boundaries
    print(f"{name}: {x.shape[0]} boundary rows, {st['three_pass_rows']} re-run by the three-pass tier, {st['rescued_rows']} by the exact tier")


@pytest.mark.parametrize("name", ["c2_slice", "c5_slice"])
@pytest.mark.parametrize("scale_exp", [0, -4, -8, -12, -16, -20, 6])
def test_rows_of_small_and_large_magnitude(name, scale_exp):
    """The split-fp16 operands are not rescaled per row: inputs far below 1 lose their low halves to fp16 underflow, so
    the gate has to send them to the exact tier (or they must still be right).  Either way the codes must not differ."""
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    if not m.fast_route_supported():
        pytest.skip("shapes without a tensor-core route")
    m.set_screen("tf32")
    x = gpu_synth(2024, 0, 200_000, cfg["in_dim"], int(g["n_total"])) * (2.0 ** scale_exp)
    exact, fast, st = both_routes(m, x.contiguous())
    bad = int((exact != fast).any(1).sum())
    assert bad == 0, f"scale 2^{scale_exp}: {bad} rows differ ({st})"


def test_module_api_on_small_batches_matches_reference_golden():
    """RQVAE.encoder / get_indices / forward on batches of 2 … 20 rows: the reference's CPU GEMM uses its small-batch
    summation order below 16 rows (csrc/small_batch.cu); latent bits, codes (with and without Sinkhorn on the last level)
    and decoded bits equal the reference's for every batch size (tests/golden/small_batch.npz, oracle/make_golden.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_batch.npz"))
    first = int(g["first_row"])
    for tag, base in (("c2", "c2_slice"), ("c3", "c3_slice")):
        gg, cfg, cbs = load_golden(base)
        from ai_education_generative_recommendation_b200 import synth
        x_all = synth.synth_items(2024, 0, 4096, cfg["in_dim"], int(gg["n_total"]))
        for key in [k for k in g["names"] if str(k).startswith(tag)]:
            key = str(key)
            M = int(key.split("_m")[1])
            x = torch.from_numpy(np.ascontiguousarray(x_all[first:first + M])).to(DEV)
            m = build_model(dict(cfg, sk_epsilons=[0.0] * len(cbs)), cbs)
            for mode in (_cabi.ENCODE_EXACT, _cabi.ENCODE_FAST):          # < 16 rows: the fast route defers to the exact kernels
                m.encode_mode = mode
                z = m.encoder(x)
                assert np.array_equal(z.cpu().numpy().view(np.int32), g[f"{key}_z"].view(np.int32)), (key, "latent")
                assert np.array_equal(m.get_indices(x, use_sk=False).cpu().numpy(), g[f"{key}_codes"].astype(np.int64)), (key, mode)
            out, rq_loss, idx = m(x, use_sk=False)
            assert np.array_equal(idx.cpu().numpy(), g[f"{key}_codes"].astype(np.int64)), (key, "forward codes")
            assert np.array_equal(out.cpu().numpy()[:, :16].view(np.int32), g[f"{key}_out"].view(np.int32)), (key, "decoded bits")
            assert abs(float(rq_loss) - float(g[f"{key}_rq_loss"])) <= 1e-5 * abs(float(g[f"{key}_rq_loss"])), key
            msk = build_model(dict(cfg, sk_epsilons=[0.0] * (len(cbs) - 1) + [0.003]), cbs)
            got = msk.get_indices(x, use_sk=True).cpu().numpy()
            assert np.array_equal(got, g[f"{key}_codes_sk"].astype(np.int64)), (key, "use_sk codes")

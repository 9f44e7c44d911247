"""Measures the tensor-core encoder's error and the margin gate's behaviour on the GPU (diagnostic)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden          # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi   # noqa: E402

DEV = "cuda:0"
# RQB200_TF32_COMP (units of 2^-11, read when the TF32 weight image is packed): compensation of the operand truncation
comps = [c for c in os.environ.get("CALIBRATE_COMPS", "").split(",") if c] or [None]
names = [a for a in sys.argv[1:]] or ["c2_slice", "c3_slice", "c5_slice"]
for name, comp in [(nm, c) for nm in names for c in comps]:
    if comp is not None:
        os.environ["RQB200_TF32_COMP"] = comp
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    n = 1_000_000 if cfg["in_dim"] == 768 else 400_000
    x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device=DEV)
    _cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, cfg["in_dim"], int(g["n_total"]), x.data_ptr(), _cabi.stream_ptr()))
    z = m.encoder(x)
    zt = m.encode_tc(x)
    rel = (zt - z).norm(dim=1) / (z.norm(dim=1) + 1e-3)
    # is the error mostly a uniform shrink (truncating accumulation)?  best scalar fit z ≈ s * z~
    s_opt = float((zt.double() * z.double()).sum() / (zt.double() * zt.double()).sum())
    rel2 = (zt * s_opt - z).norm(dim=1) / (z.norm(dim=1) + 1e-3)
    zs = m.encode_tc(x, passes=2)                      # the screening tier's latent: TF32 first layer + three-pass tail
    rel_s = (zs - z).norm(dim=1) / (z.norm(dim=1) + 1e-3)
    exact = m.get_indices(x)
    ss = float((zs.double() * z.double()).sum() / (zs.double() * zs.double()).sum())
    out = {"config": name, "tf32_comp_in_2^-11": comp, "tf32_best_residual_scale_minus_1": ss - 1.0, "tf32_screen_rel_err_max": float(rel_s.max()), "tf32_screen_rel_err_mean": float(rel_s.mean()),
           "tf32_screen_rel_err_log2_max": float(torch.log2(rel_s.max())), "rows": n, "rel_err_max": float(rel.max()), "rel_err_mean": float(rel.mean()),
           "rel_err_log2_max": float(torch.log2(rel.max())), "best_scale_minus_1": s_opt - 1.0,
           "rel_err_after_scale_max": float(rel2.max()), "rel_err_after_scale_mean": float(rel2.mean())}
    m.encode_mode = _cabi.ENCODE_FAST
    m.set_screen(False)
    for gamma_log2 in (-30, -19, -17, -16, -15, -14):
        m.set_gate(2.0 ** gamma_log2 if gamma_log2 > -30 else 0.0, 1e-3)
        fast = m.get_indices(x)
        out[f"gamma=2^{gamma_log2}"] = {"rescued": m.last_stats["rescued_rows"],
                                        "mismatching_rows": int((fast != exact).any(1).sum())}
    # screening tier (TF32 first layer): rows sent on to the three-pass tier, rows mis-coded if the tight gate were off
    m.set_gate(2.0 ** -15, 1e-3)
    for g1 in (-16, -15, -14, -13.5, -13, -12.5, -12, -11.5, -11, -10):
        m.set_screen("tf32", 2.0 ** g1)
        fast = m.get_indices(x)
        out[f"screen=2^{g1}"] = {"three_pass_rows": m.last_stats["three_pass_rows"], "rescued": m.last_stats["rescued_rows"],
                                 "mismatching_rows": int((fast != exact).any(1).sum())}
    print(json.dumps(out))

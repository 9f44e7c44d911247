"""2-CTA layer-1 kernel vs the 1-CTA kernel (diagnostic; run with RQB200_TC2=1)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden
from ai_education_generative_recommendation_b200 import _cabi
g, cfg, cbs = load_golden("c2_slice")
m = build_model(cfg, cbs)
for n in (256, 1000, 128 * 148 * 2 + 77, 1_000_000):
    x = torch.empty((n, 768), dtype=torch.float32, device="cuda:0")
    _cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, 768, 1_000_000, x.data_ptr(), _cabi.stream_ptr()))
    z = m.encoder(x)
    zt = m.encode_tc(x); torch.cuda.synchronize()
    rel = (zt - z).norm(dim=1) / (z.norm(dim=1) + 1e-3)
    print(n, "rel err max", float(rel.max()), "finite", bool(torch.isfinite(zt).all()), flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m.encode_tc(x)
e1.record(); torch.cuda.synchronize()
print("encoder (3 layers) ms:", e0.elapsed_time(e1) / 5)

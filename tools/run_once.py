"""Runs the fast get_indices route a few times on a small catalogue (profiling target for ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden          # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g, cfg, cbs = load_golden(name)
m = build_model(cfg, cbs)
x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device="cuda:0")
_cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, cfg["in_dim"], int(g["n_total"]), x.data_ptr(), _cabi.stream_ptr()))
m.encode_mode = _cabi.ENCODE_FAST
for _ in range(reps):
    codes = m.get_indices(x)
torch.cuda.synchronize()
print("ok", m.last_stats, int(codes.sum()))

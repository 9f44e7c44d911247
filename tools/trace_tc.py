"""Timeline of the tensor-core linear kernel's pipeline (CTA 0): who waits for whom (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden
from ai_education_generative_recommendation_b200 import _cabi
g, cfg, cbs = load_golden("c2_slice")
m = build_model(cfg, cbs)
n = 1_000_000
x = torch.empty((n, 768), dtype=torch.float32, device="cuda:0")
_cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, 768, n, x.data_ptr(), _cabi.stream_ptr()))
m.encode_tc(x); torch.cuda.synchronize()
buf = torch.zeros((8, 256), dtype=torch.int64, device="cuda:0")
_cabi.check(_cabi.lib().rqb200_debug_tc_trace(buf.data_ptr()))
# only the first layer: run the full MLP, the later layers overwrite the trace → use a 1-layer view by timing order
m.encode_tc(x); torch.cuda.synchronize()
_cabi.check(_cabi.lib().rqb200_debug_tc_trace(0))
t = buf.cpu().numpy()
names = ["conv:empty-ok", "conv:arrived", "mma:full_a-ok", "mma:full_w-ok", "mma:commit", "load:empty-ok", "epi:full-ok", "epi:done"]
print("trace of the N=256 kernel (encoder layer 1, K=768: 12 slabs per tile)")
t0 = t[t > 0].min()
for i in range(0, 40):
    print(i, " ".join(f"{names[k]}={t[k, i] - t0:7d}" for k in range(6) if t[k, i] > 0))
for i in range(0, 8):
    print("tile", i, " ".join(f"{names[k]}={t[k, i] - t0:7d}" for k in (6, 7) if t[k, i] > 0))

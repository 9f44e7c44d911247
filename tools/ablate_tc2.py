"""Which resource bounds the 2-CTA first-layer kernel?  Times the 768->256 layer with parts switched off (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden
from ai_education_generative_recommendation_b200 import _cabi
lib = _cabi.lib()
n = 1_000_000
g, cfg, cbs = load_golden("c2_slice")
m = build_model(cfg, cbs); m._sync()
x = torch.empty((n, 768), dtype=torch.float32, device="cuda:0")
_cabi.check(lib.rqb200_synth_items(2024, 0, n, 768, n, x.data_ptr(), _cabi.stream_ptr()))
y = torch.empty((n, 256), dtype=torch.float32, device="cuda:0")
def run(flags, passes, reps=5):
    lib.rqb200_debug_tc_flags(flags)
    call = lambda: _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), n, y.data_ptr(), passes, 1, _cabi.stream_ptr()))
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
names = {0: "production", 1: "no epilogue stores", 2: "no MMA", 4: "no producer STS", 8: "no W bulk loads", 16: "no X loads",
         1 | 8: "no epi stores + no W", 1 | 2 | 8: "no epi/MMA/W (X stream + convert only)",
         1 | 2 | 4 | 8: "X loads only", 1 | 4 | 16: "MMA + W only", 1 | 4 | 8 | 16: "MMA only", 2 | 4 | 16 | 1: "W loads only",
         2 | 4 | 8 | 16: "epilogue stores only", 1 | 2 | 4 | 8 | 16: "nothing (barrier skeleton)",
         31 | 32: "skeleton, no producer fence.proxy.async", 31 | 128: "skeleton, relaxed remote arrive",
         31 | 256: "skeleton, cta-scope peer wait", 31 | 128 | 256: "skeleton, both relaxed",
         128 | 256: "production, both relaxed", 15 | 32: "X loads only, no producer fence",
         11 | 32: "X stream + convert, no producer fence", 32: "production minus producer fence (INVALID results)"}
for passes in (3,):
    for f, nm in names.items():
        print(f"passes={passes} {f:3d} {nm:42s} {run(f, passes):7.3f} ms", flush=True)
lib.rqb200_debug_tc_flags(0)

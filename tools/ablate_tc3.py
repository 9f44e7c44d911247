"""Which resource bounds linear_tc3_kernel (A operand in tensor memory)?  Times the 768->256 layer with parts switched
off, next to linear_tc2_kernel with the same switches (diagnostic; flag 4096 selects linear_tc3_kernel).

Run under a timeout:  timeout 300 python tools/ablate_tc3.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden          # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi   # noqa: E402

lib = _cabi.lib()
n = 1_000_000
g, cfg, cbs = load_golden("c2_slice")
m = build_model(cfg, cbs)
m._sync()
x = torch.empty((n, 768), dtype=torch.float32, device="cuda:0")
_cabi.check(lib.rqb200_synth_items(2024, 0, n, 768, n, x.data_ptr(), _cabi.stream_ptr()))
y = torch.empty((n, 256), dtype=torch.float32, device="cuda:0")


def run(flags, reps=5):
    lib.rqb200_debug_tc_flags(flags)
    call = lambda: _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), n, y.data_ptr(), 3, 1, _cabi.stream_ptr()))  # noqa: E731
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    lib.rqb200_debug_tc_flags(0)
    return e0.elapsed_time(e1) / reps


names = {0: "production", 1: "no epilogue stores (drain only)", 2: "no MMA", 4: "no producer A stores", 8: "no W bulk loads",
         16: "no X loads", 1 | 2 | 8: "X stream + convert + A stores only", 1 | 2 | 4 | 8: "X loads only",
         1 | 4 | 16: "MMA + W only", 1 | 4 | 8 | 16: "MMA only", 1 | 2 | 4 | 16: "W loads only",
         2 | 4 | 8 | 16: "epilogue only", 31: "nothing (barrier skeleton + drain)"}
print(f"{'switches':44s} {'tc2 ms':>8s} {'tc3 ms':>8s}")
for f, nm in names.items():
    print(f"{f:3d} {nm:40s} {run(f):8.3f} {run(f | 4096):8.3f}", flush=True)

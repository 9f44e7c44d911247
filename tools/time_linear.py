"""Times the tensor-core Linear kernels in isolation (CUDA events), passes = 1 / 3, per layer (diagnostic)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden          # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi   # noqa: E402

DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g, cfg, cbs = load_golden(name)
m = build_model(cfg, cbs)
m._sync()
lib = _cabi.lib()
dims = [cfg["in_dim"]] + list(cfg["layers"]) + [cfg["e_dim"]]
x = torch.empty((n, dims[0]), dtype=torch.float32, device=DEV)
_cabi.check(lib.rqb200_synth_items(2024, 0, n, dims[0], int(g["n_total"]), x.data_ptr(), _cabi.stream_ptr()))
cur = x
for layer in range(len(dims) - 1):
    y = torch.empty((n, dims[layer + 1]), dtype=torch.float32, device=DEV)
    for passes in (3, 1):
        for _ in range(3):
            _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, layer, cur.data_ptr(), n, y.data_ptr(), passes, 1, _cabi.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, layer, cur.data_ptr(), n, y.data_ptr(), passes, 1, _cabi.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = n * (dims[layer] + dims[layer + 1]) * 4 / 1e9
        print(f"layer {layer} {dims[layer]}->{dims[layer + 1]} passes={passes}: {ms:.3f} ms  ({gb / ms * 1e3:.0f} GB/s of x+y traffic)", flush=True)
    cur = y

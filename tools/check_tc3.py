"""linear_tc3_kernel (A operand in tensor memory, experimental) against linear_tc2_kernel (diagnostic, GPU box).

Both kernels issue the same MMAs in the same order on the same split-fp16 operands, so every output must be
BIT-IDENTICAL: first the fp32 first-layer output in isolation (rqb200_debug_linear_tc), then the whole tensor-core
encoder (the production hand-off: first layer → split-fp16 tiles → fused layers 2+3).  Then both are timed.
Debug flag 4096 (rqb200_debug_tc_flags) routes the plain three-pass first-layer launch through linear_tc3_kernel.

Run under a timeout (an untested tcgen05 pipeline can hang):  timeout 120 python tools/check_tc3.py [c2_slice|c5_slice]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden          # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi   # noqa: E402

DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
g, cfg, cbs = load_golden(name)
m = build_model(cfg, cbs)
m._sync()
lib = _cabi.lib()
in_dim, h1 = cfg["in_dim"], cfg["layers"][0]
ok = True


def first_layer(x, flags, passes=3):
    y = torch.full((x.shape[0], h1), float("nan"), dtype=torch.float32, device=DEV)
    lib.rqb200_debug_tc_flags(flags)
    _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), x.shape[0], y.data_ptr(), passes, 1, _cabi.stream_ptr()))
    torch.cuda.synchronize()
    lib.rqb200_debug_tc_flags(0)
    return y


def encoder(x, flags):
    lib.rqb200_debug_tc_flags(flags)
    z = m.encode_tc(x)
    torch.cuda.synchronize()
    lib.rqb200_debug_tc_flags(0)
    return z


for n in (1, 255, 256, 257, 1000, 256 * 74 + 77, 256 * 74 * 3 + 5, 1_000_000):
    x = torch.empty((n, in_dim), dtype=torch.float32, device=DEV)
    _cabi.check(lib.rqb200_synth_items(2024, 0, n, in_dim, int(g["n_total"]), x.data_ptr(), _cabi.stream_ptr()))
    y2, y3 = first_layer(x, 0), first_layer(x, 4096)
    same_y = torch.equal(y2.view(torch.int32), y3.view(torch.int32))
    z2, z3 = encoder(x, 0), encoder(x, 4096)
    same_z = torch.equal(z2.view(torch.int32), z3.view(torch.int32))
    bad = int((y2.view(torch.int32) != y3.view(torch.int32)).sum())
    print(f"n={n}: first layer bit-identical={same_y} ({bad} of {y2.numel()} differ), encoder bit-identical={same_z}", flush=True)
    y2s, y3s = first_layer(x, 0, 1), first_layer(x, 4096, 1)          # one-pass (screening tier) instantiations
    same_1 = torch.equal(y2s.view(torch.int32), y3s.view(torch.int32))
    print(f"        one-pass first layer bit-identical={same_1}", flush=True)
    ok &= same_y and same_z and same_1
    if not same_y and n <= 257:
        d = (y2.view(torch.int32) != y3.view(torch.int32)).nonzero()[:8].tolist()
        print("   first differing (row, col):", d, flush=True)

for passes in (3, 1):
    for flags, label in ((0, "linear_tc2_kernel"), (4096, "linear_tc3_kernel")):
        y = torch.empty((x.shape[0], h1), dtype=torch.float32, device=DEV)
        lib.rqb200_debug_tc_flags(flags)
        for _ in range(3):
            _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), x.shape[0], y.data_ptr(), passes, 1, _cabi.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), x.shape[0], y.data_ptr(), passes, 1, _cabi.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        lib.rqb200_debug_tc_flags(0)
        ms = e0.elapsed_time(e1) / 10
        print(f"{label}<{passes}>: {ms:.3f} ms for {x.shape[0]} x {in_dim} -> {h1}  ({x.shape[0] * in_dim * 4 / ms / 1e6:.0f} GB/s of X)", flush=True)
sys.exit(0 if ok else 1)

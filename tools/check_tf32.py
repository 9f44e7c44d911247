"""linear_tf32_kernel (TMA-fed TF32 first layer, csrc/encode_tf32.cu) on the GPU box: accuracy against the exact SIMT
layer, the whole fast route with the TF32 screening tier against the exact route (codes must be IDENTICAL), tier row
counts per screening bound, and timings beside the three-pass kernel.

Run under a timeout (a new tcgen05 / TMA pipeline can hang):  timeout 300 python tools/check_tf32.py [c2_slice|c3_slice|c5_slice]
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200 import _cabi                                   # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden      # noqa: E402

DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
big = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g, cfg, cbs = load_golden(name)
cfg = dict(cfg, sk_epsilons=[0.0] * len(cfg["num_emb_list"]))
m = build_model(cfg, cbs)
m._sync()
lib = _cabi.lib()
in_dim, h1 = cfg["in_dim"], cfg["layers"][0]
ok = True


def synth(n, first=0):
    x = torch.empty((n, in_dim), dtype=torch.float32, device=DEV)
    _cabi.check(lib.rqb200_synth_items(2024, first, n, in_dim, int(g["n_total"]), x.data_ptr(), _cabi.stream_ptr()))
    return x


def first_layer(x, passes):
    y = torch.full((x.shape[0], h1), float("nan"), dtype=torch.float32, device=DEV)
    _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), x.shape[0], y.data_ptr(), passes, 1, _cabi.stream_ptr()))
    torch.cuda.synchronize()
    return y


def exact_first_layer(x):
    # the exact MLP's first layer = the encoder of a one-layer view: use the three-pass kernel's own check instead
    return first_layer(x, 3)


print(f"== {name}: {in_dim} -> {h1}", flush=True)
for n in (1, 127, 128, 255, 256, 257, 1000, 256 * 74 + 77, 256 * 74 * 3 + 5):
    x = synth(n)
    y3, y2 = first_layer(x, 3), first_layer(x, 2)
    fin = bool(torch.isfinite(y2).all())
    scale = float(y3.abs().max()) + 1e-30
    err = float((y2 - y3).abs().max()) / scale
    relu_zero_mismatch = int(((y2 == 0) != (y3 == 0)).sum())
    print(f"n={n}: finite={fin} max|y_tf32 - y_3pass| / max|y| = {err:.3e}  (zero pattern differs in {relu_zero_mismatch} of {y2.numel()})", flush=True)
    ok &= fin and err < 5e-3

x = synth(big)
for passes, label in ((3, "linear_tc2_kernel<3> (split-fp16, three passes)"), (1, "linear_tc2_kernel<1> (fp16, one pass)"),
                      (2, "linear_tf32_kernel  (TF32, one pass, TMA)")):
    y = torch.empty((big, h1), dtype=torch.float32, device=DEV)
    for _ in range(3):
        _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), big, y.data_ptr(), passes, 1, _cabi.stream_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _cabi.check(lib.rqb200_debug_linear_tc(m._handle, 0, 0, x.data_ptr(), big, y.data_ptr(), passes, 1, _cabi.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label}: {ms:.3f} ms for {big} x {in_dim} -> {h1}  ({big * in_dim * 4 / ms / 1e6:.0f} GB/s of X, fp32 rows out)", flush=True)

# the whole fast route: screening off / fp16 screen / TF32 screen at several bounds, codes against the exact route
m.encode_mode = _cabi.ENCODE_EXACT
exact = m.get_indices(x, use_sk=False)
m.encode_mode = _cabi.ENCODE_FAST
tiers = (ctypes.c_int64 * 2)()


def fast_codes(label):
    for _ in range(2):
        c = m.get_indices(x, use_sk=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c = m.get_indices(x, use_sk=False)
    e1.record()
    torch.cuda.synchronize()
    lib.rqb200_model_last_tier_rows(m._handle, tiers)
    bad = int((c != exact).any(1).sum())
    print(f"{label}: {e0.elapsed_time(e1) / 5:.3f} ms per get_indices, rows to the three-pass tier {tiers[0]} "
          f"({100.0 * tiers[0] / big:.2f} %), to the exact tier {tiers[1]} ({100.0 * tiers[1] / big:.2f} %), rows differing from "
          f"the exact route: {bad}", flush=True)
    return bad


m.set_screen(0)
ok &= fast_codes("no screening (three-pass on every row)") == 0
for ge in (9, 10, 11, 12):
    m.set_screen("tf32", 2.0 ** -ge)
    bad = fast_codes(f"TF32 screen, gamma1 = 2^-{ge}")
    if ge <= 10:
        ok &= bad == 0
m.set_screen(1, 2.0 ** -11)
fast_codes("fp16 one-pass screen, gamma1 = 2^-11")
m.set_screen(0)
print("check_tf32", "OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)

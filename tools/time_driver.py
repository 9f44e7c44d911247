"""Times the full encode driver (generate_codes: pass 1, Sinkhorn re-encode rounds, suffix) and its pieces on one GPU.

    python tools/time_driver.py [items] [config: c2_slice|c3_slice|c5_slice]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi      # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden                        # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    name = sys.argv[2] if len(sys.argv) > 2 else "c2_slice"
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    lib = _cabi.lib()
    x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device="cuda:0")
    _cabi.check(lib.rqb200_synth_items(2024, 0, n, cfg["in_dim"], n, _cabi.ptr(x), _cabi.stream_ptr(x.device)))
    for fast in (True, False):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out, stats = rq.generate_codes(build_model(cfg, cbs), x, fast=fast)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"{name} n={n} pass1={'tensor-core' if fast else 'exact'}: {dt * 1e3:.1f} ms total, {stats}", flush=True)
    # one round in isolation, stage by stage (CUDA events of the library)
    import ctypes
    codes = rq.generate_code.encode_codes_exact(m, x)
    for vq in m.rq.vq_layers[:-1]:
        vq.sk_epsilon = 0.0
    items, offsets, max_group = rq.collision_groups(m, codes)
    sizes = (offsets[1:] - offsets[:-1])
    print(f"groups {offsets.numel() - 1} items {items.numel()} max {max_group} mean {float(sizes.float().mean()):.1f}")
    for rep in range(3):
        c = codes.clone()
        lib.rqb200_profile_enable(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        found, done = rq.generate_code.reencode_round(m, c, x)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ms = (ctypes.c_double * 12)()
        cnt = (ctypes.c_longlong * 12)()
        lib.rqb200_profile_read(ms, cnt, 12)
        lib.rqb200_profile_enable(0)
        print(f"one re-encode round over {done} groups: {dt * 1e3:.2f} ms (group extraction {ms[3]:.2f}, small-batch re-encode {ms[9]:.2f}, "
              f"Sinkhorn {ms[5]:.2f})")


if __name__ == "__main__":
    main()

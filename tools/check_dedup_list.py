"""Sort-free suffix dedup (csrc/dedup_list.cu, experimental) against the sort path (diagnostic, GPU box).

Debug flag 8192 (rqb200_debug_tc_flags) routes rqb200_suffix_dedup through the hash table + item-list kernels; the ids,
the distinct count and the largest group must be identical to the sort path on every shape, including the shapes that
make it decline (a group above its walk cap falls back to the sort path inside the call).  Then both are timed.

Run under a timeout:  timeout 300 python tools/check_dedup_list.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200 import _cabi                       # noqa: E402
from ai_education_generative_recommendation_b200.generate_code import suffix_dedup  # noqa: E402

DEV = "cuda:0"
lib = _cabi.lib()
ok = True


def both(codes, Ks):
    lib.rqb200_debug_tc_flags(0)
    a, sa = suffix_dedup(None, codes, Ks)
    lib.rqb200_debug_tc_flags(8192)
    b, sb = suffix_dedup(None, codes, Ks)
    lib.rqb200_debug_tc_flags(0)
    torch.cuda.synchronize()
    return torch.equal(a, b), sa, sb


g = torch.Generator(device=DEV).manual_seed(2024)
cases = [  # n, L, K, distinct-ish cap (None = uniform codes)
    (1, 3, 256, None), (2, 3, 8, None), (707, 3, 8, None), (100_000, 3, 256, None), (1_000_000, 3, 256, None),
    (1_000_000, 4, 256, None), (1_000_000, 4, 1024, None), (300_000, 5, 4096, None),
    (200_000, 3, 4, None),          # 64 codes only: groups of ~3000 members, above the walk cap → in-call fallback
    (50_000, 2, 1, None),           # every item in one group
    (1_000_000, 3, 256, 400),       # heavy collisions: 400 distinct codes, groups ~2500
    (1_000_000, 3, 256, 20_000),    # groups ~50
]
for n, L, K, cap in cases:
    if cap is None:
        codes = torch.randint(0, K, (n, L), generator=g, device=DEV, dtype=torch.int64)
    else:
        pool = torch.randint(0, K, (cap, L), generator=g, device=DEV, dtype=torch.int64)
        codes = pool[torch.randint(0, cap, (n,), generator=g, device=DEV)]
    same, sa, sb = both(codes, [K] * L)
    stats_same = sa == sb
    print(f"n={n} L={L} K={K} pool={cap}: ids identical={same}, stats identical={stats_same} {sa}", flush=True)
    ok &= same and stats_same

codes = torch.randint(0, 256, (1_000_000, 3), generator=g, device=DEV, dtype=torch.int64)
for flags, label in ((0, "sort path"), (8192, "list path")):
    lib.rqb200_debug_tc_flags(flags)
    for _ in range(3):
        suffix_dedup(None, codes, [256] * 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        suffix_dedup(None, codes, [256] * 3)
    e1.record()
    torch.cuda.synchronize()
    lib.rqb200_debug_tc_flags(0)
    print(f"{label}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (1M x 3 codes of 256, incl. the statistics read-back)", flush=True)
sys.exit(0 if ok else 1)

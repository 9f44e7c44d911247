"""One re-encode round of the collision loop (infer.py:116-129) over the 1 M-item C2 catalogue — a short program to put under ncu:
   ncu --set full -k regex:"sinkhorn_regroup|linear_small|quantize_small" … python tools/run_round.py [c2_slice] [items] [rounds]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi      # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden      # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g, cfg, cbs = load_golden(name)
m = build_model(cfg, cbs)
x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device="cuda:0")
_cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, cfg["in_dim"], n, _cabi.ptr(x), _cabi.stream_ptr(x.device)))
codes = rq.generate_code.encode_codes_exact(m, x)
for vq in m.rq.vq_layers[:-1]:
    vq.sk_epsilon = 0.0
import ctypes
lib = _cabi.lib()
use_memo = os.environ.get("RQB200_NO_MEMO", "0") != "1"
memo = rq.generate_code._ReencodeMemo(m, n, x.device) if use_memo else None
lib.rqb200_debug_sinkhorn_variant(int(os.environ.get("RQB200_SK_VARIANT", "0")))
for r in range(rounds):
    lib.rqb200_profile_enable(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    found, done = rq.generate_code.reencode_round(m, codes, x, memo=memo)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ms = (ctypes.c_double * 12)()
    cnt = (ctypes.c_longlong * 12)()
    lib.rqb200_profile_read(ms, cnt, 12)
    lib.rqb200_profile_enable(0)
    items, offsets, mg = rq.collision_groups(m, codes)
    sz = (offsets[1:] - offsets[:-1])
    print(f"round {r}: {found} groups found, {done} re-encoded, {1e3 * dt:.1f} ms wall (re-encode {ms[9]:.1f}, Sinkhorn {ms[5]:.1f}, exact "
          f"linear {ms[0] + ms[1]:.1f}, exact quantizer {ms[2]:.1f}, sort {ms[3]:.1f}); after it: {offsets.numel() - 1} groups, "
          f"{items.numel()} members, largest {mg}, groups > 8 rows {int((sz > 8).sum())}, > 93 rows {int((sz > 93).sum())}", flush=True)

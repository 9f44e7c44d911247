"""How often does a collision group (as a member SET) come back in a later round of the re-encode loop (infer.py:116-129)?
The re-encode of a group is a pure function of its member rows, so a group seen before could be answered from a memo instead of
being recomputed.  Prints, per round, the share of groups (and of member rows) whose member set was already re-encoded in an
earlier round, split by "the round before last" (period-2 cycles) and "any earlier round".
   python tools/group_recurrence.py [c2_slice] [items] [rounds]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi      # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden      # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 30
g, cfg, cbs = load_golden(name)
m = build_model(cfg, cbs)
x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device="cuda:0")
_cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, cfg["in_dim"], n, _cabi.ptr(x), _cabi.stream_ptr(x.device)))
codes = rq.generate_code.encode_codes_exact(m, x)
for vq in m.rq.vq_layers[:-1]:
    vq.sk_epsilon = 0.0


def mix(v, c):
    v = (v + c) * -7046029254386353131          # 0x9E3779B97F4A7C15 as int64
    v = v ^ (v >> 29)
    v = v * -4658895280553007687                # 0xBF58476D1CE4E5B9
    return v ^ (v >> 32)


seen_round = {}
for r in range(rounds):
    items, offsets, mg = rq.collision_groups(m, codes)
    ng = offsets.numel() - 1
    if ng <= 0:
        print(f"round {r}: no groups")
        break
    sizes = offsets[1:] - offsets[:-1]
    h1 = torch.cumsum(mix(items, 1), 0)
    h2 = torch.cumsum(mix(items, 0x1234567), 0)
    z = torch.zeros(1, dtype=torch.int64, device=items.device)
    c1 = torch.cat([z, h1])
    c2 = torch.cat([z, h2])
    k1 = (c1[offsets[1:]] - c1[offsets[:-1]]).cpu().tolist()
    k2 = (c2[offsets[1:]] - c2[offsets[:-1]]).cpu().tolist()
    sz = sizes.cpu().tolist()
    hit_prev2 = hit_any = rows_any = 0
    for a, b, s in zip(k1, k2, sz):
        key = (a, b, s)
        last = seen_round.get(key)
        if last is not None:
            hit_any += 1
            rows_any += s
            if last == r - 2:
                hit_prev2 += 1
        seen_round[key] = r
    print(f"round {r}: {ng} groups, {items.numel()} member rows; seen before {hit_any} ({100.0 * hit_any / ng:.1f} % of groups, "
          f"{100.0 * rows_any / max(1, items.numel()):.1f} % of rows), of which re-encoded two rounds ago {hit_prev2}", flush=True)
    rq.generate_code.reencode_round(m, codes, x)

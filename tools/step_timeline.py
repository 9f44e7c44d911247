"""Kernel timeline of one encode step (get_indices on the fast route + suffix dedup) from CUPTI via torch.profiler:
start offset, duration and the idle gap before every kernel — where the step time goes beyond the kernels themselves.

    python tools/step_timeline.py [config: c2_slice|c3_slice|c5_slice] [items]
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi      # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden                        # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    m.encode_mode = _cabi.ENCODE_FAST
    x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device="cuda:0")
    _cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, cfg["in_dim"], n, _cabi.ptr(x), _cabi.stream_ptr(x.device)))

    def step():
        codes = m.get_indices(x, use_sk=False)
        return rq.suffix_dedup(m, codes)[0]

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # last step = kernels after the last first-layer launch over all rows (the TF32 screening kernel, else the three-pass kernel)
    first = "linear_tf32" if any("linear_tf32" in e.name for e in evs) else "linear_tc2"
    starts = [i for i, e in enumerate(evs) if first in e.name]
    evs = evs[starts[-1]:]
    t0 = evs[0].time_range.start
    prev_end = t0
    total_k = 0.0
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = e.time_range.start - prev_end
        total_k += d
        print(f"{s:9.1f} us  dur {d:8.1f}  gap {gap:7.1f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)
    print(f"step span {prev_end - t0:.1f} us, kernel time {total_k:.1f} us, idle {prev_end - t0 - total_k:.1f} us")


if __name__ == "__main__":
    main()

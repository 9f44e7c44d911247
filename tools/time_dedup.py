"""Suffix dedup alone on the codes of the bench catalogue: CUDA-event time per call and the kernel timeline of one call.

    python tools/time_dedup.py [config: c2_slice|c3_slice|c5_slice] [items]
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi      # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden      # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    g, cfg, cbs = load_golden(name)
    m = build_model(cfg, cbs)
    m.encode_mode = _cabi.ENCODE_FAST
    x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device="cuda:0")
    _cabi.check(_cabi.lib().rqb200_synth_items(2024, 0, n, cfg["in_dim"], n, _cabi.ptr(x), _cabi.stream_ptr(x.device)))
    codes = m.get_indices(x, use_sk=False)
    del x
    for _ in range(5):
        out, stats = rq.suffix_dedup(m, codes)
    torch.cuda.synchronize()
    reps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rq.suffix_dedup(m, codes)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} {n} items: {1e3 * e0.elapsed_time(e1) / reps:.1f} us per suffix_dedup call (incl. the statistics read-back); "
          f"distinct {stats['distinct']}, longest run {stats['max_conflicts']}")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        rq.suffix_dedup(m, codes)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    for e in evs:
        print(f"{e.time_range.start - t0:9.1f} us  dur {e.time_range.end - e.time_range.start:8.1f}  {e.name[:100]}")


def trace():
    """Variant builds with -DRQB_OS_TRACE: phase timestamps of every tile of the three digit passes of the last call."""
    import ctypes
    import numpy as np
    lib = _cabi.lib()
    if not hasattr(lib, "rqb200_debug_os_trace"):
        return
    buf = np.zeros((3, 1024, 8), dtype=np.uint64)
    lib.rqb200_debug_os_trace.argtypes = [ctypes.c_void_p]
    lib.rqb200_debug_os_trace(buf.ctypes.data)
    for p in range(3):
        t = buf[p].astype(np.int64)
        live = t[:, 0] > 0
        t = t[live]
        t0 = t[:, 0].min()
        names = ["start", "counted", "ranked-offsets", "staged", "looked-back", "stored"]
        print(f"pass {p}: {len(t)} tiles; first start 0, last start {(t[:, 0].max() - t0) / 1e3:.1f} us, last end {(t[:, 5].max() - t0) / 1e3:.1f} us")
        for j in range(1, 6):
            d = (t[:, j] - t[:, j - 1]) / 1e3
            print(f"   {names[j - 1]:>14} -> {names[j]:<14} mean {d.mean():6.2f} us  max {d.max():6.2f}  (tile of max {int(d.argmax())})")
        for tile in (0, 1, len(t) // 2, len(t) - 1):
            print(f"   tile {tile}: " + " ".join(f"{(t[tile, j] - t0) / 1e3:6.1f}" for j in range(6)))


if __name__ == "__main__":
    main()
    if os.environ.get("RQB200_OS_TRACE"):
        trace()

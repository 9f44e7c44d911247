"""Clock stamps of one steady-state batch of quantize_tc_kernel (CTA 0, its second batch), from a -DRQB_QTC_TRACE build:
   tools/build_variant.sh qt quantize_tc.cu -DRQB_QTC_TRACE;  RQB200_LIB=.../librqvae_b200_qt.so python tools/trace_qtc.py [c2_slice] [items]
Row-group events: tile start, then per level [A tile ready, per chunk (accumulator full, scan done), gate done, and below the last
level: code row loads issued, residual updated]; MMA lane: per chunk
[codebook chunk in shared memory, per group: ready to issue]."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_education_generative_recommendation_b200 import _cabi      # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden      # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_slice"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g, cfg, cbs = load_golden(name)
m = build_model(cfg, cbs)
m.encode_mode = _cabi.ENCODE_FAST
lib = _cabi.lib()
x = torch.empty((n, cfg["in_dim"]), dtype=torch.float32, device="cuda:0")
_cabi.check(lib.rqb200_synth_items(2024, 0, n, cfg["in_dim"], n, _cabi.ptr(x), _cabi.stream_ptr(x.device)))
for _ in range(3):
    m.get_indices(x, use_sk=False)
torch.cuda.synchronize()
lib.rqb200_debug_qtc_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
lib.rqb200_debug_qtc_trace(None, None, 1)
m.get_indices(x, use_sk=False)          # the first quantizer launch (tier 1, every row) fills the buffers; later launches find them full?
torch.cuda.synchronize()
buf = np.zeros((5, 512), dtype=np.int64)
cnt = np.zeros(5, dtype=np.int32)
lib.rqb200_debug_qtc_trace(buf.ctypes.data, cnt.ctypes.data, 0)
L, K = len(cbs), cbs[0].shape[0]
chunks = (K + 127) // 128
per_tile = 1 + L * (2 + 2 * chunks) + 2 * (L - 1)
t0 = min(buf[w, 0] for w in range(5) if cnt[w])
print(f"{name}: L={L} K={K} chunks/level={chunks}; events per group tile {per_tile}; counts {cnt.tolist()} (first launch = tier 1)")
for w in range(4):
    ev = buf[w, :per_tile] - t0
    print(f"group {w}: tile start {ev[0]}")
    i = 1
    for l in range(L):
        line = f"   level {l}: A ready {ev[i]:7d} |"
        i += 1
        for c in range(chunks):
            line += f" full {ev[i]:7d} scanned {ev[i + 1]:7d} (+{ev[i + 1] - ev[i]}) |"
            i += 2
        line += f" gate done {ev[i]:7d}"
        i += 1
        if l + 1 < L:
            line += f" | row loads issued {ev[i]:7d} | residual updated {ev[i + 1]:7d}"
            i += 2
        print(line)
mm = buf[4, :L * chunks * 5] - t0
print("MMA lane (per chunk: codebook ready, then ready-to-issue for groups 0..3):")
for j in range(L * chunks):
    print("   ", mm[5 * j:5 * j + 5].tolist())

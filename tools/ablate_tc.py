"""Which resource bounds the tensor-core linear kernel?  Times the 768->256 layer with parts switched off (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ai_education_generative_recommendation_b200 import _cabi, RQVAE
lib = _cabi.lib()
n = 1_000_000
m = RQVAE(in_dim=768, num_emb_list=[256], e_dim=32, layers=[256, 128], sk_epsilons=[0.0]).to("cuda:0").eval()   # all 3 layers; 768->256 dominates
x = torch.randn((n, 768), device="cuda:0")
def run(flags, reps=5):
    lib.rqb200_debug_tc_flags(flags)
    m.encode_tc(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): m.encode_tc(x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
names = {0: "production", 1: "no epilogue stores", 2: "no MMA", 4: "no producer STS", 8: "no W bulk loads", 16: "no X loads",
         1 | 8: "no epi stores + no W", 2 | 8: "no MMA + no W", 1 | 2 | 8: "no epi/MMA/W (X stream + convert only)",
         1 | 2 | 4 | 8: "X loads only", 1 | 4 | 16: "MMA + W only", 1 | 4 | 8 | 16: "MMA only", 2 | 4 | 16 | 1: "W loads only"}
for f, nm in names.items():
    print(f"{f:3d} {nm:42s} {run(f):7.3f} ms")
lib.rqb200_debug_tc_flags(0)

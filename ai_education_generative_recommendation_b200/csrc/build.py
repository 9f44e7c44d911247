"""Builds librqvae_b200.so (the C-ABI library) with nvcc for sm_100a, in-tree.

Usage: python build.py [--force]    (also called by __graft_entry__.build())
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "librqvae_b200.so")
SOURCES = ["api.cu", "linear_exact.cu", "quantize.cu", "dedup.cu", "sinkhorn.cu", "kmeans.cu", "synth.cu",
           "encode_tc.cu", "encode_tc2.cu", "quantize_tc.cu", "train.cu", "tokens.cu", "small_batch.cu", "encode_tf32.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    deps = [src, os.path.join(HERE, "common.cuh"), os.path.join(PKG, "..", "include", "rqvae_b200.h"),
            os.path.abspath(__file__)]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, rebuilt = [], False
    procs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, s):
            cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    logs = []
    for src, p in procs:
        out, _ = p.communicate()
        logs.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(objdir, "ptxas.log"), "a" if not force else "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    if rebuilt or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-lcudart", "-lcuda"]
        subprocess.check_call(cmd)
    return OUT


TORCH_OPS_OUT = os.path.join(PKG, "librqvae_b200_torch.so")


def build_torch_ops(force=False):
    """g++ → librqvae_b200_torch.so: the TORCH_LIBRARY registration of the ops (csrc/torch_ops.cpp), a thin layer over
    the C ABI linked against librqvae_b200.so ($ORIGIN) and libtorch.  No kernels in it."""
    src = os.path.join(HERE, "torch_ops.cpp")
    hdr = os.path.join(PKG, "..", "include", "rqvae_b200.h")
    if (not force and os.path.exists(TORCH_OPS_OUT)
            and os.path.getmtime(TORCH_OPS_OUT) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(__file__))):
        return TORCH_OPS_OUT
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", src, "-o", TORCH_OPS_OUT,
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in ce.include_paths("cuda")]
    cmd += [f"-L{tlib}", f"-L{PKG}", "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-lrqvae_b200", "-lcudart", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}", "-Wl,--no-as-needed"]
    subprocess.check_call(cmd)
    return TORCH_OPS_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_ops(force="--force" in sys.argv))

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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

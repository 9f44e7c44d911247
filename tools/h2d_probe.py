"""H2D copy bandwidth of the box for the e2e leg: default pinned memory vs write-combined pinned memory, one stream vs two
concurrent halves (the e2e number of bench.py is bound by exactly this copy)."""
import ctypes
import time

import torch

rt = ctypes.CDLL("libcudart.so")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
N = 3 * 1024 ** 3
dev = torch.empty(N, dtype=torch.uint8, device="cuda:0")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def alloc(flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), N, flags) == 0
    ctypes.memset(p, 1, N)
    return p


def timed(p, two):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if two:
            h = N // 2
            rt.cudaMemcpyAsync(dev.data_ptr(), p, h, 1, s1.cuda_stream)
            rt.cudaMemcpyAsync(dev.data_ptr() + h, p.value + h, N - h, 1, s2.cuda_stream)
        else:
            rt.cudaMemcpyAsync(dev.data_ptr(), p, N, 1, s1.cuda_stream)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return N / best / 1e9


for name, flags in (("pinned (default)", 0), ("pinned, portable", 1), ("pinned, write-combined", 4)):
    p = alloc(flags)
    print(f"{name}: one stream {timed(p, False):.1f} GB/s, two streams {timed(p, True):.1f} GB/s", flush=True)
    rt.cudaFreeHost(p)

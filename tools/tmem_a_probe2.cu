// tmem_a_probe2.cu — the CTA-pair (cta_group::2) version of tmem_a_probe.cu: does a 2-CTA tcgen05.mma take each CTA's half of
// the A operand from that CTA's own tensor memory, at the same address, the way linear_tc3_kernel (csrc/encode_tc3.cu)
// assumes?
//
//   D[256 x 128] (fp32) = A[256 x 64] (fp16) * B[128 x 64]^T (fp16);  CTA r of the pair owns A rows 128 r .. 128 r + 127 (its
//   tensor-memory lanes / its shared memory) and B rows (output features) 64 r .. 64 r + 63 (its shared memory), reads back
//   its own 128 rows of D.
//   mode 0: TS — A written to tensor memory with tcgen05.st.16x256b (the fragments linear_tc3_kernel writes)
//   mode 1: SS — A staged in shared memory (what linear_tc2_kernel does; validates this harness itself)
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ai_education_generative_recommendation_b200/csrc \
//              tools/tmem_a_probe2.cu -o tools/tmem_a_probe2
// run (GPU box):  timeout 60 tools/tmem_a_probe2
// STATUS: compiles; NOT yet run on a B200 (tools/r2_first_call.sh runs it).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_pair.cuh"

using namespace rqb;

constexpr int M = 256, N = 128, K = 64, MH = M / 2, NH = N / 2;

__device__ __forceinline__ void umma_f16_2cta_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ int sw128_off(int r, int c8) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe2_kernel(const __half *__restrict__ A, const __half *__restrict__ B, float *__restrict__ out, int mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sB = smem;                 // this CTA's 64 rows of B: 8 KB
    unsigned char *sA = smem + 8192;          // this CTA's 128 rows of A: 16 KB (SS form)
    uint64_t *done = reinterpret_cast<uint64_t *>(smem + 8192 + 16384);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const __half *Ar = A + (size_t)rank * MH * K;
    const __half *Br = B + (size_t)rank * NH * K;
    if (tid == 0) { mbar_init(done, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    for (int i = tid; i < NH * (K / 8); i += 128) {
        const int r = i / (K / 8), c8 = i % (K / 8);
        *reinterpret_cast<uint4 *>(sB + sw128_off(r, c8)) = *reinterpret_cast<const uint4 *>(Br + r * K + c8 * 8);
    }
    for (int i = tid; i < MH * (K / 8); i += 128) {
        const int r = i / (K / 8), c8 = i % (K / 8);
        *reinterpret_cast<uint4 *>(sA + sw128_off(r, c8)) = *reinterpret_cast<const uint4 *>(Ar + r * K + c8 * 8);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d_tmem = tmem_base;                       // columns 0..127: accumulator
    const uint32_t a_tmem = tmem_base + 128;                 // columns 128..159: A (64 fp16 per row)
    if (mode == 0) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            const int r0 = warp * 32 + blk * 16 + (lane >> 2);
            uint32_t v[16];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int k = 16 * b + 4 * (lane & 3);
                const uint32_t *row_a = reinterpret_cast<const uint32_t *>(Ar + r0 * K + k);
                const uint32_t *row_b = reinterpret_cast<const uint32_t *>(Ar + (r0 + 8) * K + k);
                v[4 * b + 0] = row_a[0]; v[4 * b + 1] = row_a[1];
                v[4 * b + 2] = row_b[0]; v[4 * b + 3] = row_b[1];
            }
            tmem_st_16x256b_x4(a_tmem + ((uint32_t)(warp * 32 + blk * 16) << 16), v);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // both CTAs' operands and barriers are in place
    tc_fence_after();
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = umma_idesc(M, N);
#pragma unroll
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint64_t db = umma_desc(smem_u32(sB) + kk * 32);
            if (mode == 0) umma_f16_2cta_ts(d_tmem, a_tmem + kk * 8, db, idesc, kk != 0);
            else umma_f16_2cta(d_tmem, umma_desc(smem_u32(sA) + kk * 32), db, idesc, kk != 0);
        }
        umma_commit_2cta(done);                              // multicast: the barrier at this offset in both CTAs
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld32(d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) out[((size_t)rank * MH + tid) * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
    }
}

int main() {
    __half *hA = (__half *)malloc(M * K * sizeof(__half)), *hB = (__half *)malloc(N * K * sizeof(__half));
    float *ref = (float *)malloc(M * N * sizeof(float)), *got = (float *)malloc(M * N * sizeof(float));
    srand(11);
    for (int i = 0; i < M * K; ++i) hA[i] = __float2half((float)(rand() % 17 - 8) * 0.125f);
    for (int i = 0; i < N * K; ++i) hB[i] = __float2half((float)(rand() % 13 - 6) * 0.25f);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += __half2float(hA[m * K + k]) * __half2float(hB[n * K + k]);
            ref[m * N + n] = s;                               // small integers x 2^-5: exact in fp32 in any order
        }
    __half *dA, *dB;
    float *dO;
    cudaMalloc(&dA, M * K * sizeof(__half)); cudaMalloc(&dB, N * K * sizeof(__half)); cudaMalloc(&dO, M * N * sizeof(float));
    cudaMemcpy(dA, hA, M * K * sizeof(__half), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, N * K * sizeof(__half), cudaMemcpyHostToDevice);
    const int smem = 8192 + 16384 + 64 + 1024;
    int rc = 0;
    for (int mode = 1; mode >= 0; --mode) {                   // the known-good SS form first
        cudaMemset(dO, 0xFF, M * N * sizeof(float));
        probe2_kernel<<<2, 128, smem>>>(dA, dB, dO, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(got, dO, M * N * sizeof(float), cudaMemcpyDeviceToHost);
        int bad = 0, bad_lo = 0;
        for (int i = 0; i < M * N; ++i) {
            const int b = !(got[i] == ref[i]);
            bad += b;
            if (i < MH * N) bad_lo += b;
        }
        printf("cta_group::2, %s: %d of %d outputs differ from the host reference (%d in the leader's rows, %d in the peer's)\n",
               mode == 0 ? "A from tensor memory (TS, 16x256b stores)" : "A from shared memory (SS)               ", bad, M * N, bad_lo,
               bad - bad_lo);
        rc |= bad != 0;
    }
    return rc;
}

// tmem_a_probe.cu — round-2 groundwork: does tcgen05.mma take its A operand from tensor memory
// the way linear_tc2_kernel would need it, and what does it cost next to the shared-memory form?
//
//   D[128 x 64] (fp32, TMEM) = A[128 x 64] (fp16) * B[64 x 64]^T (fp16, shared memory, K-major SWIZZLE_128B)
//   TS form: A written to TMEM by its own row's thread with tcgen05.st (lane = row, one 32-bit column = two K elements)
//   TS 16x256b: A written with the 16-lane fragment form (lane t: rows t/4 and t/4+8 of a 16-row block, columns 2(t%4),
//            2(t%4)+1 of every 8-column group) — the mapping linear_tc3_kernel (csrc/encode_tc3.cu) relies on
//   SS form: A staged in shared memory like B (what the kernels do today)
// Both results are compared with a host reference; then each form is issued REPS times back to back for a rate.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ai_education_generative_recommendation_b200/csrc \
//              tools/tmem_a_probe.cu -o tools/tmem_a_probe
// run (GPU box):  timeout 60 tools/tmem_a_probe
// STATUS: compiles (ptxas accepts the TS operand form and both store shapes); NOT yet run on a B200 — first thing to do
// in round 2 (tools/r2_first_call.sh).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

using namespace rqb;

constexpr int M = 128, N = 64, K = 64, REPS = 2000;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 16 lanes x 32 columns: v[4b + 0/1] = (row lane/4, columns 8b + 2(lane%4) + 0/1), v[4b + 2/3] = same columns of row lane/4 + 8
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// row r, 16-byte chunk c8 (8 fp16) of a K-major SWIZZLE_128B tile with 64 fp16 per row
__device__ __forceinline__ int sw128_off(int r, int c8) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4); }

// mode 0: TS (A from TMEM, 32x32b stores), mode 1: SS (A from shared memory), mode 2: TS with 16x256b stores; out[M][N] fp32; cycles[0] = clocks for REPS x 4 MMAs
__global__ void __launch_bounds__(128) probe_kernel(const __half *__restrict__ A, const __half *__restrict__ B, float *__restrict__ out,
                                                    long long *__restrict__ cycles, int mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sB = smem;                 // 64 rows x 128 B = 8 KB
    unsigned char *sA = smem + 8192;          // 128 rows x 128 B = 16 KB
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 8192 + 16384);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // operands → shared memory (B always, A for the SS form)
    for (int i = tid; i < N * (K / 8); i += 128) {
        const int r = i / (K / 8), c8 = i % (K / 8);
        *reinterpret_cast<uint4 *>(sB + sw128_off(r, c8)) = *reinterpret_cast<const uint4 *>(B + r * K + c8 * 8);
    }
    for (int i = tid; i < M * (K / 8); i += 128) {
        const int r = i / (K / 8), c8 = i % (K / 8);
        *reinterpret_cast<uint4 *>(sA + sw128_off(r, c8)) = *reinterpret_cast<const uint4 *>(A + r * K + c8 * 8);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d_tmem = tmem_base;                       // columns 0..63: accumulator
    const uint32_t a_tmem = tmem_base + 64;                  // columns 64..95: A (64 fp16 per row = 32 columns)
    {
        // tensor memory keeps its contents between launches: poison the A block (fp16 NaNs) so that a store that lands in the
        // wrong place cannot pass on the previous mode's data
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0x7E007E00u;
        tmem_st32(a_tmem + ((uint32_t)(warp * 32) << 16), v);
    }
    if (mode == 0) {
        // thread = row = TMEM lane; column j of the A block holds K elements 2j, 2j+1
        uint32_t v[32];
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A + tid * K);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = src[j];
        tmem_st32(a_tmem + ((uint32_t)(warp * 32) << 16), v);
    }
    if (mode == 2) {
        const int lane = tid & 31;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {                  // the warp's two 16-row blocks
            const int r0 = warp * 32 + blk * 16 + (lane >> 2);
            uint32_t v[16];
#pragma unroll
            for (int b = 0; b < 4; ++b) {                    // 8-column group b = K elements 16b .. 16b+15
                const int k = 16 * b + 4 * (lane & 3);
                const uint32_t *lo_row = reinterpret_cast<const uint32_t *>(A + r0 * K + k);
                const uint32_t *hi_row = reinterpret_cast<const uint32_t *>(A + (r0 + 8) * K + k);
                v[4 * b + 0] = lo_row[0]; v[4 * b + 1] = lo_row[1];
                v[4 * b + 2] = hi_row[0]; v[4 * b + 3] = hi_row[1];
            }
            tmem_st_16x256b_x4(a_tmem + ((uint32_t)(warp * 32 + blk * 16) << 16), v);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc(M, N);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
            for (int kk = 0; kk < K / 16; ++kk) {
                const uint64_t db = umma_desc(smem_u32(sB) + kk * 32);
                if (mode != 1) umma_f16_ts(d_tmem, a_tmem + kk * 8, db, idesc, kk != 0);
                else umma_f16(d_tmem, umma_desc(smem_u32(sA) + kk * 32), db, idesc, kk != 0);
            }
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    // read back the accumulator: lane = row, 64 columns
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld32(d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) out[tid * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

// Layout dump: every (thread, register) of a tcgen05.st.16x256b.x4 writes its own tag into a 32-column block, the block is
// read back with the 32x32b load (lane = row, register j = column j).  tags[lane][col] = (thread << 8) | register.
__global__ void __launch_bounds__(128) layout_kernel(uint32_t *__restrict__ tags) {
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_slot;
    {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0xFFFFFFFFu;
        tmem_st32(base + ((uint32_t)(warp * 32) << 16), v);
    }
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
        uint32_t v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = ((uint32_t)lane << 8) | (uint32_t)r;
        tmem_st_16x256b_x4(base + ((uint32_t)(warp * 32 + blk * 16) << 16), v);
    }
    {
        uint32_t v[32];
        tmem_ld32(base + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tags[tid * 32 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u));
}

static int check_layout() {
    uint32_t *d = nullptr, *h = (uint32_t *)malloc(128 * 32 * sizeof(uint32_t));
    cudaMalloc(&d, 128 * 32 * sizeof(uint32_t));
    layout_kernel<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("layout dump: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 128 * 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    // assumed by linear_tc3_kernel: register 4b + 2h + p of thread t → lane 16*blk + t/4 + 8h (within the warp's 32), column 8b + 2(t%4) + p
    int bad = 0;
    for (int row = 0; row < 128; ++row)
        for (int col = 0; col < 32; ++col) {
            const int rl = row & 15, t = (rl & 7) * 4 + ((col & 7) >> 1), reg = 4 * (col >> 3) + 2 * (rl >> 3) + (col & 1);
            bad += h[row * 32 + col] != (((uint32_t)t << 8) | (uint32_t)reg);
        }
    printf("tcgen05.st.16x256b.x4 layout: %d of 4096 cells differ from the mapping linear_tc3_kernel assumes\n", bad);
    if (bad) {
        printf("observed (thread:register) per column, rows 0..17 of warp 0:\n");
        for (int row = 0; row < 18; ++row) {
            printf("row %2d:", row);
            for (int col = 0; col < 32; ++col) printf(" %2u:%-2u", (h[row * 32 + col] >> 8) & 0xFFFFFF, h[row * 32 + col] & 0xFF);
            printf("\n");
        }
    }
    cudaFree(d); free(h);
    return bad != 0;
}

int main() {
    __half *hA = (__half *)malloc(M * K * sizeof(__half)), *hB = (__half *)malloc(N * K * sizeof(__half));
    float *ref = (float *)malloc(M * N * sizeof(float)), *got = (float *)malloc(M * N * sizeof(float));
    srand(7);
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
    long long *dC;
    cudaMalloc(&dA, M * K * sizeof(__half)); cudaMalloc(&dB, N * K * sizeof(__half));
    cudaMalloc(&dO, M * N * sizeof(float)); cudaMalloc(&dC, sizeof(long long));
    cudaMemcpy(dA, hA, M * K * sizeof(__half), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, N * K * sizeof(__half), cudaMemcpyHostToDevice);
    const int smem = 8192 + 16384 + 64 + 1024;
    int rc = 0;
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(dO, 0, M * N * sizeof(float));
        probe_kernel<<<1, 128, smem>>>(dA, dB, dO, dC, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long cyc = 0;
        cudaMemcpy(got, dO, M * N * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(&cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < M * N; ++i) bad += got[i] != ref[i];
        printf("%s: %d of %d outputs differ from the host reference; %lld clocks for %d x %d MMAs (%.1f clk per 128x64x16 MMA)\n",
               mode == 0 ? "A from TMEM  (TS, 32x32b stores) " : mode == 1 ? "A from shared (SS)               " : "A from TMEM  (TS, 16x256b stores)", bad, M * N, cyc, REPS, K / 16, (double)cyc / (REPS * (K / 16)));
        rc |= bad != 0;
    }
    rc |= check_layout();
    return rc;
}

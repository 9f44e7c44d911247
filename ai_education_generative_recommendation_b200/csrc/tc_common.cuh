// tc_common.cuh — inline-PTX helpers shared by the tcgen05 kernels (mbarrier, bulk copy, UMMA descriptors,
// TMEM loads, fp32 → split-fp16 conversion).  sm_100a only.
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace rqb {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart (dense tile)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}

// instruction descriptor: D=F32, A=B=F16, both K-major, M=m, N=n
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// asynchronous 16-column read: the registers are valid only after tmem_ld16_wait(v).  tcgen05.wait::ld waits for EVERY
// outstanding read of the thread, so a software pipeline is "wait(a) → issue(b) → use a → wait(b) → issue(a') → use b".
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// the "+r" operands pin the uses of v[] behind the wait
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i += 8)
        asm volatile("" : "+r"(v[i]), "+r"(v[i + 1]), "+r"(v[i + 2]), "+r"(v[i + 3]), "+r"(v[i + 4]), "+r"(v[i + 5]),
                          "+r"(v[i + 6]), "+r"(v[i + 7]));
}

// fp32 pair → packed fp16 hi and fp16 lo (x ≈ hi + lo, |x - hi - lo| ≲ 2^-22 |x|)
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
}


constexpr int EPI_LD = 36;     // floats per row of the per-warp transpose patch (16-byte aligned, conflict-free)

// Epilogue for one 128 x N accumulator (this warp: TMEM lanes 32*w .. 32*w+31 = rows row0 .. row0+31).
// TMEM gives lane = row; each 32x32 chunk is transposed through a per-warp shared-memory patch so that the global
// stores are 128 contiguous bytes per row (4 rows per STG.128 instruction, 4 L1 wavefronts) instead of 32
// scattered 16-byte pieces (32 wavefronts).  Scale / bias / ReLU are applied after the transpose.
template <int N>
__device__ __forceinline__ void epilogue_rows(uint32_t taddr, float *patch, int lane, const float *__restrict__ bias,
                                              float inv_scale, int relu, float *__restrict__ Y, int64_t row0, int64_t n,
                                              bool store) {
    const int sub = lane >> 3, q = lane & 7;
#pragma unroll 1
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4 *>(patch + lane * EPI_LD + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + c) + q);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + sub;
            const float4 a = *reinterpret_cast<const float4 *>(patch + r * EPI_LD + 4 * q);
            float4 o;
            o.x = fmaf(a.x, inv_scale, b4.x); o.y = fmaf(a.y, inv_scale, b4.y);
            o.z = fmaf(a.z, inv_scale, b4.z); o.w = fmaf(a.w, inv_scale, b4.w);
            if (relu) {
                o.x = (o.x != o.x) ? o.x : fmaxf(o.x, 0.0f); o.y = (o.y != o.y) ? o.y : fmaxf(o.y, 0.0f);
                o.z = (o.z != o.z) ? o.z : fmaxf(o.z, 0.0f); o.w = (o.w != o.w) ? o.w : fmaxf(o.w, 0.0f);
            }
            if (store && row0 + r < n) *reinterpret_cast<float4 *>(Y + (row0 + r) * (int64_t)N + c + 4 * q) = o;
        }
        __syncwarp();
    }
}


// Epilogue variant that hands the activation to the next tensor-core kernel in the form it consumes: per 128-row tile
// and 64-column slab one [hi tile | lo tile] pair (16 KB each) of fp16 in the UMMA K-major SWIZZLE_128B layout, so the
// consumer stages its A operand with ONE bulk copy per slab and needs no converter warps.  Same bytes as fp32.
// Tile (t, s) lives at tiled + ((t * (N/64) + s) * 2) * 16384.  The 32x32 chunk still goes through the per-warp smem
// patch: afterwards lane (row = 8i + lane/4, chunk = lane%4) holds 8 consecutive columns = one 16-byte piece of hi and
// of lo, and the four lanes of a row store 64 contiguous bytes.
template <int N>
__device__ __forceinline__ void epilogue_rows_split(uint32_t taddr, float *patch, int lane, const float *__restrict__ bias,
                                                    float inv_scale, int relu, unsigned char *__restrict__ tiled, int64_t tile,
                                                    int row_in_tile0) {
    const int rsub = lane >> 2, ch = lane & 3;
#pragma unroll 1
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4 *>(patch + lane * EPI_LD + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const float4 ba = __ldg(reinterpret_cast<const float4 *>(bias + c + 8 * ch));
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + c + 8 * ch + 4));
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        unsigned char *hi_t = tiled + ((size_t)tile * (N / 64) + (c >> 6)) * 32768;
        const int chunk = ((c & 63) >> 3) + ch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + rsub;
            const float4 a0 = *reinterpret_cast<const float4 *>(patch + rr * EPI_LD + 8 * ch);
            const float4 a1 = *reinterpret_cast<const float4 *>(patch + rr * EPI_LD + 8 * ch + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float h[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float y = fmaf(av[t], inv_scale, bv[t]);
                h[t] = (relu && !(y != y)) ? fmaxf(y, 0.0f) : y;
            }
            uint4 hi, lo;
            split2(h[0], h[1], hi.x, lo.x); split2(h[2], h[3], hi.y, lo.y);
            split2(h[4], h[5], hi.z, lo.z); split2(h[6], h[7], hi.w, lo.w);
            const int r = row_in_tile0 + rr;
            const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
            *reinterpret_cast<uint4 *>(hi_t + off) = hi;
            *reinterpret_cast<uint4 *>(hi_t + 16384 + off) = lo;
        }
        __syncwarp();
    }
}

}  // namespace
}  // namespace rqb

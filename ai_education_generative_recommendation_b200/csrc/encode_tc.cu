// encode_tc.cu — tensor-core encoder: tcgen05 / TMEM GEMMs with split-fp16 operands, plus the
// margin-gated fast path of get_indices.
//
// Replaces the hot loop of `RQVAE.get_indices` (reference RQ-VAE/models/rqvae.py:67-71) for the
// catalogue pass: the three `nn.Linear` GEMMs of the encoder (reference RQ-VAE/models/layers.py:23,
// 42-43) run on the 5th-gen tensor cores.  fp32 operands are split on the fly into fp16 hi + lo
// (x = hi + lo to ~22 bits) and each K-slab issues three MMAs (hi*hi, hi*lo, lo*hi) into one fp32 TMEM
// accumulator — fp32-class accuracy at 1/3 of the fp16 tensor rate.  The result z~ differs from the
// reference's FMA-chain z in the last bits only, so codes are certified by a margin gate in the
// quantizer: a row whose top-2 distance gap at any level is within the error bound is recomputed by
// the exact SIMT kernels (linear_exact.cu / quantize.cu).  Codes are therefore identical to the exact
// path; only the route differs.
//
// Kernel anatomy (one persistent CTA per SM, 448 threads, static round-robin over 128-row tiles):
//   warps 0-3   epilogue: tcgen05.ld accumulator → *2^-s + bias → ReLU → fp32 rows to HBM
//   warps 4-11  A producers: coalesced LDG of the fp32 slab (prefetched one slab ahead in registers)
//               → fp16 hi/lo → st.shared in the UMMA K-major SWIZZLE_128B layout
//   warp 12     MMA issuer (one lane): tcgen05.mma.cta_group::1.kind::f16, M=128, N=out features
//   warp 13     W loader (one lane): cp.async.bulk of the pre-packed, pre-swizzled W slab (hi|lo)
// mbarrier rings: full_a / full_w / empty per smem stage, tmem_full / tmem_empty per accumulator buffer
// (two buffers, so the epilogue of tile i overlaps the MMAs of tile i+1).
//
// In the shipped fast route this generic kernel only serves layer shapes the specialised ones do not cover:
//   * the wide first layer (out = 256) runs in the 2-CTA kernel of encode_tc2.cu, which hands its output on as
//     split-fp16 UMMA tiles (tc_common.cuh: epilogue_rows_split);
//   * the last two layers run fused in mlp23_tc_kernel below (the 128-wide activation never leaves the SM, the
//     tiles are staged by bulk copies).
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

#ifndef RQB200_TC2_DEFAULT
#define RQB200_TC2_DEFAULT 1      // 2-CTA kernel for 256-wide layers (RQB200_TC2=0 selects the 1-CTA kernel)
#endif

namespace rqb {

namespace {

constexpr int TM = 128;            // rows per tile (UMMA M)
constexpr int BK = 64;             // K elements per slab = one 128-byte swizzle row of fp16
constexpr int TC_THREADS = 448;
constexpr int EPI_WARPS = 4, CONV_WARPS = 8;
constexpr int CONV_THREADS = CONV_WARPS * 32;
constexpr int A_TILE_BYTES = TM * BK * 2;          // 16 KB (one of hi / lo)

// ---- ablation switches for tools/ablate_tc.py (0 in production): bit0 no epilogue stores, bit1 no MMA,
// bit2 no producer smem stores, bit3 no W bulk loads, bit4 no producer global loads
__device__ int g_tc_debug = 0;

template <int N, int PASSES>
struct TcCfg {
    static constexpr int NSPLIT = PASSES == 1 ? 1 : 2;                 // tiles per operand: hi [, lo]
    static constexpr int W_TILE_BYTES = N * BK * 2;                    // one of hi / lo
    static constexpr int STAGE_BYTES = NSPLIT * (A_TILE_BYTES + W_TILE_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 6 ? 6 : ((200 * 1024) / STAGE_BYTES >= 4 ? 4 : (200 * 1024) / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * N < 32) ? 32 : 2 * N;         // two accumulator buffers
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * 32 * EPI_LD * 4;
    static_assert(STAGES >= 2, "need at least two smem stages");
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N must be a multiple of 16 in [16,256]");
};

// Y[n, N] = act((X[n, K] · W^T) * 2^-s + b);  W pre-packed per 64-wide K slab (hi tile | lo tile, swizzled)
// PASSES = 3: split-fp16 operands (hi+lo), three MMAs per K step; PASSES = 1: hi halves only (tier-1 screening pass).
// n_dev (may be NULL): the row count lives on the device (written by the previous tier's gate); n is its upper bound.
template <int N, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const float *__restrict__ X, int64_t n, int K, const unsigned char *__restrict__ Wp,
                 const float *__restrict__ bias, float inv_scale, int relu, float *__restrict__ Y,
                 const unsigned long long *__restrict__ n_dev) {
    using Cfg = TcCfg<N, PASSES>;
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t *full_a = bars;                       // [STAGES]
    uint64_t *full_w = bars + Cfg::STAGES;         // [STAGES]
    uint64_t *empty = bars + 2 * Cfg::STAGES;      // [STAGES]
    uint64_t *tmem_full = bars + 3 * Cfg::STAGES;  // [2]
    uint64_t *tmem_empty = tmem_full + 2;          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dbg = g_tc_debug;
    const int KS = (K + BK - 1) / BK;
    const int64_t ntiles = (n + TM - 1) / TM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_a[s], CONV_WARPS);
            mbar_init(&full_w[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], EPI_WARPS * 32);
        }
        fence_barrier_init();
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < EPI_WARPS) {
        // ===================== epilogue =====================
        float *patch = reinterpret_cast<float *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256) + warp * (32 * EPI_LD);
        int64_t it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int buf = (int)(it & 1);
            mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            epilogue_rows<N>(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * N), patch, lane, bias, inv_scale, relu,
                             Y, tile * TM + warp * 32, n, !(dbg & 1));
            tc_fence_before();
            mbar_arrive(&tmem_empty[buf]);
        }
    } else if (warp < EPI_WARPS + CONV_WARPS) {
        // ===================== A producers =====================
        const int ct = threadIdx.x - EPI_WARPS * 32;        // 0..255
        // unit u = ct + 256*i (i<8): row = u / 16, c4 = u % 16 → one float4 (4 consecutive k) of one row.
        // A half-warp covers 256 contiguous bytes of one row (full 32-byte sectors per LDG.128) and stores
        // 128 contiguous (swizzled) bytes of fp16 — conflict-free STS.64.
        const int c4 = ct & 15;
        const int rbase = ct >> 4;                           // 0..15, rows rbase + 16*i
        auto load_slab = [&](int64_t tile, int slab, float4 (&dst)[8]) {
            const int k0 = slab * BK + c4 * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t row = tile * TM + rbase + 16 * i;
                if (row < n && k0 < K && !(dbg & 16)) dst[i] = __ldg(reinterpret_cast<const float4 *>(X + row * (int64_t)K + k0));
                else dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int stage = 0;
        uint32_t phase = 0;
        auto convert_slab = [&](const float4 (&src)[8]) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char *a_hi = smem + stage * Cfg::STAGE_BYTES;
            unsigned char *a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                const int off = (r >> 3) * 1024 + (r & 7) * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + ((c4 & 1) << 3);
                if (PASSES == 1) {
                    __half2 h0 = __floats2half2_rn(src[i].x, src[i].y), h1 = __floats2half2_rn(src[i].z, src[i].w);
                    if (!(dbg & 4))
                        *reinterpret_cast<uint2 *>(a_hi + off) = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
                } else {
                    uint2 hi, lo;
                    split2(src[i].x, src[i].y, hi.x, lo.x);
                    split2(src[i].z, src[i].w, hi.y, lo.y);
                    if (!(dbg & 4)) {
                        *reinterpret_cast<uint2 *>(a_hi + off) = hi;
                        *reinterpret_cast<uint2 *>(a_lo + off) = lo;
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[stage]);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        };
        // flattened (tile, slab) sequence, two register buffers used alternately: the loads of step i+1 are in
        // flight while step i is converted, and no register copies force an early wait on them.
        const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t steps = my_tiles * KS;
        float4 bufA[8], bufB[8];
        auto coords = [&](int64_t st, int64_t &tile, int &slab) {
            tile = blockIdx.x + (st / KS) * (int64_t)gridDim.x;
            slab = (int)(st % KS);
        };
        int64_t t0; int s0;
        if (steps > 0) { coords(0, t0, s0); load_slab(t0, s0, bufA); }
        for (int64_t st = 0; st < steps; st += 2) {
            if (st + 1 < steps) { coords(st + 1, t0, s0); load_slab(t0, s0, bufB); }
            convert_slab(bufA);
            if (st + 1 < steps) {
                if (st + 2 < steps) { coords(st + 2, t0, s0); load_slab(t0, s0, bufA); }
                convert_slab(bufB);
            }
        }
    } else if (warp == 12) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(TM, N);
            int stage = 0;
            uint32_t phase = 0;
            int64_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int buf = (int)(it & 1);
                mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait(&full_a[stage], phase);
                    mbar_wait(&full_w[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t w_hi = a_hi + Cfg::NSPLIT * A_TILE_BYTES;
                    const uint32_t w_lo = w_hi + Cfg::W_TILE_BYTES;
                    if (!(dbg & 2))
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        const uint32_t ko = kk * 32;       // 16 fp16 = 32 bytes along K inside the swizzle row
                        if (PASSES == 1) {
                            umma_f16(d_tmem, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc, (slab | kk) != 0);
                        } else {
                            // small cross terms first, the dominant hi*hi product last
                            umma_f16(d_tmem, umma_desc(a_lo + ko), umma_desc(w_hi + ko), idesc, (slab | kk) != 0);
                            umma_f16(d_tmem, umma_desc(a_hi + ko), umma_desc(w_lo + ko), idesc, 1);
                            umma_f16(d_tmem, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc, 1);
                        }
                    }
                    umma_commit(&empty[stage]);            // smem stage reusable once these MMAs retire
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[buf]);              // accumulator complete
            }
        }
    } else {
        // ===================== W loader =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            constexpr uint32_t slab_bytes = 2 * Cfg::W_TILE_BYTES;               // packed image: hi tile | lo tile per slab
            constexpr uint32_t copy_bytes = Cfg::NSPLIT * Cfg::W_TILE_BYTES;     // one pass needs the hi tile only
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (dbg & 8) {
                        mbar_arrive(&full_w[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full_w[stage], copy_bytes);
                        bulk_g2s(smem + stage * Cfg::STAGE_BYTES + Cfg::NSPLIT * A_TILE_BYTES, Wp + (size_t)slab * slab_bytes, copy_bytes,
                                 &full_w[stage]);
                    }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS));
    }
}


// ---- layers 2 + 3 fused: H1[n,K2] → relu(H1·W2ᵀ+b2) [128 wide, never leaves the SM] → ·W3ᵀ+b3 → Z[n,N3] --------------
//
// Same pipeline as linear_tc_kernel<128,3> for the second Linear; its epilogue does not store H2 but re-splits it into
// fp16 hi/lo and writes it as the A operand (UMMA K-major SWIZZLE_128B, two 64-wide slabs) of the third Linear, whose
// pre-packed weights stay resident in shared memory.  The third layer of tile i is issued behind the second layer of
// tile i+1 (software pipelining), so the tensor pipe never waits for the conversion.  Saves the H2 round trip through
// HBM (1 GB per million rows) and a launch.
template <int N3>
struct Tc23Cfg {
    static constexpr int N2 = 128;
    static constexpr int W2_TILE = N2 * BK * 2;                       // 16 KB (hi or lo)
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * W2_TILE;   // 64 KB
    static constexpr int STAGES = 2;
    static constexpr int H2_BYTES = 2 * 2 * A_TILE_BYTES;             // two K slabs x (hi | lo) = 64 KB
    static constexpr int W3_TILE = N3 * BK * 2;                        // hi or lo of one slab
    static constexpr int W3_BYTES = 2 * 2 * W3_TILE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + H2_BYTES + W3_BYTES + 1024 + 256;
    static constexpr int ACC3_COL = 2 * N2;                            // TMEM: two 128-column buffers, then the third layer's accumulator
};

// TILED_A: X is the previous layer's activation as split-fp16 UMMA tiles (epilogue_rows_split): the loader lane stages a
// whole A slab (hi | lo, 32 KB) with one bulk copy next to the W2 slab and the converter warps have nothing to do.
template <int N3, bool TILED_A>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp23_tc_kernel(const float *__restrict__ X, int64_t n, int K, const unsigned char *__restrict__ Wp2,
                const float *__restrict__ bias2, float inv_scale2, const unsigned char *__restrict__ Wp3,
                const float *__restrict__ bias3, float inv_scale3, float *__restrict__ Z,
                const unsigned long long *__restrict__ n_dev) {
    using Cfg = Tc23Cfg<N3>;
    constexpr int N2 = Cfg::N2;
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *h2 = smem + Cfg::STAGES * Cfg::STAGE_BYTES;        // slab s: hi at s*32 KB, lo at s*32 KB + 16 KB
    unsigned char *w3 = h2 + Cfg::H2_BYTES;                            // slab s: hi at s*2*W3_TILE, lo behind it
    uint64_t *bars = reinterpret_cast<uint64_t *>(w3 + Cfg::W3_BYTES);
    uint64_t *full_a = bars;                       // [STAGES]
    uint64_t *full_w = bars + Cfg::STAGES;         // [STAGES]
    uint64_t *empty = bars + 2 * Cfg::STAGES;      // [STAGES]
    uint64_t *tmem_full = bars + 3 * Cfg::STAGES;  // [2]
    uint64_t *tmem_empty = tmem_full + 2;          // [2]
    uint64_t *h2_full = tmem_empty + 2;            // epilogue → MMA: the H2 tile of a row tile is in place (count 4)
    uint64_t *z_full = h2_full + 1;                // MMA → epilogue: third-layer accumulator complete (commit)
    uint64_t *w3_full = z_full + 1;                // loader → MMA: W3 resident
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w3_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KS = (K + BK - 1) / BK;
    const int64_t ntiles = (n + TM - 1) / TM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_a[s], TILED_A ? 1 : CONV_WARPS); mbar_init(&full_w[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS * 32); }
        mbar_init(h2_full, EPI_WARPS);
        mbar_init(z_full, 1);
        mbar_init(w3_full, 1);
        fence_barrier_init();
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < EPI_WARPS) {
        // ===================== epilogue: thread = row (TMEM lane) =====================
        const int rloc = warp * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        auto store_z = [&](int64_t tile) {         // third-layer accumulator → *2^-s + bias → Z rows
            tc_fence_after();
            const int64_t row = tile * TM + rloc;
#pragma unroll
            for (int c = 0; c < N3; c += 32) {
                uint32_t v[32];
                tmem_ld32(t_lane + (uint32_t)(Cfg::ACC3_COL + c), v);
                if (row < n) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias3 + c) + j);
                        float4 o;
                        o.x = fmaf(__uint_as_float(v[4 * j]), inv_scale3, b4.x); o.y = fmaf(__uint_as_float(v[4 * j + 1]), inv_scale3, b4.y);
                        o.z = fmaf(__uint_as_float(v[4 * j + 2]), inv_scale3, b4.z); o.w = fmaf(__uint_as_float(v[4 * j + 3]), inv_scale3, b4.w);
                        *reinterpret_cast<float4 *>(Z + row * (int64_t)N3 + c + 4 * j) = o;
                    }
                }
            }
            tc_fence_before();
        };
        int64_t it = 0, prev_tile = -1;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int buf = (int)(it & 1);
            mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
            if (it > 0) {                           // the third layer of the previous tile: done ⇒ its H2 reads are done too
                mbar_wait(z_full, (uint32_t)((it - 1) & 1));
                store_z(prev_tile);
            }
            tc_fence_after();
            // second-layer accumulator → *2^-s + bias → ReLU → fp16 hi/lo → the third layer's A operand
#pragma unroll 1
            for (int c = 0; c < N2; c += 32) {
                uint32_t v[32];
                tmem_ld32(t_lane + (uint32_t)(buf * N2 + c), v);
                unsigned char *hi_t = h2 + (c >> 6) * (2 * A_TILE_BYTES);
                unsigned char *lo_t = hi_t + A_TILE_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) {       // 8 columns = one 16-byte chunk of the swizzled row
                    const float4 ba = __ldg(reinterpret_cast<const float4 *>(bias2 + c) + 2 * j);
                    const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias2 + c) + 2 * j + 1);
                    const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                    float h[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float y = fmaf(__uint_as_float(v[8 * j + t]), inv_scale2, bv[t]);
                        h[t] = (y != y) ? y : fmaxf(y, 0.0f);
                    }
                    uint4 hi, lo;
                    split2(h[0], h[1], hi.x, lo.x); split2(h[2], h[3], hi.y, lo.y);
                    split2(h[4], h[5], hi.z, lo.z); split2(h[6], h[7], hi.w, lo.w);
                    const int chunk = ((c & 63) >> 3) + j;
                    const int off = (rloc >> 3) * 1024 + (rloc & 7) * 128 + ((chunk ^ (rloc & 7)) << 4);
                    *reinterpret_cast<uint4 *>(hi_t + off) = hi;
                    *reinterpret_cast<uint4 *>(lo_t + off) = lo;
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[buf]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(h2_full);
            prev_tile = tile;
        }
        if (it > 0) {
            mbar_wait(z_full, (uint32_t)((it - 1) & 1));
            store_z(prev_tile);
        }
    } else if (warp < EPI_WARPS + CONV_WARPS) {
        // ===================== A producers (identical to linear_tc_kernel; idle when the input arrives as tiles) =====================
        if (TILED_A) goto teardown;
        const int ct = threadIdx.x - EPI_WARPS * 32;
        const int c4 = ct & 15;
        const int rbase = ct >> 4;
        auto load_slab = [&](int64_t tile, int slab, float4 (&dst)[8]) {
            const int k0 = slab * BK + c4 * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t row = tile * TM + rbase + 16 * i;
                if (row < n && k0 < K) dst[i] = __ldg(reinterpret_cast<const float4 *>(X + row * (int64_t)K + k0));
                else dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int stage = 0;
        uint32_t phase = 0;
        auto convert_slab = [&](const float4 (&src)[8]) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char *a_hi = smem + stage * Cfg::STAGE_BYTES;
            unsigned char *a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                const int off = (r >> 3) * 1024 + (r & 7) * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + ((c4 & 1) << 3);
                uint2 hi, lo;
                split2(src[i].x, src[i].y, hi.x, lo.x);
                split2(src[i].z, src[i].w, hi.y, lo.y);
                *reinterpret_cast<uint2 *>(a_hi + off) = hi;
                *reinterpret_cast<uint2 *>(a_lo + off) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[stage]);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        };
        const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t steps = my_tiles * KS;
        float4 bufA[8], bufB[8];
        auto coords = [&](int64_t st, int64_t &tile, int &slab) {
            tile = blockIdx.x + (st / KS) * (int64_t)gridDim.x;
            slab = (int)(st % KS);
        };
        int64_t t0; int s0;
        if (steps > 0) { coords(0, t0, s0); load_slab(t0, s0, bufA); }
        for (int64_t st = 0; st < steps; st += 2) {
            if (st + 1 < steps) { coords(st + 1, t0, s0); load_slab(t0, s0, bufB); }
            convert_slab(bufA);
            if (st + 1 < steps) {
                if (st + 2 < steps) { coords(st + 2, t0, s0); load_slab(t0, s0, bufA); }
                convert_slab(bufB);
            }
        }
    } else if (warp == 12) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc2 = umma_idesc(TM, N2);
            constexpr uint32_t idesc3 = umma_idesc(TM, N3);
            int stage = 0;
            uint32_t phase = 0;
            int64_t it = 0;
            auto issue_third = [&](int64_t j) {     // third layer of the j-th tile of this CTA
                if (j == 0) mbar_wait(w3_full, 0);
                mbar_wait(h2_full, (uint32_t)(j & 1));
                tc_fence_after();
                const uint32_t d3 = tmem_base + (uint32_t)Cfg::ACC3_COL;
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const uint32_t a_hi = smem_u32(h2 + sl * (2 * A_TILE_BYTES));
                    const uint32_t a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t w_hi = smem_u32(w3 + sl * (2 * Cfg::W3_TILE));
                    const uint32_t w_lo = w_hi + Cfg::W3_TILE;
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        const uint32_t ko = kk * 32;
                        umma_f16(d3, umma_desc(a_lo + ko), umma_desc(w_hi + ko), idesc3, (sl | kk) != 0);
                        umma_f16(d3, umma_desc(a_hi + ko), umma_desc(w_lo + ko), idesc3, 1);
                        umma_f16(d3, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc3, 1);
                    }
                }
                umma_commit(z_full);
            };
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int buf = (int)(it & 1);
                mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N2);
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait(&full_a[stage], phase);
                    mbar_wait(&full_w[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t w_hi = a_hi + 2 * A_TILE_BYTES;
                    const uint32_t w_lo = w_hi + Cfg::W2_TILE;
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        const uint32_t ko = kk * 32;
                        umma_f16(d_tmem, umma_desc(a_lo + ko), umma_desc(w_hi + ko), idesc2, (slab | kk) != 0);
                        umma_f16(d_tmem, umma_desc(a_hi + ko), umma_desc(w_lo + ko), idesc2, 1);
                        umma_f16(d_tmem, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc2, 1);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[buf]);
                if (it > 0) issue_third(it - 1);    // behind this tile's second layer: its H2 was converted meanwhile
            }
            if (it > 0) issue_third(it - 1);
        }
    } else {
        // ===================== W loader =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(w3_full, (uint32_t)Cfg::W3_BYTES);
            bulk_g2s(w3, Wp3, (uint32_t)Cfg::W3_BYTES, w3_full);
            int stage = 0;
            uint32_t phase = 0;
            constexpr uint32_t slab_bytes = 2 * Cfg::W2_TILE;
            const unsigned char *Xt = reinterpret_cast<const unsigned char *>(X);
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (TILED_A) {
                        mbar_arrive_expect_tx(&full_a[stage], 2 * A_TILE_BYTES);
                        bulk_g2s(smem + stage * Cfg::STAGE_BYTES, Xt + ((size_t)tile * KS + slab) * (2 * A_TILE_BYTES), 2 * A_TILE_BYTES,
                                 &full_a[stage]);
                    }
                    mbar_arrive_expect_tx(&full_w[stage], slab_bytes);
                    bulk_g2s(smem + stage * Cfg::STAGE_BYTES + 2 * A_TILE_BYTES, Wp2 + (size_t)slab * slab_bytes, slab_bytes, &full_w[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    }
teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// ---- weight packing -----------------------------------------------------------------------------

__global__ void absmax_kernel(const float *__restrict__ w, int64_t count, float *__restrict__ out) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));   // m >= 0: int order == float order
}

// W[N,K] fp32 → per slab: hi tile [N x 64 fp16, SW128 K-major] then lo tile, values scaled by `scale`
__global__ void pack_w_kernel(const float *__restrict__ W, int N, int K, int KS, float scale, unsigned char *__restrict__ out) {
    const int64_t total = (int64_t)KS * N * (BK / 8);           // 16-byte chunks per (hi or lo)
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(u % (BK / 8));
        const int nrow = (int)((u / (BK / 8)) % N);
        const int slab = (int)(u / ((int64_t)(BK / 8) * N));
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = slab * BK + c8 * 8 + 2 * t;
            float a = k < K ? W[(int64_t)nrow * K + k] * scale : 0.0f;
            float b = k + 1 < K ? W[(int64_t)nrow * K + k + 1] * scale : 0.0f;
            split2(a, b, hi[t], lo[t]);
        }
        const size_t tile_bytes = (size_t)N * BK * 2;
        unsigned char *base = out + (size_t)slab * 2 * tile_bytes;
        const size_t off = (size_t)(nrow >> 3) * 1024 + (nrow & 7) * 128 + ((c8 ^ (nrow & 7)) << 4);
        *reinterpret_cast<uint4 *>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(base + tile_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

int ensure_packed(Linear &l, cudaStream_t s) {
    if (l.W_tc) return 0;
    RQB_CHECK(l.in % 8 == 0, "tensor-core path needs in_features %% 8 == 0 (got %d)", l.in);
    RQB_CHECK(l.out == 32 || l.out == 64 || l.out == 128 || l.out == 256,
              "tensor-core path supports out_features 32/64/128/256 (got %d)", l.out);
    const int KS = (l.in + BK - 1) / BK;
    const size_t bytes = (size_t)KS * 2 * l.out * BK * 2;
    float *dmax = nullptr;
    RQB_CUDA(cudaMalloc(&dmax, sizeof(float)));
    RQB_CUDA(cudaMemsetAsync(dmax, 0, sizeof(float), s));
    count_launch();
    absmax_kernel<<<64, 256, 0, s>>>(l.W, (int64_t)l.in * l.out, dmax);
    float hmax = 0.0f;
    RQB_CUDA(cudaMemcpyAsync(&hmax, dmax, sizeof(float), cudaMemcpyDeviceToHost, s));
    RQB_CUDA(cudaStreamSynchronize(s));
    RQB_CUDA(cudaFree(dmax));
    RQB_CHECK(hmax == hmax && hmax < 3.0e38f, "non-finite weight");
    // scale so that max|W'| lies in [2^13, 2^14): the lo halves stay normal fp16 numbers
    int e = 0;
    if (hmax > 0.0f) { frexpf(hmax, &e); e = 14 - e; }
    if (e > 40) e = 40;
    if (e < -20) e = -20;
    l.tc_scale_exp = e;
    void *p = nullptr;
    RQB_CUDA(cudaMalloc(&p, bytes));
    count_launch();
    pack_w_kernel<<<kNumSMs, 256, 0, s>>>(l.W, l.out, l.in, KS, ldexpf(1.0f, e), (unsigned char *)p);
    RQB_LAUNCH_CHECK();
    l.W_tc = p;
    l.W_tc_bytes = bytes;
    return 0;
}

template <int N, int PASSES>
int launch_tc(const Linear &l, const float *x, int64_t n, float *y, bool relu, const unsigned long long *n_dev, cudaStream_t s) {
    using Cfg = TcCfg<N, PASSES>;
    auto kern = linear_tc_kernel<N, PASSES>;
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    const int64_t ntiles = (n + TM - 1) / TM;
    const unsigned grid = (unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs);
    count_launch();
    kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(x, n, l.in, (const unsigned char *)l.W_tc, l.b, ldexpf(1.0f, -l.tc_scale_exp),
                                                  relu ? 1 : 0, y, n_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

// flagged rows → compact list
__global__ void gate_kernel(const float *__restrict__ margin, int64_t n, int64_t *__restrict__ list,
                            unsigned long long *__restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool flag = i < n && !(margin[i] > 0.0f);        // margin already has the threshold subtracted; NaN flags too
    unsigned ballot = __ballot_sync(0xffffffffu, flag);
    if (ballot == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (flag) list[base + __popc(ballot & ((1u << lane) - 1u))] = i;
}

}  // namespace

static int g_tc_debug_host = 0;
int tc_debug_flags() { return g_tc_debug_host; }
int tc_set_debug(int flags) {
    RQB_CUDA(cudaMemcpyToSymbol(g_tc_debug, &flags, sizeof(flags)));
    g_tc_debug_host = flags;
    return 0;
}

int tc_set_trace(long long *buf) {
    (void)buf;
    set_error("timeline trace is not compiled into this build");
    return RQB200_EINVAL;
}

static bool env_tc2() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("RQB200_TC2"); v = (e && e[0] == '0') ? 0 : ((e && e[0] == '1') ? 1 : RQB200_TC2_DEFAULT); }
    return v == 1;
}

// passes: 3 (split-fp16, fp32-class) or 1 (fp16 screening pass).  rows: gather of the input rows (first layer of a
// re-run tier; needs the 2-CTA kernel).  n_dev: device-resident row count, n = upper bound.
static bool env_untiled() {      // RQB200_UNTILED=1: fp32 rows between the first layer and the fused kernel (A/B measurements)
    static int v = -1;
    if (v < 0) { const char *e = getenv("RQB200_UNTILED"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
static bool env_unfused() {      // RQB200_UNFUSED=1: keep the last two layers as separate kernels (A/B measurements)
    static int v = -1;
    if (v < 0) { const char *e = getenv("RQB200_UNFUSED"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int linear_tc(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, int passes, const int64_t *rows,
              const unsigned long long *n_dev, bool tiled_out) {
    if (n == 0) return 0;
    RQB_CHECK(passes == 1 || passes == 3, "passes must be 1 or 3");
    RQB_TRY(ensure_packed(l, s));
    if (linear_tc2_supported(l) && (env_tc2() || rows || tiled_out)) return linear_tc2(l, x, n, y, relu, s, passes, rows, n_dev, tiled_out);
    RQB_CHECK(rows == nullptr && !tiled_out, "row gather / tiled output need the 2-CTA kernel (out_features 256)");
#define RQB_TC_CASE(NN)                                                                              \
    case NN: return passes == 1 ? launch_tc<NN, 1>(l, x, n, y, relu, n_dev, s) : launch_tc<NN, 3>(l, x, n, y, relu, n_dev, s);
    switch (l.out) {
        RQB_TC_CASE(32)
        RQB_TC_CASE(64)
        RQB_TC_CASE(128)
        RQB_TC_CASE(256)
    }
#undef RQB_TC_CASE
    set_error("unsupported out_features %d", l.out);
    return RQB200_EINVAL;
}


template <int N3, bool TILED_A>
static int launch_mlp23(Linear &l2, Linear &l3, const float *x, int64_t n, float *z, const unsigned long long *n_dev, cudaStream_t s) {
    using Cfg = Tc23Cfg<N3>;
    auto kern = mlp23_tc_kernel<N3, TILED_A>;
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    const int64_t ntiles = (n + TM - 1) / TM;
    const unsigned grid = (unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs);
    count_launch();
    kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(x, n, l2.in, (const unsigned char *)l2.W_tc, l2.b, ldexpf(1.0f, -l2.tc_scale_exp),
                                                  (const unsigned char *)l3.W_tc, l3.b, ldexpf(1.0f, -l3.tc_scale_exp), z, n_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

// layers i and i+1 of a three-pass MLP can run as one kernel when the middle width is 128 and the output 32 or 64
static bool mlp23_supported(const Linear &l2, const Linear &l3) {
    return l2.out == 128 && l3.in == 128 && (l3.out == 32 || l3.out == 64) && l2.in % 8 == 0;
}

int mlp_tc(rqb200_model *m, int which, const float *x, int64_t n, float *y, cudaStream_t s, int passes,
           const int64_t *rows, const unsigned long long *n_dev, bool profile) {
    Linear *ls = which == 0 ? m->enc : m->dec;
    int maxdim = 0;
    for (int i = 0; i + 1 < m->n_layers; ++i) maxdim = ls[i].out > maxdim ? ls[i].out : maxdim;
    const size_t n_pad = (size_t)((n + 255) / 256) * 256;      // the tiled hand-off is written in whole 256-row pair tiles
    if (m->n_layers > 1) {
        RQB_TRY(ws_reserve(m->act[0], sizeof(float) * n_pad * maxdim));
        if (m->n_layers > 2) RQB_TRY(ws_reserve(m->act[1], sizeof(float) * n_pad * maxdim));
    }
    // passes == 2: the screening tier — first layer as ONE TF32 pass fed by TMA (encode_tf32.cu), the tail three-pass
    const bool tf32_first = passes == 2;
    if (tf32_first) {
        RQB_CHECK(which == 0 && rows == nullptr && n_dev == nullptr && linear_tf32_supported(ls[0]), "TF32 screening pass: unsupported call");
        passes = 3;
    }
    const float *cur = x;
    bool cur_tiled = false;             // `cur` holds split-fp16 UMMA tiles (written by the 2-CTA kernel for the fused kernel)
    for (int i = 0; i < m->n_layers; ++i) {
        const bool last = i == m->n_layers - 1;
        float *dst = last ? y : (float *)m->act[i & 1].ptr;
        ProfScope ps(!profile ? -1 : (i == 0 ? PROF_TC_ENCODER : PROF_TC_REST), s);      // slot 4 = the wide first layer alone
        if (i > 0 && i + 2 == m->n_layers && passes == 3 && mlp23_supported(ls[i], ls[i + 1]) && !env_unfused()) {
            // the last two layers as one kernel: the 128-wide activation between them never leaves the SM
            RQB_TRY(ensure_packed(ls[i], s));
            RQB_TRY(ensure_packed(ls[i + 1], s));
            if (ls[i + 1].out == 32) {
                if (cur_tiled) RQB_TRY((launch_mlp23<32, true>(ls[i], ls[i + 1], cur, n, y, n_dev, s)));
                else RQB_TRY((launch_mlp23<32, false>(ls[i], ls[i + 1], cur, n, y, n_dev, s)));
            } else {
                if (cur_tiled) RQB_TRY((launch_mlp23<64, true>(ls[i], ls[i + 1], cur, n, y, n_dev, s)));
                else RQB_TRY((launch_mlp23<64, false>(ls[i], ls[i + 1], cur, n, y, n_dev, s)));
            }
            return 0;
        }
        // first layer of a 3-layer three-pass MLP whose tail is fused: emit the tiles the fused kernel bulk-loads
        const bool tiled = i == 0 && m->n_layers == 3 && passes == 3 && linear_tc2_supported(ls[0]) && ls[0].out == 256 &&
                           mlp23_supported(ls[1], ls[2]) && !env_unfused() && !env_untiled();
        if (i == 0 && tf32_first) RQB_TRY(linear_tf32(ls[0], cur, n, dst, !last, s, tiled));
        else RQB_TRY(linear_tc(ls[i], cur, n, dst, !last, s, passes, i == 0 ? rows : nullptr, n_dev, tiled));
        cur = dst;
        cur_tiled = tiled;
    }
    return 0;
}

__global__ void scatter_rows_kernel(const float *__restrict__ src, const int64_t *__restrict__ rows, int64_t nr, int e,
                                    float *__restrict__ dst) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nr * e) return;
    int64_t r = p / e;
    dst[rows[r] * e + (p - r * e)] = src[p];
}
int scatter_rows(const float *src, const int64_t *rows, int64_t nr, int e, float *dst, cudaStream_t s) {
    count_launch();
    scatter_rows_kernel<<<(unsigned)((nr * e + 255) / 256), 256, 0, s>>>(src, rows, nr, e, dst);
    RQB_LAUNCH_CHECK();
    return 0;
}
__global__ void scatter_rows_dev_kernel(const float *__restrict__ src, const int64_t *__restrict__ rows,
                                        const unsigned long long *__restrict__ nr_dev, int e, float *__restrict__ dst) {
    const int64_t total = (int64_t)*nr_dev * e;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = p / e;
        dst[rows[r] * e + (p - r * e)] = src[p];
    }
}

// get_indices, fast route.  Three tiers, every one certifying the rows it keeps with the margin gate of the quantizer:
//   tier 1 (screening, optional): ONE fp16 tensor pass over every row (operands rounded to fp16: |z~ - z| <~ 2^-10 |z|),
//           rows whose top-2 gap at any level is inside the gate for gamma1 are appended to list1;
//   tier 2: list1 rows (all rows when screening is off) re-run with split-fp16, three passes (|z~ - z| <~ 2^-16.7 |z|),
//           gate gamma → list2.  Launched straight behind tier 1: its row count stays on the device;
//   tier 3: list2 rows re-run by the exact SIMT kernels in the reference's summation order.
// Codes are therefore identical to the exact route; only the route differs.
int get_indices_fast(rqb200_model *m, const float *x, int64_t n, int64_t *codes, float *z_out,
                     int64_t *stats_host, cudaStream_t s) {
    // workspace: z~[n,e] | z2[n,e] (compact, tier 2) | list1[n] | list2[n] | count1, count2
    const size_t zb = (sizeof(float) * (size_t)n * m->e + 255) & ~(size_t)255;
    const size_t lb = (sizeof(int64_t) * (size_t)n + 255) & ~(size_t)255;
    // Screening pays when the first layer dominates the row: the TF32 pass saves about a third of that layer, the rows it
    // cannot certify (5-15 %) pay the three-pass encoder AND the tensor-core quantizer a second time.  Measured
    // (profiles/r2_check_tf32_*.txt): a gain with 3 x 256 codes of 32 dims, a loss with 4 x 1024 codes of 64 dims.
    long long quant_work = 0;
    for (int l = 0; l < m->L; ++l) quant_work += (long long)m->K[l] * m->e;
    const bool screen_on = m->screen_auto ? (m->screen_kind == 1 && linear_tf32_supported(m->enc[0]) && m->n_layers == 3 &&
                                             quant_work <= 32768)
                                          : m->screen_enabled;
    const bool screen = screen_on && quantize_tc_supported(m) && !m->force_simt_quantizer;
    RQB_TRY(ws_reserve(m->misc, 2 * zb + 2 * lb + 512));
    char *base = (char *)m->misc.ptr;
    float *z1 = z_out ? z_out : (float *)base;
    float *z2 = (float *)(base + zb);
    int64_t *list1 = (int64_t *)(base + 2 * zb);
    int64_t *list2 = (int64_t *)(base + 2 * zb + lb);
    unsigned long long *counts = (unsigned long long *)(base + 2 * zb + 2 * lb);
    RQB_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), s));
    unsigned long long h[2] = {0, 0};
    if (screen) {
        const bool tf32 = m->screen_kind == 1 && linear_tf32_supported(m->enc[0]) && (((uintptr_t)x & 15) == 0) && n < ((int64_t)1 << 31);
        RQB_TRY(mlp_tc(m, 0, x, n, z1, s, tf32 ? 2 : 1, nullptr, nullptr, true));
        {
            ProfScope ps(PROF_QUANTIZE, s);
            RQB_TRY(quantize_tc(m, z1, n, codes, list1, counts, s, m->screen_gamma));
        }
        {
            ProfScope ps(PROF_TIER2, s);
            RQB_TRY(mlp_tc(m, 0, x, n, z2, s, 3, list1, counts, false));
            RQB_TRY(quantize_tc(m, z2, n, codes, list2, counts + 1, s, m->gate_gamma, list1, counts));
            if (z_out) {
                count_launch();
                scatter_rows_dev_kernel<<<kNumSMs * 4, 256, 0, s>>>(z2, list1, counts, m->e, z_out);
                RQB_LAUNCH_CHECK();
            }
        }
    } else {
        RQB_TRY(mlp_tc(m, 0, x, n, z1, s, 3, nullptr, nullptr, true));
        if (quantize_tc_supported(m) && !m->force_simt_quantizer) {
            ProfScope ps(PROF_QUANTIZE, s);
            RQB_TRY(quantize_tc(m, z1, n, codes, list2, counts + 1, s, m->gate_gamma));   // distances on the tensor cores, gate fused
        } else {
            float *margin = z2;                                                          // [n] floats, z2 is unused here
            RQB_TRY(quantize_exact(m, z1, n, codes, nullptr, nullptr, nullptr, nullptr, margin, s));   // margin - threshold
            count_launch();
            gate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(margin, n, list2, counts + 1);
            RQB_LAUNCH_CHECK();
        }
    }
    // Tier 3: the exact SIMT kernels over the rows of list2 — gather → exact MLP → exact quantizer → scatter codes.  The row
    // count stays on the device (counts[1]): the kernels run persistent grids over "row tiles below the count", their
    // activations live in the MLP's own ping-pong buffers (free by now, sized for every row) and the latent in z2, so no size
    // has to come back to the host and nothing in this call waits for the device.
    {
        ProfScope ps(PROF_RESCUE, s);
        const unsigned long long *nr_dev = counts + 1;
        // about how many rows will come through (picks the quantizer mapping and the tile shape of the exact Linears, nothing
        // else): the last count that was fetched, else the typical share of the route — 0.7 % of the rows behind the screening
        // tier, 1.6-2.1 % without it
        const int64_t hint = m->last_tier_rows[1] > 0 ? m->last_tier_rows[1] : (screen ? n / 128 : n / 48) + 1;
        const int64_t lin_hint = hint;
        float *zr = z2;                                                    // [n, e]: tier 2 has consumed it
        const float *cur = x;
        for (int i = 0; i < m->n_layers; ++i) {
            const bool last = i == m->n_layers - 1;
            float *dst = last ? zr : (float *)m->act[i & 1].ptr;
            RQB_TRY(linear_exact(m->enc[i], cur, i == 0 ? list2 : nullptr, n, dst, !last, s, n < 16 ? 16 : n, nr_dev, lin_hint));   // order of the n-row batch
            cur = dst;
        }
        RQB_TRY(quantize_exact(m, zr, n, codes, list2, nullptr, nullptr, nullptr, nullptr, s, n < 16 ? 16 : n, nr_dev, hint));
        if (z_out) {                                                       // keep the caller's latent exact on rescued rows
            count_launch();
            scatter_rows_dev_kernel<<<kNumSMs * 4, 256, 0, s>>>(zr, list2, nr_dev, m->e, z_out);
            RQB_LAUNCH_CHECK();
        }
    }
    // the tier row counts: kept on the device for rqb200_model_last_tier_rows; read back here only if the caller asks
    RQB_CUDA(cudaMemcpyAsync(m->tier_counts_dev, counts, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    m->tier_counts_stream = s;
    m->tier_counts_pending = true;
    if (stats_host) {
        RQB_CUDA(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, s));
        RQB_CUDA(cudaStreamSynchronize(s));
        stats_host[0] = (int64_t)h[1];
        m->last_tier_rows[0] = (int64_t)h[0];
        m->last_tier_rows[1] = (int64_t)h[1];
        m->tier_counts_pending = false;
    }
    return 0;
}

}  // namespace rqb

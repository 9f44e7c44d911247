// encode_tc2.cu — 2-CTA (cta_group::2) variant of the tensor-core Linear for the wide first encoder layer.
//
// Same math as linear_tc_kernel (encode_tc.cu; replaces reference RQ-VAE/models/layers.py:23 for the
// 768→256 / 1024→256 layer): split-fp16 operands, three MMAs per K-step into an fp32 TMEM accumulator.
// Why a CTA pair: with both operands in shared memory the three passes re-read A (4 KB) and B (8 KB) per
// MMA and, together with the producers' stores and the W bulk writes, saturate the 128 B/clk shared-memory
// port of one SM (tools/ablate_tc.py).  In cta_group::2 mode the pair computes a 256-row tile: each CTA
// stages its own 128 rows of A and only HALF of W (128 of the 256 output features), so per SM the B reads,
// the W bulk writes and the W traffic from L2 are halved, and the smaller stage (64 KB) allows 3 stages.
//
// Pair protocol (rank 0 = leader, issues every MMA):
//   ready[s]       local: 8 producer warps + the W bulk copy (tx) of this CTA
//   peer_ready[s]  leader only: the peer's relay lane forwards "my ready[s] completed" with a remote arrive
//   empty[s], tmem_full[b]   tcgen05.commit multicast to both CTAs
//   tmem_empty[b]  leader only: 128 local + 128 remote epilogue arrivals
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace rqb {

namespace {

constexpr int TM2 = 128;                   // rows per CTA (256 per pair)
constexpr int BK2 = 64;
constexpr int N2 = 256;                    // output features of the pair tile
constexpr int NH2 = N2 / 2;                // W rows staged per CTA
#ifndef T2_CONV_WARPS
#define T2_CONV_WARPS 16
#endif
constexpr int T2_EPI = 4, T2_CONV = T2_CONV_WARPS;       // 16 producer warps: more independent load streams per SM
constexpr int T2_THREADS = (T2_EPI + T2_CONV + 2) * 32;
constexpr int T2_NF4 = TM2 * BK2 / 4 / (T2_CONV * 32);      // float4 per producer thread per slab
constexpr int T2_RSTEP = T2_CONV * 32 / 16;                 // rows covered by one pass of the producers
constexpr int T2_MMA_WARP = T2_EPI + T2_CONV, T2_LOAD_WARP = T2_MMA_WARP + 1;
constexpr int T2_A_TILE = TM2 * BK2 * 2;   // 16 KB (hi or lo)
constexpr int T2_W_TILE = NH2 * BK2 * 2;   // 16 KB (hi or lo, this CTA's half)
// PASSES = 3: split-fp16 (hi+lo) operands, three MMAs per K step;  PASSES = 1: hi halves only (tier-1 screening pass)
template <int PASSES> struct T2Cfg {
    static constexpr int NSPLIT = PASSES == 1 ? 1 : 2;                    // tiles per operand (hi [, lo])
    static constexpr int STAGE = NSPLIT * (T2_A_TILE + T2_W_TILE);        // 64 KB / 32 KB
    static constexpr int STAGES = PASSES == 1 ? 6 : 3;
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256 + T2_EPI * 32 * EPI_LD * 4;
};
#ifndef T2_PREFETCH_N
#define T2_PREFETCH_N 2
#endif
constexpr int T2_PREFETCH = T2_PREFETCH_N;     // K slabs of X in flight per producer thread (registers)
constexpr int T2_TMEM_COLS = 512;          // two 256-column accumulator buffers

// rows (may be NULL): gather — tile row i reads X[rows[i]] (tier-2 re-run of gated rows), Y stays compact.
// n_dev (may be NULL): the row count lives on the device (written by the previous tier's gate), n is its upper bound.
template <int PASSES, bool GATHER>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
linear_tc2_kernel(const float *__restrict__ X, int64_t n, int K, const unsigned char *__restrict__ Wp2,
                  const float *__restrict__ bias, float inv_scale, int relu, float *__restrict__ Y,
                  const int64_t *__restrict__ rows, const unsigned long long *__restrict__ n_dev, int dbg, int tiled_out) {
    // tiled_out: Y receives the activation as split-fp16 UMMA tiles (epilogue_rows_split) instead of fp32 rows
    // dbg: ablation switches of tools/ablate_tc2.py (0 in production) — bit0 no epilogue stores, bit1 no MMA,
    // bit2 no producer smem stores, bit3 no W bulk loads, bit4 no X loads
    constexpr int T2_STAGE = T2Cfg<PASSES>::STAGE, T2_STAGES = T2Cfg<PASSES>::STAGES;
    if (GATHER && n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + T2_STAGES * T2_STAGE);
    uint64_t *ready = bars;                              // [STAGES]
    uint64_t *peer_ready = bars + T2_STAGES;             // [STAGES] (used in the leader)
    uint64_t *empty = bars + 2 * T2_STAGES;              // [STAGES]
    uint64_t *tmem_full = bars + 3 * T2_STAGES;          // [2]
    uint64_t *tmem_empty = tmem_full + 2;                // [2] (used in the leader)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int KS = (K + BK2 - 1) / BK2;
    const int64_t npt = (n + 2 * TM2 - 1) / (2 * TM2);   // pair tiles of 256 rows
    const int64_t pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_STAGES; ++s) {
            mbar_init(&ready[s], T2_CONV + 1);
            mbar_init(&peer_ready[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 2 * T2_EPI * 32);
        }
        fence_barrier_init();
    }
    if (warp == T2_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)T2_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < T2_EPI) {
        // ===================== epilogue (each CTA drains its own 128 rows) =====================
        float *patch = reinterpret_cast<float *>(smem + T2_STAGES * T2_STAGE + 256) + warp * (32 * EPI_LD);
        int64_t it = 0;
        for (int64_t pt = pair0; pt < npt; pt += npairs, ++it) {
            const int buf = (int)(it & 1);
            mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            if (tiled_out)
                epilogue_rows_split<N2>(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * N2), patch, lane, bias, inv_scale,
                                        relu, reinterpret_cast<unsigned char *>(Y), pt * 2 + rank, warp * 32);
            else
                epilogue_rows<N2>(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * N2), patch, lane, bias, inv_scale, relu,
                                  Y, pt * (2 * TM2) + rank * TM2 + warp * 32, n, !(dbg & 1));
            tc_fence_before();
            if (rank == 0) mbar_arrive(&tmem_empty[buf]);
            else mbar_arrive_remote(&tmem_empty[buf], 0);
        }
    } else if (warp < T2_EPI + T2_CONV) {
        // ===================== A producers (own 128 rows) =====================
        const int ct = threadIdx.x - T2_EPI * 32;
        const int c4 = ct & 15;
        const int rbase = ct >> 4;                           // rows rbase + T2_RSTEP*i
        const int64_t my_tiles = pair0 < npt ? (npt - pair0 + npairs - 1) / npairs : 0;
        const int64_t steps = my_tiles * KS;
        auto load_slab = [&](int64_t st, float4 (&dst)[T2_NF4]) {
            const int64_t pt = pair0 + (st / KS) * npairs;
            const int k0 = (int)(st % KS) * BK2 + c4 * 4;
            int64_t src[T2_NF4];
#pragma unroll
            for (int i = 0; i < T2_NF4; ++i) {                   // (the row indices first, as one batch of independent loads)
                const int64_t row = pt * (2 * TM2) + rank * TM2 + rbase + T2_RSTEP * i;
                src[i] = row;
                if (GATHER) src[i] = row < n ? __ldg(rows + row) : 0;
            }
#pragma unroll
            for (int i = 0; i < T2_NF4; ++i) {
                const int64_t row = pt * (2 * TM2) + rank * TM2 + rbase + T2_RSTEP * i;
                if (row < n && k0 < K && !(dbg & 16)) dst[i] = __ldg(reinterpret_cast<const float4 *>(X + src[i] * (int64_t)K + k0));
                else dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int stage = 0;
        uint32_t phase = 0;
        auto convert_slab = [&](const float4 (&src)[T2_NF4]) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char *a_hi = smem + stage * T2_STAGE;
            unsigned char *a_lo = a_hi + T2_A_TILE;
#pragma unroll
            for (int i = 0; i < T2_NF4; ++i) {
                const int r = rbase + T2_RSTEP * i;
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
            if (!(dbg & 32)) fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[stage]);
            if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
        };
        // T2_PREFETCH register buffers in rotation; a buffer is refilled right after it has been converted.
        // (Measured: 3 buffers or an L2 bulk prefetch of the next tile are both slower than 2 buffers.)
        float4 buf[T2_PREFETCH][T2_NF4];
#pragma unroll
        for (int d = 0; d < T2_PREFETCH; ++d)
            if (d < steps) load_slab(d, buf[d]);
        for (int64_t st = 0; st < steps; st += T2_PREFETCH) {
#pragma unroll
            for (int d = 0; d < T2_PREFETCH; ++d) {
                if (st + d < steps) {
                    convert_slab(buf[d]);
                    if (st + d + T2_PREFETCH < steps) load_slab(st + d + T2_PREFETCH, buf[d]);
                }
            }
        }
    } else if (warp == T2_MMA_WARP) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if (rank == 0) {
                // ===================== MMA issuer (leader) =====================
                const uint32_t idesc = umma_idesc(2 * TM2, N2);
                int64_t it = 0;
                for (int64_t pt = pair0; pt < npt; pt += npairs, ++it) {
                    const int buf = (int)(it & 1);
                    mbar_wait_cluster(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N2);
                    for (int slab = 0; slab < KS; ++slab) {
                        mbar_wait(&ready[stage], phase);
                        if (dbg & 256) mbar_wait(&peer_ready[stage], phase);
                        else mbar_wait_cluster(&peer_ready[stage], phase);
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(smem + stage * T2_STAGE);
                        const uint32_t a_lo = a_hi + T2_A_TILE;
                        const uint32_t w_hi = a_hi + T2Cfg<PASSES>::NSPLIT * T2_A_TILE;
                        const uint32_t w_lo = w_hi + T2_W_TILE;
                        if (!(dbg & 2))
#pragma unroll
                        for (int kk = 0; kk < BK2 / 16; ++kk) {
                            const uint32_t ko = kk * 32;
                            if (PASSES == 1) {
                                umma_f16_2cta(d_tmem, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc, (slab | kk) != 0);
                            } else {
                                umma_f16_2cta(d_tmem, umma_desc(a_lo + ko), umma_desc(w_hi + ko), idesc, (slab | kk) != 0);
                                umma_f16_2cta(d_tmem, umma_desc(a_hi + ko), umma_desc(w_lo + ko), idesc, 1);
                                umma_f16_2cta(d_tmem, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc, 1);
                            }
                        }
                        umma_commit_2cta(&empty[stage]);
                        if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_2cta(&tmem_full[buf]);
                }
            } else {
                // ===================== relay (peer): forward "stage ready" to the leader =====================
                for (int64_t pt = pair0; pt < npt; pt += npairs) {
                    for (int slab = 0; slab < KS; ++slab) {
                        mbar_wait(&ready[stage], phase);
                        if (dbg & 128) mbar_arrive_remote_relaxed(&peer_ready[stage], 0);
                        else mbar_arrive_remote(&peer_ready[stage], 0);
                        if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ===================== W loader (this CTA's half of the output features) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            constexpr uint32_t half_bytes = 2 * T2_W_TILE;              // hi | lo of this CTA's 128 features (packed image)
            constexpr uint32_t copy_bytes = T2Cfg<PASSES>::NSPLIT * T2_W_TILE;   // one pass needs the hi tile only
            for (int64_t pt = pair0; pt < npt; pt += npairs) {
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (dbg & 8) {
                        mbar_arrive(&ready[stage]);
                    } else {
                        mbar_arrive_expect_tx(&ready[stage], copy_bytes);
                        bulk_g2s(smem + stage * T2_STAGE + T2Cfg<PASSES>::NSPLIT * T2_A_TILE, Wp2 + ((size_t)slab * 2 + rank) * half_bytes,
                                 copy_bytes, &ready[stage]);
                    }
                    if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    }
    // ---- teardown: neither CTA may leave while the pair can still touch its smem / TMEM
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == T2_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)T2_TMEM_COLS));
    }
}

// W[256,K] fp32 → per slab: [features 0..127: hi | lo][features 128..255: hi | lo], SW128 K-major tiles of 128 x 64 fp16
__global__ void pack_w2_kernel(const float *__restrict__ W, int K, int KS, float scale, unsigned char *__restrict__ out) {
    const int64_t total = (int64_t)KS * N2 * (BK2 / 8);
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(u % (BK2 / 8));
        const int nrow = (int)((u / (BK2 / 8)) % N2);
        const int slab = (int)(u / ((int64_t)(BK2 / 8) * N2));
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = slab * BK2 + c8 * 8 + 2 * t;
            const float a = k < K ? W[(int64_t)nrow * K + k] * scale : 0.0f;
            const float b = k + 1 < K ? W[(int64_t)nrow * K + k + 1] * scale : 0.0f;
            split2(a, b, hi[t], lo[t]);
        }
        const int half = nrow / NH2, r = nrow % NH2;
        unsigned char *base = out + ((size_t)slab * 2 + half) * (2 * T2_W_TILE);
        const size_t off = (size_t)(r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4);
        *reinterpret_cast<uint4 *>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(base + T2_W_TILE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

}  // namespace

bool linear_tc2_supported(const Linear &l) { return l.out == N2 && l.in % 8 == 0; }

int tc_debug_flags();     // encode_tc.cu

template <int PASSES, bool GATHER>
static int launch_tc2(Linear &l, const float *x, int64_t n, float *y, bool relu, const int64_t *rows,
                      const unsigned long long *n_dev, cudaStream_t s, bool tiled_out) {
    auto kern = linear_tc2_kernel<PASSES, GATHER>;
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Cfg<PASSES>::SMEM));
    }
    const int64_t npt = (n + 2 * TM2 - 1) / (2 * TM2);
    int64_t pairs = npt < kNumSMs / 2 ? npt : kNumSMs / 2;
    count_launch();
    kern<<<(unsigned)(pairs * 2), T2_THREADS, T2Cfg<PASSES>::SMEM, s>>>(x, n, l.in, (const unsigned char *)l.W_tc2, l.b,
                                                                       ldexpf(1.0f, -l.tc_scale_exp), relu ? 1 : 0, y, rows, n_dev,
                                                                       tc_debug_flags(), tiled_out ? 1 : 0);
    RQB_LAUNCH_CHECK();
    return 0;
}

// requires l.tc_scale_exp to be set (ensure_packed of encode_tc.cu ran).  n is an upper bound when n_dev is given.
int linear_tc2(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, int passes,
               const int64_t *rows, const unsigned long long *n_dev, bool tiled_out) {
    if (n == 0) return 0;
    const int KS = (l.in + BK2 - 1) / BK2;
    if (!l.W_tc2) {
        void *p = nullptr;
        RQB_CUDA(cudaMalloc(&p, (size_t)KS * 2 * (2 * T2_W_TILE)));
        count_launch();
        pack_w2_kernel<<<kNumSMs, 256, 0, s>>>(l.W, l.in, KS, ldexpf(1.0f, l.tc_scale_exp), (unsigned char *)p);
        RQB_LAUNCH_CHECK();
        l.W_tc2 = p;
    }
    if (rows) {
        RQB_CHECK(passes == 3, "row gather is only built for the three-pass kernel");
        return launch_tc2<3, true>(l, x, n, y, relu, rows, n_dev, s, tiled_out);
    }
    RQB_CHECK(n_dev == nullptr, "a device-side row count needs the gather variant");
    return passes == 1 ? launch_tc2<1, false>(l, x, n, y, relu, nullptr, nullptr, s, tiled_out)
                       : launch_tc2<3, false>(l, x, n, y, relu, nullptr, nullptr, s, tiled_out);
}

}  // namespace rqb

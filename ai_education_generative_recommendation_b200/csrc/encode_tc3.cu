// encode_tc3.cu — CTA-pair tensor-core Linear for the wide first encoder layer with the A operand in TENSOR MEMORY.
//
// STATUS: EXPERIMENTAL, OFF BY DEFAULT (RQB200_TC3=1 selects it).  Written at the end of round 1 from the round-1
// measurements; it compiles for sm_100a but has NOT run on a B200 yet.  tools/check_tc3.py is the first thing to run:
// the kernel issues the same MMAs in the same order as linear_tc2_kernel, so its output must be BIT-IDENTICAL.
// The tensor-memory layout it relies on (tcgen05.st.16x256b fragments forming the MMA's A operand) is checked on its own
// by tools/tmem_a_probe.cu (mode "TS 16x256b") and tools/tmem_a_probe2.cu (CTA pair).  Both layout assumptions were cross-read
// against the CuTe headers vendored in the image (read only, nothing included): the (thread, register) → (lane, column) map of
// tcgen05.st.16x256b.x2 is Copy_Traits<SM100_TMEM_LOAD_16dp256b2x>::DstLayout (cute/atom/copy_traits_sm100.hpp: thread =
// (t%4 → 64-bit chunk of the row, t/4 → row), values = (64 bits, row + 8, next 8-column block)), and the A operand of a 2-SM
// M = 256 TS-form MMA is tmem_frg_2sm's "4x1" atom (cute/atom/mma_traits_sm100.hpp: lane = row of the CTA's 128, 16-bit
// elements packed along K in consecutive columns, same lane addresses as the accumulator).
//
// Same math as linear_tc2_kernel (encode_tc2.cu; replaces reference RQ-VAE/models/layers.py:23 for the 768→256 / 1024→256
// layer): split-fp16 operands, three MMAs per 16-wide K step into one fp32 accumulator, cta_group::2 (256-row pair tile,
// each CTA stages its own 128 rows of A and half of W).
//
// Why: linear_tc2_kernel is bound by the shared-memory port (DESIGN.md §4 "What bounds the dominant kernel"): per 64-wide
// K slab and SM it writes 32 KB of A (hi+lo) and 32 KB of W and the three MMA passes read 48 KB of A and 48 KB of W —
// 160 KB at 128 B/clk is 92 % of the 1536 clocks the MMAs need, so loads and MMAs cannot overlap (0.70 ms each alone,
// 1.09 ms together).  Here the converter warps write the split-fp16 A fragments straight from registers into tensor
// memory (tcgen05.st, 256 B/clk) and the MMAs take A from there ("TS" form): shared memory carries only W, 80 KB per
// slab (52 % of the port).
//
// What changes around that:
//   * Tensor memory (512 columns): the accumulator is single-buffered (columns 0..255) and columns 256..447 hold a
//     three-stage ring of A slabs (per stage 32 columns hi + 32 columns lo: 64 fp16 per row each).
//   * Because the accumulator is no longer double-buffered, the epilogue first drains it into shared memory (128 KB,
//     the space the A stages used to take; about 2 us) and releases it, then converts / stores from shared memory while
//     the next tile's MMAs run.
//   * Producer mapping follows the 16x256b tensor-memory fragment: lane t of a warp owns rows t/4 and t/4+8 of a 16-row
//     block and 4 consecutive K elements per 16-wide K group, so one LDG.128 feeds one fragment pair (hi, lo) and the
//     four lanes of a row read 64 contiguous bytes.
//
// PASSES = 1 (hi halves only, six stages of half the size) is the screening-tier instantiation: with shared memory out of the
// way a one-pass first layer may become HBM-bound, which is what would make the screening tier of the fast route pay
// (DESIGN.md §4 "Screening tier": the one-pass linear_tc2_kernel was not faster because it was not tensor-bound).
//
// Pair protocol: as in encode_tc2.cu (ready / peer_ready / empty per stage, tmem_full / tmem_empty per accumulator).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace rqb {

namespace {

constexpr int TM3 = 128;                   // rows per CTA (256 per pair)
constexpr int BK3 = 64;
constexpr int N3 = 256;                    // output features of the pair tile
constexpr int NH3 = N3 / 2;                // W rows staged per CTA
#ifndef T3_EPI_WARPS
#define T3_EPI_WARPS 4
#endif
// Epilogue warps: 4 (one per tensor-memory lane quarter, all 256 columns) or 8 (two per quarter, 128 columns each; A/B build
// -DT3_EPI_WARPS=8: 832 threads leave 72 registers per thread).  The producers follow, so their warp id % 4 stays pw % 4.
constexpr int T3_EPI = T3_EPI_WARPS, T3_CONV = 16;
static_assert(T3_EPI == 4 || T3_EPI == 8, "4 or 8 epilogue warps");
constexpr int T3_EPI_COLS = 256 / (T3_EPI / 4);          // accumulator columns per epilogue warp
constexpr int T3_THREADS = (T3_EPI + T3_CONV + 2) * 32;
constexpr int T3_MMA_WARP = T3_EPI + T3_CONV;
constexpr int T3_W_TILE = NH3 * BK3 * 2;   // 16 KB (hi or lo, this CTA's half of the features)
constexpr int T3_RING_BYTES = 6 * T3_W_TILE;             // shared memory of the W ring (96 KB), whatever the stage size
constexpr int T3_DRAIN_ROW = N3 * 4;       // bytes per accumulator row in the drain buffer (no padding: 16-byte chunks
                                           // are XOR-swizzled with the row number instead)
constexpr int T3_DRAIN_WARP = 32 * T3_DRAIN_ROW;      // 32 KB per epilogue warp (its 32 rows)
constexpr int T3_SMEM = T3_RING_BYTES + 4 * T3_DRAIN_WARP + 256 + 1024;
static_assert(T3_SMEM <= 227 * 1024, "linear_tc3_kernel: shared memory over the per-CTA limit");
constexpr int T3_TMEM_COLS = 512;
constexpr int T3_A_COL0 = N3;              // first tensor-memory column of the A ring
constexpr int T3_A_COLS = BK3 / 2;         // 32 columns = 64 fp16 per row (hi or lo)
// PASSES = 3: split-fp16 (hi + lo) operands, three MMAs per K step, three stages of [W hi | W lo] / [A hi | A lo];
// PASSES = 1: hi halves only (screening tier of the fast route), six stages of half the size
template <int PASSES> struct T3Cfg {
    static constexpr int NSPLIT = PASSES == 1 ? 1 : 2;
    static constexpr int STAGE = NSPLIT * T3_W_TILE;                 // shared memory per stage
    static constexpr int STAGES = T3_RING_BYTES / STAGE;             // 3 / 6
    static constexpr int A_STAGE_COLS = NSPLIT * T3_A_COLS;          // tensor-memory columns per stage of the A ring
    static_assert(T3_A_COL0 + STAGES * A_STAGE_COLS <= T3_TMEM_COLS, "A ring does not fit in tensor memory");
    static_assert(4 * STAGES * 8 + 16 + 4 <= 256, "barrier block");
};
#ifndef T3_PREFETCH_N
#define T3_PREFETCH_N 2
#endif
constexpr int T3_PREFETCH = T3_PREFETCH_N; // K slabs of X in flight per producer thread (registers); -DT3_PREFETCH_N=3 for an
                                           // A/B build loaded through RQB200_LIB (tools/r2_first_call.sh)
constexpr int T3_NF4 = 4;                  // float4 per producer thread per slab: 2 rows x 2 K groups

// D[tmem] (+)= A[tmem] * B[smem desc]^T, issued for the CTA pair
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

// 16 lanes x 16 columns: v[4b + 0/1] = (row lane/4, columns 8b + 2(lane%4) + 0/1), v[4b + 2/3] = same columns of row lane/4 + 8
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// address of the 16-byte chunk `chunk` (0..63) of row `r` (0..31) in a warp's drain buffer
__device__ __forceinline__ unsigned char *drain_at(unsigned char *drain, int r, int chunk) {
    return drain + r * T3_DRAIN_ROW + ((chunk ^ (r & 7)) << 4);
}

// Accumulator rows of this warp's lane quarter, columns c0 .. c0 + T3_EPI_COLS - 1 → the quarter's drain buffer; lane = row.
__device__ __forceinline__ void drain_accumulator(uint32_t taddr, unsigned char *drain, int lane, int c0) {
#pragma unroll 1
    for (int c = c0; c < c0 + T3_EPI_COLS; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4 *>(drain_at(drain, lane, (c >> 2) + j)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}

// Second half of epilogue_rows (tc_common.cuh) reading the drained accumulator: fp32 rows, 128 contiguous bytes per row.
__device__ __forceinline__ void store_rows_from_drain(const unsigned char *drain, int lane, const float *__restrict__ bias,
                                                      float inv_scale, int relu, float *__restrict__ Y, int64_t row0, int64_t n,
                                                      int c0) {
    const int sub = lane >> 3, q = lane & 7;
#pragma unroll 1
    for (int c = c0; c < c0 + T3_EPI_COLS; c += 32) {
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + c) + q);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + sub;
            const float4 a = *reinterpret_cast<const float4 *>(drain_at(const_cast<unsigned char *>(drain), r, (c >> 2) + q));
            float4 o;
            o.x = fmaf(a.x, inv_scale, b4.x); o.y = fmaf(a.y, inv_scale, b4.y);
            o.z = fmaf(a.z, inv_scale, b4.z); o.w = fmaf(a.w, inv_scale, b4.w);
            if (relu) {
                o.x = (o.x != o.x) ? o.x : fmaxf(o.x, 0.0f); o.y = (o.y != o.y) ? o.y : fmaxf(o.y, 0.0f);
                o.z = (o.z != o.z) ? o.z : fmaxf(o.z, 0.0f); o.w = (o.w != o.w) ? o.w : fmaxf(o.w, 0.0f);
            }
            if (row0 + r < n) *reinterpret_cast<float4 *>(Y + (row0 + r) * (int64_t)N3 + c + 4 * q) = o;
        }
    }
}

// Second half of epilogue_rows_split (tc_common.cuh): the activation as split-fp16 UMMA tiles for mlp23_tc_kernel.
__device__ __forceinline__ void store_split_from_drain(const unsigned char *drain, int lane, const float *__restrict__ bias,
                                                       float inv_scale, int relu, unsigned char *__restrict__ tiled, int64_t tile,
                                                       int row_in_tile0, int c0) {
    const int rsub = lane >> 2, ch = lane & 3;
#pragma unroll 1
    for (int c = c0; c < c0 + T3_EPI_COLS; c += 32) {
        const float4 ba = __ldg(reinterpret_cast<const float4 *>(bias + c + 8 * ch));
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + c + 8 * ch + 4));
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        unsigned char *hi_t = tiled + ((size_t)tile * (N3 / 64) + (c >> 6)) * 32768;
        const int chunk = ((c & 63) >> 3) + ch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + rsub;
            const float4 a0 = *reinterpret_cast<const float4 *>(drain_at(const_cast<unsigned char *>(drain), rr, (c >> 2) + 2 * ch));
            const float4 a1 = *reinterpret_cast<const float4 *>(drain_at(const_cast<unsigned char *>(drain), rr, (c >> 2) + 2 * ch + 1));
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
    }
}

template <int PASSES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
linear_tc3_kernel(const float *__restrict__ X, int64_t n, int K, const unsigned char *__restrict__ Wp2,
                  const float *__restrict__ bias, float inv_scale, int relu, float *__restrict__ Y, int tiled_out, int dbg) {
    // dbg: ablation switches of tools/ablate_tc3.py (0 in production; same meaning as in linear_tc2_kernel) — bit0 no
    // epilogue conversion / global stores (the accumulator is still drained), bit1 no MMA, bit2 no tensor-memory stores by
    // the producers, bit3 no W bulk loads, bit4 no X loads
    constexpr int T3_STAGE = T3Cfg<PASSES>::STAGE, T3_STAGES = T3Cfg<PASSES>::STAGES, T3_A_STAGE_COLS = T3Cfg<PASSES>::A_STAGE_COLS;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *drain_all = smem + T3_RING_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(drain_all + 4 * T3_DRAIN_WARP);
    uint64_t *ready = bars;                              // [STAGES] 16 producer warps + the W bulk copy (tx) of this CTA
    uint64_t *peer_ready = bars + T3_STAGES;             // [STAGES] (used in the leader)
    uint64_t *empty = bars + 2 * T3_STAGES;              // [STAGES] commit multicast: W stage and A ring stage are free
    uint64_t *tmem_full = bars + 3 * T3_STAGES;          // [1]
    uint64_t *tmem_empty = tmem_full + 1;                // [1] (used in the leader)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int KS = (K + BK3 - 1) / BK3;
    const int64_t npt = (n + 2 * TM3 - 1) / (2 * TM3);   // pair tiles of 256 rows
    const int64_t pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T3_STAGES; ++s) {
            mbar_init(&ready[s], T3_CONV + 1);
            mbar_init(&peer_ready[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * T3_EPI * 32);
        fence_barrier_init();
    }
    if (warp == T3_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)T3_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < T3_EPI) {
        // ===================== epilogue (each CTA drains its own 128 rows) =====================
        const int quarter = warp & 3, c0 = (warp >> 2) * T3_EPI_COLS;      // lane quarter, first accumulator column
        unsigned char *drain = drain_all + quarter * T3_DRAIN_WARP;      // (two warps of a quarter use disjoint columns of it)
        int64_t it = 0;
        for (int64_t pt = pair0; pt < npt; pt += npairs, ++it) {
            mbar_wait(tmem_full, (uint32_t)(it & 1));
            tc_fence_after();
            drain_accumulator(tmem_base + ((uint32_t)(quarter * 32) << 16), drain, lane, c0);
            tc_fence_before();
            if (rank == 0) mbar_arrive(tmem_empty);
            else mbar_arrive_remote(tmem_empty, 0);
            __syncwarp();                                // the warp's rows are all in its drain buffer
            if (dbg & 1) continue;
            if (tiled_out)
                store_split_from_drain(drain, lane, bias, inv_scale, relu, reinterpret_cast<unsigned char *>(Y), pt * 2 + rank,
                                       quarter * 32, c0);
            else
                store_rows_from_drain(drain, lane, bias, inv_scale, relu, Y, pt * (2 * TM3) + rank * TM3 + quarter * 32, n, c0);
            __syncwarp();                                // before the next tile overwrites the buffer
        }
    } else if (warp < T3_EPI + T3_CONV) {
        // ===================== A producers (own 128 rows) → tensor memory =====================
        // Warp pw: tensor-memory lane quarter pw % 4 (the hardware lets a warp touch lanes 32 (warp id % 4) ..+31; the
        // producers are warps 4..19, so warp id % 4 == pw % 4), 16-row block (pw / 4) % 2 of it, K half pw / 8 of the slab.
        const int pw = warp - T3_EPI;
        const int quarter = pw & 3, rowblk = (pw >> 2) & 1, khalf = pw >> 3;
        const int r_lo = 32 * quarter + 16 * rowblk + (lane >> 2);          // this lane's rows: r_lo and r_lo + 8
        const int kq = 32 * khalf + 4 * (lane & 3);                         // + 16 g for K group g = 0, 1
        const uint32_t a_lane = (uint32_t)(32 * quarter + 16 * rowblk) << 16;
        const int64_t my_tiles = pair0 < npt ? (npt - pair0 + npairs - 1) / npairs : 0;
        const int64_t steps = my_tiles * KS;
        auto load_slab = [&](int64_t st, float4 (&dst)[T3_NF4]) {
            const int64_t pt = pair0 + (st / KS) * npairs;
            const int k0 = (int)(st % KS) * BK3 + kq;
            const int64_t row0 = pt * (2 * TM3) + rank * TM3 + r_lo;
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t row = row0 + 8 * h;
                    const int k = k0 + 16 * g;
                    if (row < n && k < K && !(dbg & 16)) dst[2 * g + h] = __ldg(reinterpret_cast<const float4 *>(X + row * (int64_t)K + k));
                    else dst[2 * g + h] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
        };
        int stage = 0;
        uint32_t phase = 0;
        auto convert_slab = [&](const float4 (&src)[T3_NF4]) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (PASSES == 1) {                  // hi halves only: plain round-to-nearest fp16 pairs (as linear_tc2_kernel<1>)
                    const __half2 p0 = __floats2half2_rn(src[2 * g].x, src[2 * g].y), p1 = __floats2half2_rn(src[2 * g].z, src[2 * g].w);
                    const __half2 p2 = __floats2half2_rn(src[2 * g + 1].x, src[2 * g + 1].y),
                                  p3 = __floats2half2_rn(src[2 * g + 1].z, src[2 * g + 1].w);
                    hi[4 * g + 0] = *reinterpret_cast<const uint32_t *>(&p0); hi[4 * g + 1] = *reinterpret_cast<const uint32_t *>(&p1);
                    hi[4 * g + 2] = *reinterpret_cast<const uint32_t *>(&p2); hi[4 * g + 3] = *reinterpret_cast<const uint32_t *>(&p3);
                    lo[4 * g + 0] = lo[4 * g + 1] = lo[4 * g + 2] = lo[4 * g + 3] = 0u;
                } else {
                    split2(src[2 * g].x, src[2 * g].y, hi[4 * g + 0], lo[4 * g + 0]);             // row r_lo,     k .. k+1
                    split2(src[2 * g].z, src[2 * g].w, hi[4 * g + 1], lo[4 * g + 1]);             // row r_lo,     k+2 .. k+3
                    split2(src[2 * g + 1].x, src[2 * g + 1].y, hi[4 * g + 2], lo[4 * g + 2]);     // row r_lo + 8, k .. k+1
                    split2(src[2 * g + 1].z, src[2 * g + 1].w, hi[4 * g + 3], lo[4 * g + 3]);     // row r_lo + 8, k+2 .. k+3
                }
            }
            mbar_wait(&empty[stage], phase ^ 1);
            tc_fence_after();
            const uint32_t a_hi = tmem_base + a_lane + (uint32_t)(T3_A_COL0 + stage * T3_A_STAGE_COLS + 16 * khalf);
            if (!(dbg & 4)) {
                tmem_st_16x256b_x2(a_hi, hi);
                if (PASSES != 1) tmem_st_16x256b_x2(a_hi + T3_A_COLS, lo);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[stage]);
            if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
        };
        float4 buf[T3_PREFETCH][T3_NF4];
#pragma unroll
        for (int d = 0; d < T3_PREFETCH; ++d)
            if (d < steps) load_slab(d, buf[d]);
        for (int64_t st = 0; st < steps; st += T3_PREFETCH) {
#pragma unroll
            for (int d = 0; d < T3_PREFETCH; ++d) {
                if (st + d < steps) {
                    convert_slab(buf[d]);
                    if (st + d + T3_PREFETCH < steps) load_slab(st + d + T3_PREFETCH, buf[d]);
                }
            }
        }
    } else if (warp == T3_MMA_WARP) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if (rank == 0) {
                // ===================== MMA issuer (leader) =====================
                const uint32_t idesc = umma_idesc(2 * TM3, N3);
                const uint32_t d_tmem = tmem_base;
                int64_t it = 0;
                for (int64_t pt = pair0; pt < npt; pt += npairs, ++it) {
                    mbar_wait_cluster(tmem_empty, (uint32_t)((it & 1) ^ 1));
                    tc_fence_after();
                    for (int slab = 0; slab < KS; ++slab) {
                        mbar_wait(&ready[stage], phase);
                        mbar_wait_cluster(&peer_ready[stage], phase);
                        tc_fence_after();
                        const uint32_t w_hi = smem_u32(smem + stage * T3_STAGE);
                        const uint32_t w_lo = w_hi + T3_W_TILE;
                        const uint32_t a_hi = tmem_base + (uint32_t)(T3_A_COL0 + stage * T3_A_STAGE_COLS);
                        const uint32_t a_lo = a_hi + T3_A_COLS;
                        if (!(dbg & 2))
#pragma unroll
                        for (int kk = 0; kk < BK3 / 16; ++kk) {
                            const uint32_t ko = kk * 32;             // bytes along K in the SWIZZLE_128B W tile
                            const uint32_t ac = kk * 8;              // tensor-memory columns along K (two fp16 each)
                            if (PASSES == 1) {
                                umma_f16_2cta_ts(d_tmem, a_hi + ac, umma_desc(w_hi + ko), idesc, (slab | kk) != 0);
                            } else {
                                umma_f16_2cta_ts(d_tmem, a_lo + ac, umma_desc(w_hi + ko), idesc, (slab | kk) != 0);
                                umma_f16_2cta_ts(d_tmem, a_hi + ac, umma_desc(w_lo + ko), idesc, 1);
                                umma_f16_2cta_ts(d_tmem, a_hi + ac, umma_desc(w_hi + ko), idesc, 1);
                            }
                        }
                        umma_commit_2cta(&empty[stage]);
                        if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_2cta(tmem_full);
                }
            } else {
                // ===================== relay (peer): forward "stage ready" to the leader =====================
                for (int64_t pt = pair0; pt < npt; pt += npairs) {
                    for (int slab = 0; slab < KS; ++slab) {
                        mbar_wait(&ready[stage], phase);
                        tc_fence_after();
                        tc_fence_before();
                        mbar_arrive_remote(&peer_ready[stage], 0);
                        if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ===================== W loader (this CTA's half of the output features) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            constexpr uint32_t half_bytes = 2 * T3_W_TILE;              // hi | lo of this CTA's 128 features (packed image)
            constexpr uint32_t copy_bytes = T3Cfg<PASSES>::NSPLIT * T3_W_TILE;   // one pass needs the hi tile only
            for (int64_t pt = pair0; pt < npt; pt += npairs) {
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (dbg & 8) {
                        mbar_arrive(&ready[stage]);
                    } else {
                        mbar_arrive_expect_tx(&ready[stage], copy_bytes);
                        bulk_g2s(smem + stage * T3_STAGE, Wp2 + ((size_t)slab * 2 + rank) * half_bytes, copy_bytes, &ready[stage]);
                    }
                    if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    }
    // ---- teardown: neither CTA may leave while the pair can still touch its smem / TMEM
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == T3_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)T3_TMEM_COLS));
    }
}

}  // namespace

// RQB200_TC3=1 routes the plain (no gather) first-layer launches of linear_tc2 through this kernel.
// rqb200_debug_tc_flags(4096) does the same inside a running process (tools/check_tc3.py compares the two kernels).
int tc_debug_flags();     // encode_tc.cu

bool linear_tc3_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("RQB200_TC3"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1 || (tc_debug_flags() & 4096) != 0;
}

template <int PASSES>
static int launch_tc3(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, bool tiled_out) {
    auto kern = linear_tc3_kernel<PASSES>;
    static rqb::DeviceOnce attr_once;          // per template instantiation
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM));
    }
    const int64_t npt = (n + 2 * TM3 - 1) / (2 * TM3);
    const int64_t pairs = npt < kNumSMs / 2 ? npt : kNumSMs / 2;
    count_launch();
    kern<<<(unsigned)(pairs * 2), T3_THREADS, T3_SMEM, s>>>(x, n, l.in, (const unsigned char *)l.W_tc2, l.b,
                                                           ldexpf(1.0f, -l.tc_scale_exp), relu ? 1 : 0, y, tiled_out ? 1 : 0,
                                                           tc_debug_flags() & 31);
    RQB_LAUNCH_CHECK();
    return 0;
}

// Same contract as linear_tc2(l, x, n, y, relu, s, passes, nullptr, nullptr, tiled_out) with passes = 3 (split-fp16) or 1
// (hi halves only: the screening tier); l.W_tc2 (the packed W image of encode_tc2.cu) must exist.
int linear_tc3(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, bool tiled_out, int passes) {
    if (n == 0) return 0;
    RQB_CHECK(l.W_tc2 != nullptr && l.out == N3 && l.in % 8 == 0, "linear_tc3: layer not packed for the CTA-pair kernel");
    return passes == 1 ? launch_tc3<1>(l, x, n, y, relu, s, tiled_out) : launch_tc3<3>(l, x, n, y, relu, s, tiled_out);
}

}  // namespace rqb

// quantize_tc.cu — residual quantizer on the tensor cores, with the margin gate fused in.
//
// Fast-route twin of quantize.cu for `ResidualVectorQuantizer.forward(use_sk=False)` (reference
// RQ-VAE/models/rq.py:39-56, RQ-VAE/models/vq.py:63-99): per level the codebook-distance contraction
// r·Cᵀ (vq.py:73) runs as split-fp16 tcgen05 MMAs (codebook chunk staged in shared memory by
// cp.async.bulk, residual tile written by the row warps), the accumulator is read back from TMEM and
// reduced to the top-2 distances per row, the chosen code is gathered and the residual is updated in
// registers (vq.py:95, rq.py:47) — the residual never leaves the SM between levels.
//
// Results here are APPROXIMATE by construction (the input z~ comes from the tensor-core encoder), so
// the kernel also decides which rows can be trusted: a row is appended to the rescue list when, at any
// level, the gap between its two smallest distances is within the error bound of the approximate
// arithmetic (see the derivation at `tau` below).  Rows on the list are recomputed by the exact
// kernels; all other rows carry the reference's codes (the bound is calibrated and checked at scale, not proven: DESIGN.md §4).
//
// CTA = 576 threads: four independent "row groups" of four warps each (one 128-row tile per group, thread
// = row), warp 16 issues the MMAs for all groups, warp 17 streams codebook chunks.  While one group is
// in its top-2 epilogue the tensor core works for the others.
//
// Diagnostic builds (tools/build_variant.sh … quantize_tc.cu -D…): RQB_QTC_TRACE records clock stamps of one steady-state
// batch (tools/trace_qtc.py); RQB_QTC_NO_ALU / RQB_QTC_NO_LD compile the scan's arithmetic / its tensor-memory reads out
// (timing only, wrong codes) — the measurements behind "what bounds this kernel" in DESIGN.md §4.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rqb {

namespace {

constexpr int QTM = 128;                   // rows per group tile
constexpr int QCH = 128;                   // codes per MMA (UMMA N)
constexpr int QB = 1;                      // TMEM accumulator buffers per group (QG * QB * QCH <= 512 columns).  Measured on C2/C3/C5
                                           // (ms per 1M rows): 4 groups x 1 x 128 codes 0.43/0.83/2.7 (this), 4 x 2 x 64 0.50/0.95/3.2
                                           // (N=64 MMAs are shared-memory-bound), 2 x 2 x 128 0.51/0.88/2.8, 2 x 1 x 256 0.47/1.05/3.0
constexpr int QG = 4;                      // row groups per CTA (one 128-row tile each, thread = row)
// (Measured and dropped in round 2: the hi half of the residual tile additionally kept in tensor memory as the A operand of two
// of the three MMA passes, three row groups — bit-identical, 2.61 vs 2.67 ms on the 4 x 1024-code shape, no gain at 3 x 256:
// profiles/r2_first_call_summary.txt.)
static_assert(QG * QB * QCH <= 512, "tensor memory: accumulators");
constexpr int Q_MMA_WARP = 4 * QG, Q_LOAD_WARP = 4 * QG + 1;
constexpr int QTC_THREADS = (4 * QG + 2) * 32;              // 576
constexpr int Q_STAGES = 2;
constexpr int Q_A_BYTES = QTM * 128;       // one of hi / lo, one 64-wide K slab (16 KB)
constexpr int Q_CB_TILE = QCH * 128;       // one of hi / lo (16 KB)
constexpr int Q_STAGE_BYTES = 2 * Q_CB_TILE;               // hi | lo
constexpr int Q_CC_MAX = 4096;             // scaled code norms of all levels kept in shared memory when they fit (16 KB)
constexpr int Q_SMEM = QG * 2 * Q_A_BYTES + Q_STAGES * Q_STAGE_BYTES + 1024 + 256 + Q_CC_MAX * 4;

struct QtcArgs {
    const unsigned char *cbp[RQB200_MAX_LEVELS];   // packed chunks of QCH codes: [hi tile | lo tile]
    const float *ccs[RQB200_MAX_LEVELS];           // |c_j|^2 2^s per code, padded to QCH per chunk with +huge
    uint32_t aug_half[RQB200_MAX_LEVELS];          // fp16 bits of 2^t: value of the augmented A column (see pack_codebook_kernel)
    const float *cb[RQB200_MAX_LEVELS];            // fp32 codebooks for the gather
    const float *cc[RQB200_MAX_LEVELS];
    int K[RQB200_MAX_LEVELS];
    float inv_scale[RQB200_MAX_LEVELS];
    int L;
    uint32_t keep_mask;                            // 0xFFFFFFF0 (see pack_col)
};

#ifdef RQB_QTC_TRACE
// per-phase clock stamps of the SECOND batch of CTA 0 (tools/trace_qtc.py): [group 0..QG-1 | QG = MMA lane][event]
__device__ long long g_qtc_trace[QG + 1][512];
__device__ int g_qtc_trace_n[QG + 1];
#define QTRACE(who, cond) do { if ((cond) && blockIdx.x == 0 && trace_on) { int i_ = g_qtc_trace_n[who]; if (i_ < 512) { g_qtc_trace[who][i_] = clock64(); g_qtc_trace_n[who] = i_ + 1; } } } while (0)
#else
#define QTRACE(who, cond) do { } while (0)
#endif

// (x & keep) | col: with the mask in a register (opaque to the compiler) the two operations fuse into ONE LOP3 with the column
// as its immediate; with two literals they stay two instructions
__device__ __forceinline__ float pack_col(float x, uint32_t keep, uint32_t col) {
    return __uint_as_float((__float_as_uint(x) & keep) | col);
}

__device__ __forceinline__ int q_swz(int rloc, int c) {          // byte offset of 16-byte chunk c of row rloc (SW128 K-major tile)
    return (rloc >> 3) * 1024 + (rloc & 7) * 128 + ((c ^ (rloc & 7)) << 4);
}

// CTA = 576 threads: warps 0-15 are four independent "row groups" (one 128-row tile each, thread = row), warp 16
// issues the MMAs for all groups, warp 17 streams codebook chunks.  While a group scans its distances or updates
// its residual (the next level needs the argmin of this one) the tensor core works for the other groups.  The residual of a row lives ONLY in the group's A tile in shared memory, as the
// split-fp16 pair the MMA reads (r ~ hi + lo to 2^-22, far inside the gate's error budget): no register copy, so
// sixteen row warps fit and e_dim 64 does not spill.
template <int E>
__global__ void __launch_bounds__(QTC_THREADS, 1)
quantize_tc_kernel(const float *__restrict__ z, int64_t n, QtcArgs qa, int64_t *__restrict__ codes,
                   int64_t *__restrict__ list, unsigned long long *__restrict__ list_count, float gate_gamma,
                   float gate_floor, const int64_t *__restrict__ rows, const unsigned long long *__restrict__ n_dev) {
    static_assert(E % 8 == 0 && E <= 64, "tensor-core quantizer supports e_dim <= 64");
    constexpr bool AUG = E < 64;           // the 64-wide K slab has a free column: the MMA itself adds the code norm (adding it in
                                           // the scan instead, three MMA k-steps -> two, measured slower at e = 32: 0.456 vs 0.411 ms)
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *a_base = smem;                                     // [group][hi|lo] 16 KB each
    unsigned char *cb_base = smem + QG * 2 * Q_A_BYTES;               // [stage] hi | lo
    uint64_t *bars = reinterpret_cast<uint64_t *>(cb_base + Q_STAGES * Q_STAGE_BYTES);
    uint64_t *a_full = bars;                       // [QG]     row warps → MMA   (count 4)
    uint64_t *d_full = bars + QG;                  // [QG][QB]  MMA → row warps   (count 1, tcgen05.commit)
    uint64_t *d_empty = bars + (1 + QB) * QG;      // [QG][QB]  row warps → MMA   (count 4: one arrival per warp)
    uint64_t *cb_full = bars + (1 + 2 * QB) * QG;  // [Q_STAGES] loader → MMA (tx)
    uint64_t *cb_empty = cb_full + Q_STAGES;       // [Q_STAGES] MMA → loader (commit)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(cb_empty + Q_STAGES);
    float *cc_s = reinterpret_cast<float *>(cb_base + Q_STAGES * Q_STAGE_BYTES + 256);      // per level, padded to QCH: |c_j|^2 2^s
    int cc_total = 0;
    for (int l = 0; l < qa.L; ++l) cc_total += (qa.K[l] + QCH - 1) / QCH * QCH;
    const bool cc_in_smem = cc_total <= Q_CC_MAX;
    if (cc_in_smem) {
        int off = 0;
        for (int l = 0; l < qa.L; ++l) {
            const int kp = (qa.K[l] + QCH - 1) / QCH * QCH;
            for (int i = threadIdx.x; i < kp; i += QTC_THREADS) cc_s[off + i] = qa.ccs[l][i];
            off += kp;
        }
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (n + QTM - 1) / QTM;
    const int64_t nbatches = (ntiles + QG - 1) / QG;

    if (threadIdx.x == 0) {
        for (int g = 0; g < QG; ++g) {
            mbar_init(&a_full[g], 4);
            for (int b = 0; b < QB; ++b) { mbar_init(&d_full[QB * g + b], 1); mbar_init(&d_empty[QB * g + b], 4); }
        }
        for (int s = 0; s < Q_STAGES; ++s) { mbar_init(&cb_full[s], 1); mbar_init(&cb_empty[s], 1); }
        fence_barrier_init();
    }
    // zero the A tiles once: columns >= E (+ the augmented one) of the 64-wide K slab stay zero for the whole kernel
    for (int i = threadIdx.x; i < QG * 2 * Q_A_BYTES / 16; i += QTC_THREADS) reinterpret_cast<uint4 *>(a_base)[i] = make_uint4(0, 0, 0, 0);
    if (warp == Q_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4 * QG) {
        // ===================== row groups =====================
        const int g = warp >> 2;
        const int rloc = (warp & 3) * 32 + lane;
        unsigned char *a_hi = a_base + g * 2 * Q_A_BYTES;
        unsigned char *a_lo = a_hi + Q_A_BYTES;
        const uint32_t t_group = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * QB * QCH);
        uint32_t round = 0;                       // chunks consumed by this group so far: buffer = round % QB
        for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
            const int64_t tile = batch * QG + g;
            if (tile >= ntiles) continue;         // idle group in the last batch (the MMA warp skips it too)
#ifdef RQB_QTC_TRACE
            const bool trace_on = batch == (int64_t)blockIdx.x + gridDim.x;
            const bool tr = (warp & 3) == 0 && lane == 0;
#endif
            QTRACE(g, tr);
            const int64_t row = tile * QTM + rloc;
            const bool live = row < n;
            const int64_t item = live ? (rows ? __ldg(rows + row) : row) : 0;      // where the codes of this row go
            // latent row → split fp16 → A tile; |z|^2 on the way
            float xx;
            {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int c = 0; c < E / 8; ++c) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (live) {
                        v0 = *reinterpret_cast<const float4 *>(z + row * E + 8 * c);
                        v1 = *reinterpret_cast<const float4 *>(z + row * E + 8 * c + 4);
                    }
                    s0 = fmaf(v0.x, v0.x, s0); s1 = fmaf(v0.y, v0.y, s1); s0 = fmaf(v0.z, v0.z, s0); s1 = fmaf(v0.w, v0.w, s1);
                    s0 = fmaf(v1.x, v1.x, s0); s1 = fmaf(v1.y, v1.y, s1); s0 = fmaf(v1.z, v1.z, s0); s1 = fmaf(v1.w, v1.w, s1);
                    uint4 hi, lo;
                    split2(v0.x, v0.y, hi.x, lo.x); split2(v0.z, v0.w, hi.y, lo.y);
                    split2(v1.x, v1.y, hi.z, lo.z); split2(v1.z, v1.w, hi.w, lo.w);
                    *reinterpret_cast<uint4 *>(a_hi + q_swz(rloc, c)) = hi;
                    *reinterpret_cast<uint4 *>(a_lo + q_swz(rloc, c)) = lo;
                }
                xx = s0 + s1;
            }
            float min_margin = __int_as_float(0x7f800000);
            const uint32_t keep_mask = qa.keep_mask;      // 0xFFFFFFF0 from the host: a literal would be folded back into two instructions
            const float gate_eps = gate_gamma * (sqrtf(xx) + gate_floor);
            int cc_base = 0;                                                   // offset of this level's norms in cc_s
            for (int l = 0; l < qa.L; cc_base += (qa.K[l] + QCH - 1) / QCH * QCH, ++l) {
                if (AUG)   // augmented K column E: A = 2^t, B = |c_j|^2 * 2^(s-t)  ⇒  the MMA itself adds the code norm
                    *reinterpret_cast<uint4 *>(a_hi + q_swz(rloc, E / 8)) = make_uint4(qa.aug_half[l], 0, 0, 0);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[g]);
                QTRACE(g, tr);
                const float inv_s = qa.inv_scale[l];
                // Scan of the accumulator (= 2^s (|c_j|^2 - 2 r.c_j), the distance up to the row constant |r|^2).  The scan is what
                // bounds this kernel (profiles/r2_qtc_phase_trace.txt: with the ALU work compiled out a 128-column chunk takes 355
                // clocks per group, with it 5.5 k), so it is written for the fewest issue slots per distance:
                //   * the column number inside its 16-column block replaces the four low mantissa bits of the value (one LOP3;
                //     relative change < 2^-20 once cleared again, covered by the gate's rounding term), so the arg-min needs no compare + select per
                //     distance — only "which block held the best so far", once per block;
                //   * two distances a time: lo = min(a, b), hi = max(a, b), second = min3(second, max(best, lo), hi),
                //     best = min(best, lo) — 2.5 slots per distance (min3 is one FMNMX3);
                //   * two independent (best, second) trackers (even / odd pairs) keep the dependent chains short.
                float B[2], S[2];
                int blk[2] = {0, 0};
                B[0] = B[1] = S[0] = S[1] = __int_as_float(0x7f800000);
                const int K = qa.K[l];
                for (int c0 = 0; c0 < K; c0 += QCH) {
                    const int ncols = min(QCH, ((K - c0) + 31) & ~31);
                    const uint32_t buf = round % QB;
                    const uint32_t t_addr = t_group + buf * QCH;
                    mbar_wait(&d_full[QB * g + buf], (round / QB) & 1);
                    ++round;
                    tc_fence_after();
                    QTRACE(g, tr);
                    const float *ccs_l = cc_in_smem ? cc_s + cc_base + c0 : qa.ccs[l] + c0;      // scaled norms of this chunk
                    // The accumulator comes out of tensor memory 16 columns at a time, software-pipelined: the read of the next
                    // half block is in flight while this one is scanned.
                    auto scan16 = [&](const uint32_t (&v)[16], const int cc0, const int half) {      // columns cc0 + half .. + 15 of the chunk (half: literal 0 / 16)
#ifdef RQB_QTC_NO_ALU
                        return;
#endif
                        const float b0 = B[0], b1 = B[1];
                        float nn[16];
                        if (!AUG) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {       // same address in every lane
                                const float4 n4 = *reinterpret_cast<const float4 *>(ccs_l + cc0 + half + 4 * q);
                                nn[4 * q] = n4.x; nn[4 * q + 1] = n4.y; nn[4 * q + 2] = n4.z; nn[4 * q + 3] = n4.w;
                            }
                        }
#pragma unroll
                        for (int p = 0; p < 8; ++p) {
                            float x = __uint_as_float(v[2 * p]), y = __uint_as_float(v[2 * p + 1]);
                            if (!AUG) {
                                x += nn[2 * p];
                                y += nn[2 * p + 1];
                            }
                            x = pack_col(x, keep_mask, (uint32_t)(2 * p));
                            y = pack_col(y, keep_mask, (uint32_t)(2 * p + 1));
                            const float lo = fminf(x, y), hi = fmaxf(x, y);
                            const int t = p & 1;
                            S[t] = fminf(S[t], fminf(fmaxf(B[t], lo), hi));
                            B[t] = fminf(B[t], lo);
                        }
                        const int here = c0 + cc0 + half;
                        if (B[0] < b0) blk[0] = here;
                        if (B[1] < b1) blk[1] = here;
                    };
#ifdef RQB_QTC_NO_LD
#define tmem_ld16_async(a, v) do { for (int q_ = 0; q_ < 16; ++q_) v[q_] = (uint32_t)(a) + q_ * 0x3f000u + (uint32_t)xx; } while (0)
#define tmem_ld16_wait(v) do { } while (0)
#endif
                    if (ncols == QCH) {
                        // straight-line, one register window per 16-column read: with a rolled loop (or a conditional read) the
                        // register allocator copies the windows around — one move per distance
                        uint32_t v[QCH / 16][16];
                        tmem_ld16_async(t_addr, v[0]);
#pragma unroll
                        for (int j = 0; j < QCH / 16; ++j) {
                            tmem_ld16_wait(v[j]);
                            if (j + 1 < QCH / 16) tmem_ld16_async(t_addr + (uint32_t)(16 * (j + 1)), v[j + 1]);
                            scan16(v[j], (j >> 1) * 32, (j & 1) * 16);
                        }
                    } else {                            // ragged last chunk of a codebook whose size is no multiple of 128
                        uint32_t va[16];
#pragma unroll 1
                        for (int cc = 0; cc < ncols; cc += 16) {
                            tmem_ld16_async(t_addr + (uint32_t)cc, va);
                            tmem_ld16_wait(va);
                            scan16(va, cc & ~31, cc & 16);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[QB * g + buf]);
                    QTRACE(g, tr);
                }
                // merge the trackers: global best (its low four bits = column inside the block), second = min(other best, both seconds);
                // equal bests leave a zero gap, i.e. the row goes to the next tier
                const bool odd = B[1] < B[0];
                float bestd = odd ? B[1] : B[0];
                float second = fminf(fminf(S[0], S[1]), odd ? B[0] : B[1]);
                int best = (odd ? blk[1] : blk[0]) + (int)(__float_as_uint(bestd) & 15u);
                // back to distance units: d = acc * 2^-s + |r|^2
                const bool bad_index = best >= K;                 // a padded code won: only possible for wild inputs
                if (bad_index) best = 0;
                // the index bits are cleared again before the gap is formed: both values are then truncated the same way (towards
                // zero by less than 2^-20 of their magnitude), so the gap moves by less than 2^-20 (|r|^2 + 2 |c|^2)
                second = fmaf(__uint_as_float(__float_as_uint(second) & 0xFFFFFFF0u), inv_s, xx);
                bestd = fmaf(__uint_as_float(__float_as_uint(bestd) & 0xFFFFFFF0u), inv_s, xx);
                const float ccb = cc_in_smem ? cc_s[cc_base + best] * inv_s : __ldg(qa.cc[l] + best);
                if (live) codes[item * qa.L + l] = best;
                {
                    // Let eps bound |r~ - r| (tensor-core encoder) and rho = |r - c_best|.  A code j can overtake `best`
                    // only if |c_j - c_best| <= 2 rho + 2 eps, and then (d_j - d_best) moves by at most
                    // 2 eps (2 rho + 2 eps); the remaining term covers the rounding of the split-fp16 distance GEMM,
                    // of the fp16-pair residual and of the reference's own fp32 evaluation of d.
                    const float rho = sqrtf(fmaxf(bestd, 0.0f)) + gate_eps;
                    const float tau = 4.0f * gate_eps * (rho + gate_eps) + 6.0e-6f * (xx + fabsf(ccb));    // 4e-6 rounding + 2e-6 for the index bits
                    const float mg = (second - bestd) - tau;
                    min_margin = (mg == mg && min_margin == min_margin && !bad_index) ? fminf(min_margin, mg) : __int_as_float(0x7fc00000);
                }
                QTRACE(g, tr);
                if (l + 1 < qa.L) {
                    // gather + straight-through residual update (vq.py:95 / rq.py:47) on the A tile, 8 dims at a time:
                    // r = hi + lo, xres = r + (q - r), r' = r - xres → split → back into the tile; |r'|^2 for the next level
                    const float4 *q4 = reinterpret_cast<const float4 *>(qa.cb[l] + (int64_t)best * E);
                    float s0 = 0.f, s1 = 0.f;
                    // the chosen code's fp32 row comes from L2: all its loads are issued together, ahead of the first use (the warp
                    // barrier keeps the register-bound schedule from sinking each load to its use, i.e. one L2 round trip per 8 dims)
                    constexpr int HALF = E / 8;                     // 8-dim pieces per batch of loads: the whole row
#pragma unroll
                    for (int h0 = 0; h0 < E / 8; h0 += HALF) {
                        float4 qv[2 * HALF];
#pragma unroll
                        for (int c = 0; c < 2 * HALF; ++c)
                            qv[c] = (h0 + c / 2 < E / 8) ? __ldg(q4 + 2 * h0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                        __syncwarp();
                        QTRACE(g, tr);
#pragma unroll
                        for (int cc = 0; cc < HALF; ++cc) {
                            const int c = h0 + cc;
                            if (c >= E / 8) break;
                            const float4 qa4 = qv[2 * cc], qb4 = qv[2 * cc + 1];
                            const uint4 hi = *reinterpret_cast<const uint4 *>(a_hi + q_swz(rloc, c));
                            const uint4 lo = *reinterpret_cast<const uint4 *>(a_lo + q_swz(rloc, c));
                            const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
                            const float qv8[8] = {qa4.x, qa4.y, qa4.z, qa4.w, qb4.x, qb4.y, qb4.z, qb4.w};
                            float rn[8];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float2 h2 = __half22float2(*reinterpret_cast<const __half2 *>(&hw[t]));
                                const float2 l2 = __half22float2(*reinterpret_cast<const __half2 *>(&lw[t]));
                                const float r0 = h2.x + l2.x, r1 = h2.y + l2.y;
                                const float x0 = __fadd_rn(r0, __fsub_rn(qv8[2 * t], r0)), x1 = __fadd_rn(r1, __fsub_rn(qv8[2 * t + 1], r1));
                                rn[2 * t] = __fsub_rn(r0, x0);
                                rn[2 * t + 1] = __fsub_rn(r1, x1);
                                s0 = fmaf(rn[2 * t], rn[2 * t], s0);
                                s1 = fmaf(rn[2 * t + 1], rn[2 * t + 1], s1);
                            }
                            uint4 nh, nl;
                            split2(rn[0], rn[1], nh.x, nl.x); split2(rn[2], rn[3], nh.y, nl.y);
                            split2(rn[4], rn[5], nh.z, nl.z); split2(rn[6], rn[7], nh.w, nl.w);
                            *reinterpret_cast<uint4 *>(a_hi + q_swz(rloc, c)) = nh;
                            *reinterpret_cast<uint4 *>(a_lo + q_swz(rloc, c)) = nl;
                        }
                    }
                    xx = s0 + s1;
                    QTRACE(g, tr);
                }
            }
            // rows that cannot be certified → rescue list (warp-aggregated append)
            const bool flag = live && !(min_margin > 0.0f);
            const unsigned ballot = __ballot_sync(0xffffffffu, flag);
            if (ballot) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(list_count, (unsigned long long)__popc(ballot));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (flag) list[base + __popc(ballot & ((1u << lane) - 1u))] = item;
            }
        }
    } else if (warp == Q_MMA_WARP) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t a_round[QG], d_round[QG];
            uint64_t da_hi[QG], da_lo[QG];            // the A tiles never move: descriptors built once (+2 per 32-byte K step)
#pragma unroll
            for (int g = 0; g < QG; ++g) {
                a_round[g] = 0; d_round[g] = 0;
                da_hi[g] = umma_desc(smem_u32(a_base + g * 2 * Q_A_BYTES));
                da_lo[g] = umma_desc(smem_u32(a_base + g * 2 * Q_A_BYTES + Q_A_BYTES));
            }
            uint32_t cb_round = 0;
            for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
#ifdef RQB_QTC_TRACE
                const bool trace_on = batch == (int64_t)blockIdx.x + gridDim.x;
#endif
                for (int l = 0; l < qa.L; ++l) {
                    const int K = qa.K[l];
                    for (int c0 = 0; c0 < K; c0 += QCH, ++cb_round) {
                        const int ncols = min(QCH, ((K - c0) + 31) & ~31);
                        const int st = cb_round % Q_STAGES;
                        mbar_wait(&cb_full[st], (cb_round / Q_STAGES) & 1);
                        QTRACE(QG, true);
                        const uint64_t dw_hi = umma_desc(smem_u32(cb_base + st * Q_STAGE_BYTES));
                        const uint64_t dw_lo = umma_desc(smem_u32(cb_base + st * Q_STAGE_BYTES + Q_CB_TILE));
                        const uint32_t idesc = umma_idesc(QTM, ncols);
#pragma unroll
                        for (int g = 0; g < QG; ++g) {
                            if (batch * QG + g >= ntiles) continue;
                            if (c0 == 0) { mbar_wait(&a_full[g], a_round[g] & 1); ++a_round[g]; }
                            const uint32_t buf = d_round[g] % QB;
                            mbar_wait(&d_empty[QB * g + buf], ((d_round[g] / QB) & 1) ^ 1);
                            ++d_round[g];
                            tc_fence_after();
                            QTRACE(QG, true);
                            const uint32_t d_tmem = tmem_base + (uint32_t)((g * QB + buf) * QCH);
#pragma unroll
                            for (int kk = 0; kk < (E + (AUG ? 1 : 0) + 15) / 16; ++kk) {      // E residual columns (+ the augmented norm column)
                                const uint64_t ko = (uint64_t)(2 * kk);                      // 32 bytes along K = 2 descriptor units
                                umma_f16(d_tmem, da_lo[g] + ko, dw_hi + ko, idesc, kk != 0);
                                umma_f16(d_tmem, da_hi[g] + ko, dw_lo + ko, idesc, 1);
                                umma_f16(d_tmem, da_hi[g] + ko, dw_hi + ko, idesc, 1);
                            }
                            umma_commit(&d_full[QB * g + buf]);
                        }
                        umma_commit(&cb_empty[st]);
                    }
                }
            }
        }
    } else {
        // ===================== codebook loader =====================
        if (lane == 0) {
            uint32_t cb_round = 0;
            for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
                for (int l = 0; l < qa.L; ++l) {
                    const int K = qa.K[l];
                    int chunk = 0;
                    for (int c0 = 0; c0 < K; c0 += QCH, ++cb_round, ++chunk) {
                        const int st = cb_round % Q_STAGES;
                        mbar_wait(&cb_empty[st], ((cb_round / Q_STAGES) & 1) ^ 1);
                        mbar_arrive_expect_tx(&cb_full[st], (uint32_t)Q_STAGE_BYTES);
                        bulk_g2s(cb_base + st * Q_STAGE_BYTES, qa.cbp[l] + (size_t)chunk * Q_STAGE_BYTES, (uint32_t)Q_STAGE_BYTES,
                                 &cb_full[st]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == Q_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

}  // namespace
#ifdef RQB_QTC_TRACE
extern "C" __attribute__((visibility("default"))) int rqb200_debug_qtc_trace(long long *host_out, int *n_out, int reset) {
    if (reset) { int z[QG + 1] = {0}; return (int)cudaMemcpyToSymbol(g_qtc_trace_n, z, sizeof(z)); }
    cudaMemcpyFromSymbol(n_out, g_qtc_trace_n, sizeof(int) * (QG + 1));
    return (int)cudaMemcpyFromSymbol(host_out, g_qtc_trace, sizeof(long long) * (QG + 1) * 512);
}
#endif
namespace {
// codebook [K,e] fp32 → per QCH-code chunk: hi tile | lo tile (SW128 K-major, 64-wide K slab).  Column k < e holds
// -2 c_jk 2^s, column e holds |c_j|^2 2^(s-t) (the A operand carries 2^t there), padded codes get a huge norm.
__global__ void pack_codebook_kernel(const float *__restrict__ cb, const float *__restrict__ cc, int K, int e, float scale,
                                     float cc_scale, unsigned char *__restrict__ out, float *__restrict__ ccs) {
    const int nchunks = (K + QCH - 1) / QCH;
    const int64_t total = (int64_t)nchunks * QCH * 8;       // 16-byte units per (hi or lo)
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(u & 7);
        const int rowc = (int)((u >> 3) % QCH);
        const int chunk = (int)(u / (8 * QCH));
        const int code = chunk * QCH + rowc;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float ab[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = c8 * 8 + 2 * t + h;
                float v = 0.0f;
                if (k < e) v = code < K ? -2.0f * cb[(int64_t)code * e + k] * scale : 0.0f;
                else if (k == e) v = code < K ? cc[code] * cc_scale : 60000.0f;
                ab[h] = v;
            }
            split2(ab[0], ab[1], hi[t], lo[t]);
        }
        unsigned char *base = out + (size_t)chunk * Q_STAGE_BYTES;
        const size_t off = (size_t)(rowc >> 3) * 1024 + (rowc & 7) * 128 + ((c8 ^ (rowc & 7)) << 4);
        *reinterpret_cast<uint4 *>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(base + Q_CB_TILE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        if (ccs && c8 == 0) ccs[chunk * QCH + rowc] = code < K ? cc[code] * scale : 3.0e38f;
    }
}

__global__ void absmax2_kernel(const float *__restrict__ w, int64_t count, float *__restrict__ out) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));
}

}  // namespace

// (re)build the packed fp16 images of all codebooks
int quantize_tc_prepare(rqb200_model *m, cudaStream_t s) {
    for (int l = 0; l < m->L; ++l) {
        if (m->cb_tc[l]) continue;
        const int nchunks = (m->K[l] + QCH - 1) / QCH;
        float *dmax = nullptr;
        RQB_CUDA(cudaMalloc(&dmax, 2 * sizeof(float)));
        RQB_CUDA(cudaMemsetAsync(dmax, 0, 2 * sizeof(float), s));
        count_launch();
        absmax2_kernel<<<32, 256, 0, s>>>(m->cb[l], (int64_t)m->K[l] * m->e, dmax);
        count_launch();
        absmax2_kernel<<<8, 256, 0, s>>>(m->cc[l], (int64_t)m->K[l], dmax + 1);
        float hmax[2] = {0.0f, 0.0f};
        RQB_CUDA(cudaMemcpyAsync(hmax, dmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        RQB_CUDA(cudaStreamSynchronize(s));
        RQB_CUDA(cudaFree(dmax));
        RQB_CHECK(hmax[0] == hmax[0] && hmax[0] < 1.0e18f && hmax[1] == hmax[1] && hmax[1] < 1.0e36f, "non-finite codebook entry");
        // s: max|2 c| 2^s in [2^13, 2^14);  t: |c|^2 2^(s-t) < 2^14 with the A column carrying 2^t (t in [0, 15])
        int ex = 0;
        if (hmax[0] > 0.0f) { frexpf(2.0f * hmax[0], &ex); ex = 14 - ex; }
        if (ex > 40) ex = 40;
        if (ex < -20) ex = -20;
        int t = 0;
        if (hmax[1] > 0.0f) { int ec = 0; frexpf(hmax[1], &ec); t = ec + ex - 14; }
        if (t < 0) t = 0;
        RQB_CHECK(t <= 15, "codebook %d: norms too large relative to entries for the fp16 tensor-core quantizer", l);
        m->cb_tc_scale_exp[l] = ex;
        m->cb_tc_aug_exp[l] = t;
        void *p = nullptr;
        RQB_CUDA(cudaMalloc(&p, (size_t)nchunks * Q_STAGE_BYTES));
        if (!m->ccs_tc[l]) RQB_CUDA(cudaMalloc(&m->ccs_tc[l], sizeof(float) * (size_t)nchunks * QCH));
        count_launch();
        pack_codebook_kernel<<<64, 256, 0, s>>>(m->cb[l], m->cc[l], m->K[l], m->e, ldexpf(1.0f, ex), ldexpf(1.0f, ex - t),
                                                (unsigned char *)p, m->ccs_tc[l]);
        RQB_LAUNCH_CHECK();
        m->cb_tc[l] = p;
    }
    return 0;
}

// e_dim 32 / 64: the widths the tensor-core encoder can end in and the kernel is tested with (the template also compiles
// for 16 / 48, but nothing upstream produces such a latent on the fast route, so they are not instantiated)
bool quantize_tc_supported(const rqb200_model *m) { return m->e == 32 || m->e == 64; }

template <int E>
static int launch_qtc(rqb200_model *m, const float *z, int64_t n, int64_t *codes, int64_t *list,
                      unsigned long long *count, cudaStream_t s, float gamma, const int64_t *rows,
                      const unsigned long long *n_dev) {
    auto kern = quantize_tc_kernel<E>;
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM));
    }
    QtcArgs qa;
    qa.L = m->L;
    qa.keep_mask = 0xFFFFFFF0u;
    for (int l = 0; l < RQB200_MAX_LEVELS; ++l) {
        const bool on = l < m->L;
        qa.cbp[l] = on ? (const unsigned char *)m->cb_tc[l] : nullptr;
        qa.cb[l] = on ? m->cb[l] : nullptr;
        qa.cc[l] = on ? m->cc[l] : nullptr;
        qa.K[l] = on ? m->K[l] : 0;
        qa.inv_scale[l] = on ? ldexpf(1.0f, -m->cb_tc_scale_exp[l]) : 0.0f;
        qa.ccs[l] = on ? m->ccs_tc[l] : nullptr;
        qa.aug_half[l] = on ? (uint32_t)__half_as_ushort(__float2half(ldexpf(1.0f, m->cb_tc_aug_exp[l]))) : 0u;
    }
    const int64_t nbatches = ((n + QTM - 1) / QTM + QG - 1) / QG;
    const unsigned grid = (unsigned)(nbatches < kNumSMs ? nbatches : kNumSMs);
    count_launch();
    kern<<<grid, QTC_THREADS, Q_SMEM, s>>>(z, n, qa, codes, list, count, gamma, m->gate_floor, rows, n_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

// z~[n,e] → codes[n,L] (approximate route) + list of rows that need the exact route
int quantize_tc(rqb200_model *m, const float *z, int64_t n, int64_t *codes, int64_t *list, unsigned long long *count,
                cudaStream_t s, float gamma, const int64_t *rows, const unsigned long long *n_dev) {
    if (n == 0) return 0;
    RQB_TRY(quantize_tc_prepare(m, s));
    switch (m->e) {
        case 32: return launch_qtc<32>(m, z, n, codes, list, count, s, gamma, rows, n_dev);
        case 64: return launch_qtc<64>(m, z, n, codes, list, count, s, gamma, rows, n_dev);
    }
    set_error("tensor-core quantizer: e_dim %d not supported", m->e);
    return RQB200_EINVAL;
}

}  // namespace rqb

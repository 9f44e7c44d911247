// quantize.cu — the residual quantizer with use_sk=False, fused over all levels.
//
// Replaces ResidualVectorQuantizer.forward / VectorQuantizer.forward (reference
// RQ-VAE/models/rq.py:39-56, vq.py:63-99): per level distance (vq.py:71-73), first-index argmin
// (vq.py:75), gather (vq.py:87), mse numerators (vq.py:90-92), straight-through output
// (vq.py:95) and residual / accumulate updates (rq.py:47-48) — one kernel, the residual never
// leaves registers between levels, the [n,K] distance matrix is never materialised.
//
// Arithmetic is the reference's, operation by operation:
//   xx   = torch.sum(r**2, dim=1)      → ATen vectorized_inner_sum: 8 lanes x 4 accumulators
//   dot  = one sequential FMA chain over k
//   d    = (xx + cc_j) - (2*dot)       → separate roundings, no contraction
//   idx  = first minimum, NaN wins
//   xres = r + (q - r);  r = r - xres;  x_q = x_q + xres
//
// Mapping: one row per thread (E values in registers), codebook chunk + its norms staged in shared
// memory and read as warp-wide broadcasts (conflict free), four independent chains in flight per
// thread.  Bound: fp32 FMA pipe (L*K*E FMA per row); HBM traffic is 4E+8L bytes per row.
#include <stdlib.h>

#include "common.cuh"

namespace rqb {

namespace {

constexpr int QT = 128;            // threads (= rows) per CTA
constexpr int CHUNK_FLOATS = 16384;   // 64 KB codebook chunk

// torch.sum(v*v) over E contiguous fp32 values, ATen order (see oracle/rqvae_oracle.c).
template <int E>
__device__ __forceinline__ float sumsq_aten(const float (&v)[E]) {
    static_assert(E < 512, "cascade levels of multi_row_sum are not needed below 512");
    constexpr int VEC = E / 8;          // 8-lane vectors
    constexpr int SIZE_ILP = VEC / 4;   // groups of 4 vectors
    float part[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) part[k][l] = 0.0f;
#pragma unroll
    for (int i = 0; i < SIZE_ILP; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                float x = v[i * 32 + k * 8 + l];
                part[k][l] = __fadd_rn(part[k][l], __fmul_rn(x, x));
            }
#pragma unroll
    for (int i = SIZE_ILP * 4; i < VEC; ++i)
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            float x = v[i * 8 + l];
            part[0][l] = __fadd_rn(part[0][l], __fmul_rn(x, x));
        }
#pragma unroll
    for (int k = 1; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) part[0][l] = __fadd_rn(part[0][l], part[k][l]);
    float fin = 0.0f;
#pragma unroll
    for (int k = VEC * 8; k < E; ++k) fin = __fadd_rn(fin, __fmul_rn(v[k], v[k]));
#pragma unroll
    for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, part[0][l]);
    return fin;
}

template <int E>
__global__ void __launch_bounds__(256) sumsq_rows_kernel(const float *__restrict__ v, int n,
                                                         float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r[E];
#pragma unroll
    for (int k = 0; k < E; ++k) r[k] = v[(int64_t)i * E + k];
    out[i] = sumsq_aten<E>(r);
}

struct QuantArgs {
    const float *cb[RQB200_MAX_LEVELS];
    const float *cc[RQB200_MAX_LEVELS];
    int K[RQB200_MAX_LEVELS];
    int L;
};

__device__ __forceinline__ bool better(float d, float best) {
    // torch.argmin: NaN is the minimum; first index wins ties
    return !(best != best) && ((d != d) || d < best);
}

// MODE 0: codes (+ optional xq / sumsq / last residual / margin).  MODE 1: distances of one level.
template <int E, int MODE>
__global__ void __launch_bounds__(QT)
quantize_kernel(const float *__restrict__ z, int64_t n, QuantArgs qa, int64_t *__restrict__ codes,
                const int64_t *__restrict__ rows_out, float *__restrict__ xq_out,
                double *__restrict__ sumsq_out, float *__restrict__ last_residual,
                float *__restrict__ margin_out, float *__restrict__ dist_out, int dist_level,
                float gate_gamma, float gate_floor) {
    extern __shared__ __align__(16) float smem[];
    constexpr int CH = CHUNK_FLOATS / E;    // codes per chunk
    float *s_cb = smem;                     // [CH][E]
    float *s_cc = smem + CH * E;            // [CH]
    __shared__ double s_loss[RQB200_MAX_LEVELS];

    const int tid = threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * QT + tid;
    const bool live = row < n;

    if (tid < RQB200_MAX_LEVELS) s_loss[tid] = 0.0;

    float r[E], xq[E];
#pragma unroll
    for (int k = 0; k < E; k += 4) {
        float4 v = live ? *reinterpret_cast<const float4 *>(z + row * E + k) : make_float4(0, 0, 0, 0);
        r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
    }
    float min_margin = __int_as_float(0x7f800000);
    float gate_eps = 0.0f;      // bound on |z~ - z| for this row (tensor-core route only)

    const int L = (MODE == 1) ? 1 : qa.L;
    for (int l0 = 0; l0 < L; ++l0) {
        const int l = (MODE == 1) ? dist_level : l0;
        const int K = qa.K[l];
        const float *cb = qa.cb[l];
        const float *cc = qa.cc[l];
        if (MODE == 0 && last_residual && l == qa.L - 1 && live) {
#pragma unroll
            for (int k = 0; k < E; k += 4)
                *reinterpret_cast<float4 *>(last_residual + row * E + k) =
                    make_float4(r[k], r[k + 1], r[k + 2], r[k + 3]);
        }
        const float xx = sumsq_aten<E>(r);
        if (MODE == 0 && l0 == 0) gate_eps = gate_gamma * (sqrtf(xx) + gate_floor);
        int best = 0;
        float bestd = 0.0f, second = __int_as_float(0x7f800000);
        for (int c0 = 0; c0 < K; c0 += CH) {
            const int cn = min(CH, K - c0);
            __syncthreads();
            for (int i = tid; i < cn * E / 4; i += QT)
                reinterpret_cast<float4 *>(s_cb)[i] =
                    reinterpret_cast<const float4 *>(cb + (int64_t)c0 * E)[i];
            for (int i = tid; i < cn; i += QT) s_cc[i] = cc[c0 + i];
            __syncthreads();
            int j = 0;
            for (; j + 4 <= cn; j += 4) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                const float4 *c4 = reinterpret_cast<const float4 *>(s_cb + j * E);
#pragma unroll
                for (int k = 0; k < E / 4; ++k) {
                    float4 v0 = c4[k], v1 = c4[k + E / 4], v2 = c4[k + 2 * (E / 4)], v3 = c4[k + 3 * (E / 4)];
                    a0 = __fmaf_rn(r[4 * k], v0.x, a0); a1 = __fmaf_rn(r[4 * k], v1.x, a1);
                    a2 = __fmaf_rn(r[4 * k], v2.x, a2); a3 = __fmaf_rn(r[4 * k], v3.x, a3);
                    a0 = __fmaf_rn(r[4 * k + 1], v0.y, a0); a1 = __fmaf_rn(r[4 * k + 1], v1.y, a1);
                    a2 = __fmaf_rn(r[4 * k + 1], v2.y, a2); a3 = __fmaf_rn(r[4 * k + 1], v3.y, a3);
                    a0 = __fmaf_rn(r[4 * k + 2], v0.z, a0); a1 = __fmaf_rn(r[4 * k + 2], v1.z, a1);
                    a2 = __fmaf_rn(r[4 * k + 2], v2.z, a2); a3 = __fmaf_rn(r[4 * k + 2], v3.z, a3);
                    a0 = __fmaf_rn(r[4 * k + 3], v0.w, a0); a1 = __fmaf_rn(r[4 * k + 3], v1.w, a1);
                    a2 = __fmaf_rn(r[4 * k + 3], v2.w, a2); a3 = __fmaf_rn(r[4 * k + 3], v3.w, a3);
                }
                float dd[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float d = __fsub_rn(__fadd_rn(xx, s_cc[j + t]), __fmul_rn(2.0f, dd[t]));
                    if (MODE == 1) {
                        if (live) dist_out[row * K + c0 + j + t] = d;
                    } else if (c0 + j + t == 0) {
                        bestd = d;
                    } else if (better(d, bestd)) {
                        second = bestd; bestd = d; best = c0 + j + t;
                    } else if (d < second) {
                        second = d;
                    }
                }
            }
            for (; j < cn; ++j) {
                float a0 = 0.f;
#pragma unroll
                for (int k = 0; k < E; ++k) a0 = __fmaf_rn(r[k], s_cb[j * E + k], a0);
                float d = __fsub_rn(__fadd_rn(xx, s_cc[j]), __fmul_rn(2.0f, a0));
                if (MODE == 1) {
                    if (live) dist_out[row * K + c0 + j] = d;
                } else if (c0 + j == 0) {
                    bestd = d;
                } else if (better(d, bestd)) {
                    second = bestd; bestd = d; best = c0 + j;
                } else if (d < second) {
                    second = d;
                }
            }
        }
        if (MODE == 1) return;
        if (live) {
            int64_t orow = rows_out ? rows_out[row] : row;
            codes[orow * qa.L + l] = best;
        }
        {
            // A code j can overtake `best` only if |c_j - c_best| <= 2*rho + 2*eps (rho = |r - c_best|), and then the
            // error of (d_j - d_best) caused by |dr| <= eps is at most 2*eps*(2*rho + 2*eps); plus fp32 rounding of d.
            const float rho = sqrtf(fmaxf(bestd, 0.0f)) + gate_eps;
            const float tau = 4.0f * gate_eps * (rho + gate_eps) + 1.0e-6f * (xx + fabsf(cc[best]));
            const float mg = (second - bestd) - tau;
            min_margin = (mg == mg) ? fminf(min_margin, mg) : mg;      // NaN sticks
            if (min_margin != min_margin) min_margin = __int_as_float(0x7fc00000);
        }
        // gather q, losses, straight-through residual update
        double lsum = 0.0;
        const float4 *q4 = reinterpret_cast<const float4 *>(cb + (int64_t)best * E);
#pragma unroll
        for (int k = 0; k < E; k += 4) {
            float4 q = q4[k / 4];
            float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float diff = __fsub_rn(qv[t], r[k + t]);
                lsum += (double)diff * (double)diff;
                float xres = __fadd_rn(r[k + t], diff);
                r[k + t] = __fsub_rn(r[k + t], xres);
                xq[k + t] = (l == 0) ? __fadd_rn(0.0f, xres) : __fadd_rn(xq[k + t], xres);
            }
        }
        if (sumsq_out) {
            if (!live) lsum = 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
            if ((tid & 31) == 0) atomicAdd(&s_loss[l], lsum);
        }
    }
    if (MODE == 0) {
        if (xq_out && live) {
#pragma unroll
            for (int k = 0; k < E; k += 4)
                *reinterpret_cast<float4 *>(xq_out + row * E + k) =
                    make_float4(xq[k], xq[k + 1], xq[k + 2], xq[k + 3]);
        }
        if (margin_out && live) margin_out[row] = min_margin;
        if (sumsq_out) {
            __syncthreads();
            if (tid < qa.L) atomicAdd(&sumsq_out[tid], s_loss[tid]);
        }
    }
}

QuantArgs make_args(const rqb200_model *m) {
    QuantArgs qa;
    qa.L = m->L;
    for (int l = 0; l < RQB200_MAX_LEVELS; ++l) {
        qa.cb[l] = l < m->L ? m->cb[l] : nullptr;
        qa.cc[l] = l < m->L ? m->cc[l] : nullptr;
        qa.K[l] = l < m->L ? m->K[l] : 0;
    }
    return qa;
}

template <int E, int MODE>
int launch_quant(const rqb200_model *m, const float *z, int64_t n, int64_t *codes,
                 const int64_t *rows_out, float *xq, double *sumsq, float *last_residual,
                 float *margin, float *dist, int dist_level, cudaStream_t s) {
    auto kern = quantize_kernel<E, MODE>;
    constexpr int CH = CHUNK_FLOATS / E;
    size_t smem = sizeof(float) * (CH * E + CH);
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    unsigned grid = (unsigned)((n + QT - 1) / QT);
    rqb::count_launch();
    kern<<<grid, QT, smem, s>>>(z, n, make_args(m), codes, rows_out, xq, sumsq, last_residual, margin,
                               dist, dist_level, m->gate_gamma, m->gate_floor);
    RQB_LAUNCH_CHECK();
    return 0;
}


// ---- few rows (the exact rescue tier of the fast route, collision-group re-encodes): with one thread per row a few
// thousand rows leave most SMs idle and every thread walks L*K*E dependent FMAs.  Here a row is shared by QS
// consecutive lanes; lane s evaluates the codes j = s (mod QS) in ascending order and the lanes merge their
// (distance, index) candidates with the same rule as the sequential scan — NaN first, then the smaller distance, then
// the smaller index — so the chosen code is bit-identical to quantize_kernel's.  Codes only (MODE 0 without the
// optional outputs).  The codebook chunk is staged with a row pitch of E+4 floats: the QS codes a warp reads at once
// then fall into different banks.
constexpr int QS = 8;                      // lanes per row
constexpr int QS_ROWS = QT / QS;           // rows per CTA

template <int E>
__global__ void __launch_bounds__(QT)
quantize_sliced_kernel(const float *__restrict__ z, int64_t n, QuantArgs qa, int64_t *__restrict__ codes,
                       const int64_t *__restrict__ rows_out, const unsigned long long *__restrict__ n_dev) {
    // n_dev (may be NULL): device-resident row count (n = capacity), persistent grid over the row tiles
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    extern __shared__ __align__(16) float smem[];
    constexpr int PITCH = E + 4;
    constexpr int CH = (CHUNK_FLOATS / PITCH) / QS * QS;    // codes per chunk (multiple of QS)
    float *s_cb = smem;                     // [CH][PITCH]
    float *s_cc = smem + CH * PITCH;        // [CH]
    const int tid = threadIdx.x, sl = tid % QS;
    const int64_t n_row_tiles = (n + QS_ROWS - 1) / QS_ROWS;
    for (int64_t row_tile = blockIdx.x; row_tile < n_row_tiles; row_tile += gridDim.x) {
    const int64_t row = row_tile * QS_ROWS + tid / QS;
    const bool live = row < n;
    float r[E];
#pragma unroll
    for (int k = 0; k < E; k += 4) {
        float4 v = live ? *reinterpret_cast<const float4 *>(z + row * E + k) : make_float4(0, 0, 0, 0);
        r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
    }
    for (int l = 0; l < qa.L; ++l) {
        const int K = qa.K[l];
        const float *cb = qa.cb[l];
        const float *cc = qa.cc[l];
        const float xx = sumsq_aten<E>(r);
        int best = 0x7fffffff;              // this lane has not seen a code yet
        float bestd = 0.0f;
        for (int c0 = 0; c0 < K; c0 += CH) {
            const int cn = min(CH, K - c0);
            __syncthreads();
            // asynchronous 16-byte copies (LDGSTS): all of them in flight at once instead of one load-store round trip
            // per iteration — with few rows per CTA the staging would otherwise dominate
            for (int i = tid; i < cn * (E / 4); i += QT) {
                const int j = i / (E / 4), k4 = i % (E / 4);
                const unsigned dst = (unsigned)__cvta_generic_to_shared(s_cb + j * PITCH + 4 * k4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(cb + (int64_t)(c0 + j) * E + 4 * k4) : "memory");
            }
            for (int i = tid; i < cn; i += QT) s_cc[i] = cc[c0 + i];
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            // four codes (j, j+QS, j+2QS, j+3QS) at a time: four independent FMA chains per lane
            int j = sl;
            for (; j + 3 * QS < cn; j += 4 * QS) {
                float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < E / 4; ++k) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float4 v = *reinterpret_cast<const float4 *>(s_cb + (j + t * QS) * PITCH + 4 * k);
                        a[t] = __fmaf_rn(r[4 * k], v.x, a[t]);
                        a[t] = __fmaf_rn(r[4 * k + 1], v.y, a[t]);
                        a[t] = __fmaf_rn(r[4 * k + 2], v.z, a[t]);
                        a[t] = __fmaf_rn(r[4 * k + 3], v.w, a[t]);
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float d = __fsub_rn(__fadd_rn(xx, s_cc[j + t * QS]), __fmul_rn(2.0f, a[t]));
                    if (best == 0x7fffffff || better(d, bestd)) { bestd = d; best = c0 + j + t * QS; }
                }
            }
            for (; j < cn; j += QS) {
                float a0 = 0.f;
                const float4 *c4 = reinterpret_cast<const float4 *>(s_cb + j * PITCH);
#pragma unroll
                for (int k = 0; k < E / 4; ++k) {
                    const float4 v = c4[k];
                    a0 = __fmaf_rn(r[4 * k], v.x, a0);
                    a0 = __fmaf_rn(r[4 * k + 1], v.y, a0);
                    a0 = __fmaf_rn(r[4 * k + 2], v.z, a0);
                    a0 = __fmaf_rn(r[4 * k + 3], v.w, a0);
                }
                const float d = __fsub_rn(__fadd_rn(xx, s_cc[j]), __fmul_rn(2.0f, a0));
                if (best == 0x7fffffff || better(d, bestd)) { bestd = d; best = c0 + j; }
            }
        }
        // merge the QS lanes of the row (xor butterfly inside the aligned group of QS lanes)
#pragma unroll
        for (int o = 1; o < QS; o <<= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bestd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best, o);
            if (oi != 0x7fffffff) {
                bool take;
                if (best == 0x7fffffff) take = true;
                else {
                    const bool mn = bestd != bestd, on = od != od;
                    if (mn || on) take = on && (!mn || oi < best);              // NaN is the minimum; the first NaN wins
                    else take = od < bestd || (od == bestd && oi < best);      // first index wins ties
                }
                if (take) { bestd = od; best = oi; }
            }
        }
        if (live && sl == 0) codes[(rows_out ? rows_out[row] : row) * qa.L + l] = best;
        const float4 *q4 = reinterpret_cast<const float4 *>(cb + (int64_t)best * E);
#pragma unroll
        for (int k = 0; k < E; k += 4) {
            const float4 q = __ldg(q4 + k / 4);
            const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float xres = __fadd_rn(r[k + t], __fsub_rn(qv[t], r[k + t]));
                r[k + t] = __fsub_rn(r[k + t], xres);
            }
        }
    }
    }   // row tiles (every chunk staging starts with a block barrier, so the shared buffers are free again)
}

template <int E>
int launch_quant_sliced(const rqb200_model *m, const float *z, int64_t n, int64_t *codes, const int64_t *rows_out,
                        cudaStream_t s, const unsigned long long *n_dev = nullptr) {
    auto kern = quantize_sliced_kernel<E>;
    constexpr int PITCH = E + 4;
    constexpr int CH = (CHUNK_FLOATS / PITCH) / QS * QS;
    size_t smem = sizeof(float) * (CH * PITCH + CH);
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    unsigned grid = (unsigned)((n + QS_ROWS - 1) / QS_ROWS);
    if (n_dev && grid > (unsigned)kNumSMs * 8) grid = (unsigned)kNumSMs * 8;
    rqb::count_launch();
    kern<<<grid, QT, smem, s>>>(z, n, make_args(m), codes, rows_out, n_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}


// ---- many rows, codes only (the exact rescue tier at catalogue scale: tens of thousands of gated rows, K = 1024):
// one row per thread walks L*K*E dependent FMAs with a handful of warps per SM, and the sliced kernel re-stages the
// whole codebook for every 16 rows.  Here a CTA owns 64 rows as a register-tiled contraction: thread (ty, tx) holds
// 4 rows x 4 codes = 16 independent chains, each the same sequential FMA chain over k as the scan kernels, so every
// distance has the same bits; the 16 code lanes of a row merge their candidates with the sequential rule (NaN first,
// smaller distance, smaller index).  Codebook chunks are double-buffered with cp.async; the residual tile lives in
// shared memory between levels.
constexpr int QTL_ROWS = 64, QTL_THREADS = 256, QTL_CODES = 128;     // rows per CTA, threads, codes per staged chunk

template <int E>
__global__ void __launch_bounds__(QTL_THREADS, 2)
quantize_tiled_kernel(const float *__restrict__ z, int64_t n, QuantArgs qa, int64_t *__restrict__ codes,
                      const int64_t *__restrict__ rows_out, const unsigned long long *__restrict__ n_dev) {
    // n_dev (may be NULL): device-resident row count (n = capacity), persistent grid over the row tiles
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    extern __shared__ __align__(16) float smem[];
    constexpr int PITCH = E + 4;
    float *s_r = smem;                                   // [64][PITCH]  residual tile
    float *s_cb = s_r + QTL_ROWS * PITCH;                // [2][QTL_CODES][PITCH]
    float *s_cc = s_cb + 2 * QTL_CODES * PITCH;          // [2][QTL_CODES]
    float *s_xx = s_cc + 2 * QTL_CODES;                  // [64]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t n_row_tiles = (n + QTL_ROWS - 1) / QTL_ROWS;
    for (int64_t row_tile = blockIdx.x; row_tile < n_row_tiles; row_tile += gridDim.x) {
    if (row_tile != (int64_t)blockIdx.x) __syncthreads();         // the previous tile's shared-memory reads are done
    const int64_t row0 = row_tile * QTL_ROWS;
    for (int i = tid; i < QTL_ROWS * (E / 4); i += QTL_THREADS) {
        const int rr = i / (E / 4), k4 = i % (E / 4);
        const float4 v = row0 + rr < n ? *reinterpret_cast<const float4 *>(z + (row0 + rr) * E + 4 * k4) : make_float4(0, 0, 0, 0);
        *reinterpret_cast<float4 *>(s_r + rr * PITCH + 4 * k4) = v;
    }
    auto stage = [&](const float *cb, const float *cc, int c0, int cn, int buf) {
        float *dst = s_cb + buf * QTL_CODES * PITCH;
        for (int i = tid; i < cn * (E / 4); i += QTL_THREADS) {
            const int j = i / (E / 4), k4 = i % (E / 4);
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + j * PITCH + 4 * k4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(cb + (int64_t)(c0 + j) * E + 4 * k4) : "memory");
        }
        for (int i = tid; i < cn; i += QTL_THREADS) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(s_cc + buf * QTL_CODES + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(cc + c0 + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int l = 0; l < qa.L; ++l) {
        const int K = qa.K[l];
        const float *cb = qa.cb[l];
        const float *cc = qa.cc[l];
        __syncthreads();                                 // residual tile complete (initial load / previous level's update)
        if (tx < 4) {                                    // row norms in the reference's order, one lane per row
            float rv[E];
            const int rr = ty * 4 + tx;
#pragma unroll
            for (int k = 0; k < E; ++k) rv[k] = s_r[rr * PITCH + k];
            s_xx[rr] = sumsq_aten<E>(rv);
        }
        const int nchunks = (K + QTL_CODES - 1) / QTL_CODES;
        stage(cb, cc, 0, min(QTL_CODES, K), 0);
        int best[4];
        float bestd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { best[i] = 0x7fffffff; bestd[i] = 0.0f; }
        float xx[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * QTL_CODES, cn = min(QTL_CODES, K - c0), buf = ch & 1;
            if (ch + 1 < nchunks) {
                stage(cb, cc, c0 + QTL_CODES, min(QTL_CODES, K - c0 - QTL_CODES), buf ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();                             // chunk `buf` (and, the first time, s_xx) visible to everybody
            if (ch == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) xx[i] = s_xx[ty * 4 + i];
            }
            const float *cbuf = s_cb + buf * QTL_CODES * PITCH;
            const float *ccbuf = s_cc + buf * QTL_CODES;
            for (int j0 = 0; j0 < cn; j0 += 64) {        // 64 codes per step: this lane takes j0 + tx + 16 t
                float a[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int t = 0; t < 4; ++t) a[i][t] = 0.0f;
#pragma unroll 4
                for (int k4 = 0; k4 < E / 4; ++k4) {
                    float4 rv[4], cv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) rv[i] = *reinterpret_cast<const float4 *>(s_r + (ty * 4 + i) * PITCH + 4 * k4);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int j = min(j0 + tx + 16 * t, cn - 1);         // clamped lanes are ignored below
                        cv[t] = *reinterpret_cast<const float4 *>(cbuf + j * PITCH + 4 * k4);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            a[i][t] = __fmaf_rn(rv[i].x, cv[t].x, a[i][t]);
                            a[i][t] = __fmaf_rn(rv[i].y, cv[t].y, a[i][t]);
                            a[i][t] = __fmaf_rn(rv[i].z, cv[t].z, a[i][t]);
                            a[i][t] = __fmaf_rn(rv[i].w, cv[t].w, a[i][t]);
                        }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int j = j0 + tx + 16 * t;
                    if (j < cn) {
                        const float ccj = ccbuf[j];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float d = __fsub_rn(__fadd_rn(xx[i], ccj), __fmul_rn(2.0f, a[i][t]));
                            if (best[i] == 0x7fffffff || better(d, bestd[i])) { bestd[i] = d; best[i] = c0 + j; }
                        }
                    }
                }
            }
            __syncthreads();                             // everybody is done with `buf` before it is staged again
        }
        // merge the 16 code lanes of every row (xor butterfly inside the aligned half warp)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bestd[i], o);
                const int oi = __shfl_xor_sync(0xffffffffu, best[i], o);
                if (oi != 0x7fffffff) {
                    bool take;
                    if (best[i] == 0x7fffffff) take = true;
                    else {
                        const bool mn = bestd[i] != bestd[i], on = od != od;
                        if (mn || on) take = on && (!mn || oi < best[i]);
                        else take = od < bestd[i] || (od == bestd[i] && oi < best[i]);
                    }
                    if (take) { bestd[i] = od; best[i] = oi; }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = row0 + ty * 4 + i;
            if (tx == 0 && row < n) codes[(rows_out ? rows_out[row] : row) * qa.L + l] = best[i];
            // straight-through residual update (vq.py:95, rq.py:47): lane tx owns dims 4tx.. (+64, ...) of the row
            for (int k = 4 * tx; k < E; k += 64) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(cb + (int64_t)best[i] * E + k));
                float4 rv = *reinterpret_cast<float4 *>(s_r + (ty * 4 + i) * PITCH + k);
                rv.x = __fsub_rn(rv.x, __fadd_rn(rv.x, __fsub_rn(q.x, rv.x)));
                rv.y = __fsub_rn(rv.y, __fadd_rn(rv.y, __fsub_rn(q.y, rv.y)));
                rv.z = __fsub_rn(rv.z, __fadd_rn(rv.z, __fsub_rn(q.z, rv.z)));
                rv.w = __fsub_rn(rv.w, __fadd_rn(rv.w, __fsub_rn(q.w, rv.w)));
                *reinterpret_cast<float4 *>(s_r + (ty * 4 + i) * PITCH + k) = rv;
            }
        }
    }
    }   // row tiles
}

template <int E>
int launch_quant_tiled(const rqb200_model *m, const float *z, int64_t n, int64_t *codes, const int64_t *rows_out,
                       cudaStream_t s, const unsigned long long *n_dev = nullptr) {
    auto kern = quantize_tiled_kernel<E>;
    constexpr int PITCH = E + 4;
    size_t smem = sizeof(float) * (QTL_ROWS * PITCH + 2 * QTL_CODES * PITCH + 2 * QTL_CODES + QTL_ROWS);
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    unsigned grid = (unsigned)((n + QTL_ROWS - 1) / QTL_ROWS);
    if (n_dev && grid > (unsigned)kNumSMs * 4) grid = (unsigned)kNumSMs * 4;
    rqb::count_launch();
    kern<<<grid, QTL_THREADS, smem, s>>>(z, n, make_args(m), codes, rows_out, n_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

#define RQB_DISPATCH_E(e, CALL)                                                     \
    switch (e) {                                                                    \
        case 8:   { constexpr int E = 8;   CALL; } break;                            \
        case 16:  { constexpr int E = 16;  CALL; } break;                            \
        case 32:  { constexpr int E = 32;  CALL; } break;                            \
        case 48:  { constexpr int E = 48;  CALL; } break;                            \
        case 64:  { constexpr int E = 64;  CALL; } break;                            \
        case 96:  { constexpr int E = 96;  CALL; } break;                            \
        case 128: { constexpr int E = 128; CALL; } break;                            \
        default:                                                                    \
            set_error("e_dim %d not supported (8,16,32,48,64,96,128)", e);          \
            return RQB200_EINVAL;                                                   \
    }

__global__ void recon_error_kernel(const float *__restrict__ out, const float *__restrict__ x,
                                   int64_t count, double *__restrict__ acc) {
    double s2 = 0.0, s1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        double d = (double)out[i] - (double)x[i];
        s2 += d * d;
        s1 += fabs(d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    __shared__ double w2[32], w1[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { w2[wid] = s2; w1[wid] = s1; }
    __syncthreads();
    if (wid == 0) {
        int nw = blockDim.x >> 5;
        s2 = lane < nw ? w2[lane] : 0.0;
        s1 = lane < nw ? w1[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0) { atomicAdd(&acc[0], s2); atomicAdd(&acc[1], s1); }
    }
}

}  // namespace

int codebook_norms(const float *cb, int K, int e, float *cc, cudaStream_t s) {
    count_launch();
    RQB_DISPATCH_E(e, (sumsq_rows_kernel<E><<<(K + 255) / 256, 256, 0, s>>>(cb, K, cc)));
    RQB_LAUNCH_CHECK();
    return 0;
}

int quantize_exact(const rqb200_model *m, const float *z, int64_t n, int64_t *codes,
                   const int64_t *rows_out, float *xq, double *sumsq, float *last_residual,
                   float *margin_out, cudaStream_t s, int64_t batch_rows, const unsigned long long *n_dev, int64_t n_hint) {
    if (n == 0) return 0;
    ProfScope ps(PROF_QUANTIZE, s);
    if (n_dev) {
        // device-counted rows (rescue tier; codes only, rows of a large batch): n is the capacity, n_hint the expected count
        RQB_CHECK(!xq && !sumsq && !last_residual && !margin_out && m->e <= 64, "device-counted quantizer: codes only, e_dim <= 64");
        RQB_CHECK(!small_batch_lane16(batch_rows, m->e), "device-counted rows must belong to a batch of 16 or more rows");
        if (n_hint <= (int64_t)kNumSMs * QT / 2) { RQB_DISPATCH_E(m->e, return (launch_quant_sliced<E>(m, z, n, codes, rows_out, s, n_dev))); }
        else { RQB_DISPATCH_E(m->e, return (launch_quant_tiled<E>(m, z, n, codes, rows_out, s, n_dev))); }
        return 0;
    }
    // a batch of 2..15 rows with 24·n <= e: matmul(latent, E.t()) runs in the reference's small-batch order (small_batch.cu)
    if (small_batch_lane16(batch_rows < 0 ? n : batch_rows, m->e) && !last_residual && !margin_out)
        return quantize_small(m, z, rows_out, nullptr, (int)n, n, m->L, codes, nullptr, xq, sumsq, nullptr, 0, s);
    // few rows and codes only: share a row between QS lanes so that the whole GPU works on it (measured: faster than a row
    // per thread up to about half a wave of 128-row CTAs, slower beyond — the row-per-thread kernel then fills the SMs)
    if (n <= (int64_t)kNumSMs * QT / 2 && !xq && !sumsq && !last_residual && !margin_out && m->e <= 64) {
        RQB_DISPATCH_E(m->e, return (launch_quant_sliced<E>(m, z, n, codes, rows_out, s)));
        return 0;
    }
    // many rows, codes only (rescue tier at catalogue scale): register-tiled kernel, same bits
    static int tiled_mode = -1;            // RQB200_QUANT_TILED=0 keeps the row-per-thread kernel (A/B, tests)
    if (tiled_mode < 0) {
        const char *ev = getenv("RQB200_QUANT_TILED");
        tiled_mode = (ev && ev[0] == '0') ? 0 : 1;
    }
    if (tiled_mode && !xq && !sumsq && !last_residual && !margin_out && m->e <= 128) {
        RQB_DISPATCH_E(m->e, return (launch_quant_tiled<E>(m, z, n, codes, rows_out, s)));
        return 0;
    }
    RQB_DISPATCH_E(m->e, return (launch_quant<E, 0>(m, z, n, codes, rows_out, xq, sumsq, last_residual,
                                                   margin_out, nullptr, 0, s)));
    return 0;
}

int distances_exact(const rqb200_model *m, int level, const float *r, int64_t n, float *d,
                    cudaStream_t s) {
    if (n == 0) return 0;
    RQB_CHECK(level >= 0 && level < m->L, "level %d out of range", level);
    if (small_batch_lane16(n, m->e))
        return quantize_small(m, r, nullptr, nullptr, (int)n, n, 1, nullptr, nullptr, nullptr, nullptr, d, level, s);
    RQB_DISPATCH_E(m->e, return (launch_quant<E, 1>(m, r, n, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                   nullptr, d, level, s)));
    return 0;
}

int recon_error(const float *out, const float *x, int64_t count, double *recon_sum, cudaStream_t s) {
    if (count == 0) return 0;
    int64_t blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    rqb::count_launch();
    recon_error_kernel<<<(unsigned)blocks, 256, 0, s>>>(out, x, count, recon_sum);
    RQB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rqb

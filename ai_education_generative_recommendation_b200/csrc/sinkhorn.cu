// sinkhorn.cu — the use_sk branch of VectorQuantizer.forward and the collision re-encode step.
//
// Replaces (reference RQ-VAE/models/vq.py:51-61,74-83 and RQ-VAE/models/layers.py:85-108):
//     d  = center_distance_for_constraint(d)          fp32: (d - mid) / (max - mid + 1e-5)
//     Q  = sinkhorn_algorithm(d.double(), eps, iters)  fp64: exp(-d/eps), normalise total, then
//          iters x { Q /= rowsum; Q /= B; Q /= colsum; Q /= K };  Q *= B
//     idx = argmax(Q, -1)
// and its only caller on the encode path, the per-group re-encode loop of reference
// RQ-VAE/infer.py:109-130 (Sinkhorn on the LAST level only, one call per collision group).
//
// The reference launches ~200 tiny kernels per group; here one CTA owns one group: the [B,K] fp64
// matrix lives in shared memory, distances are recomputed in the reference's fp32 order from the
// residual entering the last level, and every division of the reference is kept as a separate
// fp64 division so the iteration follows the same trajectory.  Latency/fp64-ALU bound by nature;
// reported as time, not against a roofline.
#include <stdlib.h>

#include "common.cuh"

namespace rqb {

namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_ONE_LEVEL_MAX = 64;    // grid Sinkhorn: up to this many CTAs every CTA sums all column partials itself

// torch.sum(v*v) for a runtime length e < 512 (ATen order; see oracle/rqvae_oracle.c)
__device__ float sumsq_aten_rt(const float *v, int e) {
    const int vec = e / 8, size_ilp = vec / 4;
    float part[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) part[k][l] = 0.0f;
    for (int i = 0; i < size_ilp; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                float x = v[i * 32 + k * 8 + l];
                part[k][l] = __fadd_rn(part[k][l], __fmul_rn(x, x));
            }
    for (int i = size_ilp * 4; i < vec; ++i)
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
    for (int k = vec * 8; k < e; ++k) fin = __fadd_rn(fin, __fmul_rn(v[k], v[k]));
#pragma unroll
    for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, part[0][l]);
    return fin;
}

// <r, c> in the order the reference's matmul(latent, E.t()) (vq.py:73) uses for a batch of B rows: one sequential fma
// chain, or for 2 <= B <= 15 with 24·B <= e the small-batch order of small_batch.cu (16 interleaved chains, folded)
__device__ __forceinline__ float dot_for_batch(const float *r, const float *c, int e, int B) {
    if (!small_batch_lane16(B, e)) {
        float acc = 0.0f;
        for (int k = 0; k < e; ++k) acc = __fmaf_rn(r[k], c[k], acc);
        return acc;
    }
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0f;
    int k = 0;
    for (; k + 16 <= e; k += 16)
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = __fmaf_rn(r[k + q], c[k + q], acc[q]);
#pragma unroll
    for (int q = 0; q < 16; ++q)
        if (k + q < e) acc[q] = __fmaf_rn(r[k + q], c[k + q], acc[q]);
    float s4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) s4[q] = __fadd_rn(__fadd_rn(__fadd_rn(acc[q], acc[q + 4]), acc[q + 8]), acc[q + 12]);
    return __fadd_rn(__fadd_rn(s4[0], s4[1]), __fadd_rn(s4[2], s4[3]));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// a / b for MANY numerators over ONE denominator, without a division per element: y = RN(1/b) once, then
//   q0 = RN(a·y);  r0 = a − b·q0 (exact, fma);  q1 = RN(q0 + r0·y)  → faithful;  r1 = a − b·q1 (exact);  q = RN(q1 + r1·y)
// which is the correctly rounded quotient RN(a/b) (Markstein's final-step theorem for a faithful q1 and a correctly rounded
// reciprocal) as long as nothing in the sequence under- or overflows: a, b, q normal with exponents inside ±960, which makes
// the residuals exactly representable.  The window is checked on the NUMERATOR alone: for a given b the exponent of q is that
// of a minus that of b (or one less), so "a and q inside ±960" is one range test on the high word of a against bounds that
// are computed once per denominator (two integer instructions per quotient; a zero, subnormal, negative, infinite or NaN
// numerator falls outside every window).  Everything outside takes the hardware division.  Same bits as `a / b`, one multiply
// + four fmas instead of a division: the Sinkhorn kernels are bound by exactly these.  rqb200_debug_check_division compares
// the two on 2^32 operand pairs per call (tests/test_gpu_parity.py).
__device__ __noinline__ double div_slow(double a, double b) { return a / b; }

struct DivCtx {
    double b, y;            // denominator and RN(1/b)
    unsigned lo, span;      // numerators with (hi32(a) - lo) < span take the fma sequence
};

__device__ __forceinline__ DivCtx make_div(double b) {
    DivCtx c;
    c.b = b;
    c.y = 1.0 / b;
    // a >= 2^-968 keeps both residuals (multiples of 2^(e_a - 104)) representable; q >= 2^-1021 keeps every quotient normal;
    // |e_b| <= 960 keeps b, 1/b and the products far from the ends of the range
    const unsigned eb = (unsigned)__double2hiint(b) >> 20;         // sign bit included: a negative b is "out of range"
    const bool b_ok = eb - 63u <= 1920u;                           // 2^-960 <= b < 2^961, finite, positive
    const int lo_e = max(55, (int)eb - 1020), hi_e = min(1983, (int)eb + 960);
    c.lo = b_ok ? (unsigned)lo_e << 20 : 0xffffffffu;
    c.span = b_ok ? (unsigned)(hi_e + 1 - lo_e) << 20 : 0u;
    return c;
}

// the fma sequence alone (the caller has checked the window for a whole batch of numerators: straight-line code, so the
// independent quotients of a row or column overlap in the fp64 pipe)
__device__ __forceinline__ double div_fast(double a, const DivCtx &c) {
    const double q0 = a * c.y;
    const double q1 = __fma_rn(__fma_rn(-q0, c.b, a), c.y, q0);
    return __fma_rn(__fma_rn(-q1, c.b, a), c.y, q1);
}
__device__ __forceinline__ bool div_outside(double a, const DivCtx &c) { return (unsigned)__double2hiint(a) - c.lo >= c.span; }

__device__ __forceinline__ double div_ctx(double a, const DivCtx &c) {
    const double q = div_fast(a, c);
    if (__builtin_expect(div_outside(a, c), 0)) return div_slow(a, c.b);   // a real branch
    return q;
}

__device__ double block_sum(double v, double *s_red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < SK_THREADS / 32; ++w) t += s_red[w];
    return t;
}

// Sinkhorn iterations on a [B,K] fp64 matrix owned by this CTA (shared or global memory).
__device__ void sinkhorn_cta(double *Q, int B, int K, int iters, double *s_red) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = SK_THREADS / 32;
    // sum_Q = Q.sum(-1).sum(-2);  Q /= sum_Q
    double part = 0.0;
    for (int i = wid; i < B; i += NW) {
        double rs = 0.0;
        for (int j = lane; j < K; j += 32) rs += Q[(size_t)i * K + j];
        rs = warp_sum(rs);
        if (lane == 0) part += rs;
    }
    double total = block_sum(part, s_red);
    {
        const DivCtx ct = make_div(total);
        for (int i = tid; i < B * K; i += SK_THREADS) Q[i] = div_ctx(Q[i], ct);
    }
    __syncthreads();
    const double dB = (double)B, dK = (double)K;
    // `Q /= B` / `Q /= K` with a power-of-two divisor: x / 2^k and x * 2^-k are the correctly rounded value of the same real
    // number, hence the same bits — one fp64 division less per element and half-iteration (the kernel is division-bound)
    const bool b_pow2 = (B & (B - 1)) == 0, k_pow2 = (K & (K - 1)) == 0;
    const double invB = 1.0 / dB, invK = 1.0 / dK;
    const DivCtx cB = make_div(dB), cK = make_div(dK);
    for (int it = 0; it < iters; ++it) {
        // Q /= Q.sum(dim=1, keepdim=True);  Q /= B
        for (int i = wid; i < B; i += NW) {
            double rs = 0.0;
            for (int j = lane; j < K; j += 32) rs += Q[(size_t)i * K + j];
            rs = warp_sum(rs);
            const DivCtx cr = make_div(rs);
            if (b_pow2) { for (int j = lane; j < K; j += 32) Q[(size_t)i * K + j] = div_ctx(Q[(size_t)i * K + j], cr) * invB; }
            else { for (int j = lane; j < K; j += 32) Q[(size_t)i * K + j] = div_ctx(div_ctx(Q[(size_t)i * K + j], cr), cB); }
        }
        __syncthreads();
        // Q /= Q.sum(dim=0, keepdim=True);  Q /= K
        for (int j = tid; j < K; j += SK_THREADS) {
            double cs = 0.0;
            for (int i = 0; i < B; ++i) cs += Q[(size_t)i * K + j];
            const DivCtx cc_ = make_div(cs);
            if (k_pow2) { for (int i = 0; i < B; ++i) Q[(size_t)i * K + j] = div_ctx(Q[(size_t)i * K + j], cc_) * invK; }
            else { for (int i = 0; i < B; ++i) Q[(size_t)i * K + j] = div_ctx(div_ctx(Q[(size_t)i * K + j], cc_), cK); }
        }
        __syncthreads();
    }
    for (int i = tid; i < B * K; i += SK_THREADS) Q[i] = Q[i] * dB;
    __syncthreads();
}

// argmax over K for each row (first maximum wins; NaN counts as maximum like torch.argmax)
__device__ void argmax_rows(const double *Q, int B, int K, const int64_t *items, int64_t *codes, int L,
                            int64_t *idx_out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = wid; i < B; i += SK_THREADS / 32) {
        double best = 0.0;
        int bj = -1;
        for (int j = lane; j < K; j += 32) {
            double v = Q[(size_t)i * K + j];
            bool take = bj < 0 || (!(best != best) && ((v != v) || v > best));
            if (take) { best = v; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (oj >= 0) {
                bool onan = ov != ov, bnan = best != best;
                bool take = bj < 0 || (onan && (!bnan || oj < bj)) ||
                            (!onan && !bnan && (ov > best || (ov == best && oj < bj)));
                if (take) { best = ov; bj = oj; }
            }
        }
        if (lane == 0) {
            if (codes) codes[items[i] * L + (L - 1)] = bj;
            else idx_out[i] = bj;
        }
    }
}

// centre (fp32) + exp (fp64) in place: Q holds fp32 distances widened to double on entry
__device__ void center_and_exp(double *Q, int count, double epsilon, float *s_mm) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float mx = -__int_as_float(0x7f800000), mn = __int_as_float(0x7f800000);
    for (int i = tid; i < count; i += SK_THREADS) {
        float d = (float)Q[i];
        mx = fmaxf(mx, d);
        mn = fminf(mn, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    __syncthreads();
    if (lane == 0) { s_mm[wid] = mx; s_mm[32 + wid] = mn; }
    __syncthreads();
    for (int w = 0; w < SK_THREADS / 32; ++w) { mx = fmaxf(mx, s_mm[w]); mn = fminf(mn, s_mm[32 + w]); }
    const float middle = __fdiv_rn(__fadd_rn(mx, mn), 2.0f);
    const float amplitude = __fadd_rn(__fsub_rn(mx, middle), 1e-5f);
    for (int i = tid; i < count; i += SK_THREADS) {
        float c = __fdiv_rn(__fsub_rn((float)Q[i], middle), amplitude);
        Q[i] = exp(-((double)c) / epsilon);
    }
    __syncthreads();
}

// one CTA per collision group (infer.py:120-129 for one `collision_items` list)
__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_regroup_kernel(const float *__restrict__ residual, const int64_t *__restrict__ items,
                        const int64_t *__restrict__ offsets, int e, const float *__restrict__ cb,
                        const float *__restrict__ cc, int K, int L, int min_rows, int cap_rows, double epsilon, int iters,
                        int64_t *__restrict__ codes) {
    extern __shared__ __align__(16) unsigned char sk_smem[];
    __shared__ double s_red[SK_THREADS / 32];
    __shared__ float s_mm[64];
    const int64_t g0 = offsets[blockIdx.x], g1 = offsets[blockIdx.x + 1];
    const int64_t B64 = g1 - g0;
    // this launch serves one size class (its shared memory is sized for cap_rows); larger groups are taken by the
    // next class or by sinkhorn_regroup_large_kernel
    if (B64 < 2 || B64 < min_rows || B64 > cap_rows) return;
    const int B = (int)B64;
    double *Q = reinterpret_cast<double *>(sk_smem);                    // [B][K]
    float *s_r = reinterpret_cast<float *>(Q + (size_t)cap_rows * K);   // [B][e]
    float *s_xx = s_r + (size_t)cap_rows * e;                           // [B]
    const int64_t *gi = items + g0;
    const int tid = threadIdx.x;
    for (int i = tid; i < B * e; i += SK_THREADS) s_r[i] = residual[gi[i / e] * e + (i % e)];
    __syncthreads();
    for (int i = tid; i < B; i += SK_THREADS) s_xx[i] = sumsq_aten_rt(s_r + (size_t)i * e, e);
    __syncthreads();
    // d[i][j] = (xx_i + cc_j) - 2 * <r_i, c_j>   (vq.py:71-73)
    for (int p = tid; p < B * K; p += SK_THREADS) {
        const int i = p / K, j = p % K;
        const float *r = s_r + (size_t)i * e;
        const float *c = cb + (size_t)j * e;
        const float acc = dot_for_batch(r, c, e, B);
        Q[p] = (double)__fsub_rn(__fadd_rn(s_xx[i], cc[j]), __fmul_rn(2.0f, acc));
    }
    __syncthreads();
    center_and_exp(Q, B * K, epsilon, s_mm);
    sinkhorn_cta(Q, B, K, iters, s_red);
    argmax_rows(Q, B, K, gi, codes, L, nullptr);
}


// Groups of 2 … SKW_MAX_ROWS rows — the bulk of every round after the first — one WARP per group, no block barriers.
// Bit-identical to sinkhorn_regroup_kernel by construction: every sum runs in the same order (row sums: lane-strided
// partials + xor butterfly; the total: row sums added in row order; column sums: rows in order), every quotient through the
// same div_by, the same centre / exp expressions, the same arg-max rule.
constexpr int SKW_MAX_ROWS = 8;
__global__ void __launch_bounds__(128)
sinkhorn_regroup_warp_kernel(const float *__restrict__ residual, const int64_t *__restrict__ items,
                             const int64_t *__restrict__ offsets, int64_t n_groups, int e, const float *__restrict__ cb,
                             const float *__restrict__ cc, int K, int L, double epsilon, int iters, int warps_per_cta,
                             int64_t *__restrict__ codes) {
    extern __shared__ __align__(16) unsigned char sk_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid >= warps_per_cta) return;
    const int64_t g = (int64_t)blockIdx.x * warps_per_cta + wid;
    if (g >= n_groups) return;
    const int64_t g0 = offsets[g];
    const int64_t B64 = offsets[g + 1] - g0;
    if (B64 < 2 || B64 > SKW_MAX_ROWS) return;
    const int B = (int)B64;
    const size_t per_warp = sizeof(double) * SKW_MAX_ROWS * (size_t)K + sizeof(float) * SKW_MAX_ROWS * (size_t)(e + 1);
    double *Q = reinterpret_cast<double *>(sk_smem + (size_t)wid * per_warp);          // [B][K]
    float *s_r = reinterpret_cast<float *>(Q + (size_t)SKW_MAX_ROWS * K);              // [B][e]
    float *s_xx = s_r + (size_t)SKW_MAX_ROWS * e;                                      // [B]
    const int64_t *gi = items + g0;
    for (int i = lane; i < B * e; i += 32) s_r[i] = residual[gi[i / e] * e + (i % e)];
    __syncwarp();
    if (lane < B) s_xx[lane] = sumsq_aten_rt(s_r + (size_t)lane * e, e);
    __syncwarp();
    float mx = -__int_as_float(0x7f800000), mn = __int_as_float(0x7f800000);
    for (int p = lane; p < B * K; p += 32) {
        const int i = p / K, j = p % K;
        const float acc = dot_for_batch(s_r + (size_t)i * e, cb + (size_t)j * e, e, B);
        const float d = __fsub_rn(__fadd_rn(s_xx[i], cc[j]), __fmul_rn(2.0f, acc));
        Q[p] = (double)d;
        mx = fmaxf(mx, d);
        mn = fminf(mn, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    const float middle = __fdiv_rn(__fadd_rn(mx, mn), 2.0f);
    const float amplitude = __fadd_rn(__fsub_rn(mx, middle), 1e-5f);
    for (int p = lane; p < B * K; p += 32) {
        const float c = __fdiv_rn(__fsub_rn((float)Q[p], middle), amplitude);
        Q[p] = exp(-((double)c) / epsilon);
    }
    __syncwarp();
    // sum_Q = Q.sum(-1).sum(-2);  Q /= sum_Q
    double total = 0.0;
    for (int i = 0; i < B; ++i) {
        double rs = 0.0;
        for (int j = lane; j < K; j += 32) rs += Q[(size_t)i * K + j];
        total += warp_sum(rs);
    }
    {
        const DivCtx ct = make_div(total);
        for (int p = lane; p < B * K; p += 32) Q[p] = div_ctx(Q[p], ct);
    }
    __syncwarp();
    const double dB = (double)B, dK = (double)K;
    const bool b_pow2 = (B & (B - 1)) == 0, k_pow2 = (K & (K - 1)) == 0;
    const double invB = 1.0 / dB, invK = 1.0 / dK;
    const DivCtx cB = make_div(dB), cK = make_div(dK);
    for (int it = 0; it < iters; ++it) {
        for (int i = 0; i < B; ++i) {
            double rs = 0.0;
            for (int j = lane; j < K; j += 32) rs += Q[(size_t)i * K + j];
            rs = warp_sum(rs);
            const DivCtx cr = make_div(rs);
            if (b_pow2) { for (int j = lane; j < K; j += 32) Q[(size_t)i * K + j] = div_ctx(Q[(size_t)i * K + j], cr) * invB; }
            else { for (int j = lane; j < K; j += 32) Q[(size_t)i * K + j] = div_ctx(div_ctx(Q[(size_t)i * K + j], cr), cB); }
        }
        __syncwarp();
        for (int j = lane; j < K; j += 32) {
            double cs = 0.0;
            for (int i = 0; i < B; ++i) cs += Q[(size_t)i * K + j];
            const DivCtx cc_ = make_div(cs);
            if (k_pow2) { for (int i = 0; i < B; ++i) Q[(size_t)i * K + j] = div_ctx(Q[(size_t)i * K + j], cc_) * invK; }
            else { for (int i = 0; i < B; ++i) Q[(size_t)i * K + j] = div_ctx(div_ctx(Q[(size_t)i * K + j], cc_), cK); }
        }
        __syncwarp();
    }
    for (int p = lane; p < B * K; p += 32) Q[p] = Q[p] * dB;
    __syncwarp();
    // arg-max per row (first maximum wins; NaN counts as maximum like torch.argmax)
    for (int i = 0; i < B; ++i) {
        double best = 0.0;
        int bj = -1;
        for (int j = lane; j < K; j += 32) {
            const double v = Q[(size_t)i * K + j];
            const bool take = bj < 0 || (!(best != best) && ((v != v) || v > best));
            if (take) { best = v; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (oj >= 0) {
                const bool onan = ov != ov, bnan = best != best;
                const bool take = bj < 0 || (onan && (!bnan || oj < bj)) || (!onan && !bnan && (ov > best || (ov == best && oj < bj)));
                if (take) { best = ov; bj = oj; }
            }
        }
        if (lane == 0) codes[gi[i] * L + (L - 1)] = bj;
    }
}

// ---- column-owner kernels: the fp64 matrix lives in REGISTERS ------------------------------------------------------------
// A team of WARPS warps owns one group; thread (w, l) of the team holds the columns j = l + 32·(w + WARPS·u), u < CPT, of every
// row (K = 32·WARPS·CPT), BMAX·CPT doubles in registers.  The column pass (sum over the rows in order, divide) is then
// register-local: no shared memory, no barrier.  The row pass needs Σ_j Q[i][j] in the order every Sinkhorn kernel of this
// file uses — lane l adds its columns l, l+32, l+64 … in ascending order, then the xor butterfly — which for WARPS = 1 is
// again register-local (one warp per group, several groups per CTA; EXACT: the launch serves groups of exactly BMAX rows, so
// every row loop is resolved at compile time and the loop body stays small enough for the instruction cache); for WARPS = 8
// the values pass once through a shared exchange tile ([row][column], conflict-free both ways), warp i mod 8 folds row i and
// publishes the row's division context.
// Same sums in the same order, the same quotients, centre / exp expressions and arg-max rule as sinkhorn_regroup_kernel ⇒
// bit-identical codes (tests/test_gpu_zz_late_fixtures.py compares the two on real groups).  Groups are taken from a per-size-
// class list (sk_class_lists_kernel) through an atomic ticket, so a launch is persistent and self-balancing.
constexpr int SK_MAX_CLASSES = 12;
struct SkClassArgs {
    int n_classes;
    int lo[SK_MAX_CLASSES], hi[SK_MAX_CLASSES];       // class c holds groups with lo[c] <= rows <= hi[c]
};

__global__ void sk_class_lists_kernel(const int64_t *__restrict__ offsets, int64_t n_groups, SkClassArgs ca,
                                      int *__restrict__ lists, unsigned *__restrict__ counts) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = -1;
    if (g < n_groups) {
        const int64_t B = offsets[g + 1] - offsets[g];
        for (int k = 0; k < ca.n_classes; ++k)
            if (B >= ca.lo[k] && B <= ca.hi[k]) c = k;
    }
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < ca.n_classes; ++k) {
        const unsigned mk = __ballot_sync(0xffffffffu, c == k);
        if (!mk) continue;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&counts[k], (unsigned)__popc(mk));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (c == k) lists[(size_t)k * n_groups + base + __popc(mk & ((1u << lane) - 1))] = (int)g;
    }
}

// A pass whose numerators are not all inside the windows of their contexts (rare: an entry next to the bottom of the fp64
// range): the reference's literal two divisions, element by element, on a copy of the thread's elements in shared memory —
// compact looped code that stays out of the hot loop body.  elem(i, u) at el[i·is + u·us]; den[d·dpitch] is the row sum
// (by_row) or the column sum of the element.
__device__ __noinline__ void sk_slow_pass(double *el, int is, int us, const double *den, int dpitch, int B, int ncol, bool by_row,
                                          bool pow2, double inv, double dcount) {
    const DivCtx cD = make_div(dcount);
    for (int i = 0; i < B; ++i)
        for (int u = 0; u < ncol; ++u) {
            const double d = den[(by_row ? i : u) * dpitch];
            double *x = el + (size_t)i * is + (size_t)u * us;
            *x = pow2 ? (*x / d) * inv : div_ctx(*x / d, cD);
        }
}

template <int CPT, int WARPS, int BMAX, int E, bool EXACT>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : 32 * WARPS)
sinkhorn_regroup_own_kernel(const float *__restrict__ residual, const int64_t *__restrict__ items,
                            const int64_t *__restrict__ offsets, const int *__restrict__ glist,
                            const unsigned *__restrict__ gcount, unsigned *__restrict__ gnext,
                            const float *__restrict__ cb, const float *__restrict__ cc, int L, double epsilon, int iters,
                            int64_t *__restrict__ codes) {
    constexpr int K = 32 * WARPS * CPT;
    constexpr int TEAMS = WARPS == 1 ? 4 : 1;
    constexpr bool K_POW2 = (K & (K - 1)) == 0;
    static_assert(E % 16 == 0, "e_dim must be a multiple of 16");
    static_assert(!EXACT || WARPS == 1, "exact row counts are a property of the one-warp launches");
    extern __shared__ __align__(16) unsigned char sk_smem[];
    const int lane = threadIdx.x & 31;
    const int w = WARPS == 1 ? 0 : (threadIdx.x >> 5);
    const int team = WARPS == 1 ? (threadIdx.x >> 5) : 0;
    const int tid_team = WARPS == 1 ? lane : threadIdx.x;
    constexpr int TEAM_THREADS = 32 * WARPS;
    // shared memory per team: rows float[BMAX][E], norms float[BMAX], the row sums of the pass double[BMAX], the thread's
    // column sums double[CPT][threads]; WARPS = 1: a scratch copy of every thread's elements for the slow pass; WARPS > 1: the
    // exchange tile double[BMAX][K] (doubles as that scratch), one division context per row, per-warp partials
    constexpr size_t ROWS_BYTES = (sizeof(float) * BMAX * (E + 1) + 15) / 16 * 16;
    constexpr size_t TEAM_DOUBLES = (size_t)BMAX + (size_t)CPT * TEAM_THREADS + (WARPS == 1 ? (size_t)BMAX * CPT * 32 : 0);
    constexpr size_t TEAM_BYTES = (ROWS_BYTES + sizeof(double) * TEAM_DOUBLES + 15) / 16 * 16;
    float *s_r = reinterpret_cast<float *>(sk_smem + (size_t)team * TEAM_BYTES);
    float *s_xx = s_r + BMAX * E;
    double *den_row = reinterpret_cast<double *>(sk_smem + (size_t)team * TEAM_BYTES + ROWS_BYTES);          // [BMAX]
    double *den_col = den_row + BMAX + tid_team;                                                             // [CPT], pitch TEAM_THREADS
    double *ex = reinterpret_cast<double *>(sk_smem + (size_t)TEAMS * TEAM_BYTES);                           // WARPS > 1 only
    DivCtx *s_ctx = reinterpret_cast<DivCtx *>(ex + (size_t)BMAX * K);
    double *s_part = reinterpret_cast<double *>(s_ctx + BMAX);           // [WARPS]
    float *s_mm = reinterpret_cast<float *>(s_part + WARPS);             // [2][WARPS]
    // where this thread's element (i, u) sits while a slow pass works on it
    double *el = WARPS == 1 ? den_row + BMAX + (size_t)CPT * TEAM_THREADS + lane : ex + lane + 32 * w;
    constexpr int EL_IS = WARPS == 1 ? CPT * 32 : K, EL_US = WARPS == 1 ? 32 : 32 * WARPS;
    __shared__ unsigned s_ticket;
    auto team_sync = [&]() { if (WARPS == 1) __syncwarp(); else __syncthreads(); };
    const unsigned n_listed = *gcount;

    for (;;) {
        unsigned ticket;
        if (WARPS == 1) {
            __syncwarp();
            ticket = 0;
            if (lane == 0) ticket = atomicAdd(gnext, 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
        } else {
            __syncthreads();                                  // the previous group's arg-max has left the exchange tile
            if (threadIdx.x == 0) s_ticket = atomicAdd(gnext, 1u);
            __syncthreads();
            ticket = s_ticket;
        }
        if (ticket >= n_listed) return;
        const int64_t g = glist[ticket];
        const int64_t g0 = offsets[g];
        const int B = EXACT ? BMAX : (int)(offsets[g + 1] - g0);
        const int64_t *gi = items + g0;
        for (int i = tid_team; i < B * (E / 4); i += TEAM_THREADS) {
            const int r = i / (E / 4), q4 = i % (E / 4);
            reinterpret_cast<float4 *>(s_r + r * E)[q4] = reinterpret_cast<const float4 *>(residual + gi[r] * E)[q4];
        }
        team_sync();
        for (int i = tid_team; i < B; i += TEAM_THREADS) s_xx[i] = sumsq_aten_rt(s_r + i * E, E);
        team_sync();

        // d[i][j] = (xx_i + cc_j) - 2 <r_i, c_j>  (vq.py:71-73), the product in the order of a batch of B rows
        double Q[BMAX][CPT];
        float mx = -__int_as_float(0x7f800000), mn = __int_as_float(0x7f800000);
        const bool l16 = small_batch_lane16(B, E);
#pragma unroll 1
        for (int u = 0; u < CPT; ++u) {
            const int j = lane + 32 * (w + WARPS * u);
            const float *c = cb + (size_t)j * E;
            const float ccj = cc[j];
            float dist[BMAX];
            if (!l16) {
                float acc[BMAX];
#pragma unroll
                for (int i = 0; i < BMAX; ++i) acc[i] = 0.0f;
#pragma unroll 1
                for (int k0 = 0; k0 < E; k0 += 16) {
                    float cv[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a = *reinterpret_cast<const float4 *>(c + k0 + 4 * q);
                        cv[4 * q] = a.x; cv[4 * q + 1] = a.y; cv[4 * q + 2] = a.z; cv[4 * q + 3] = a.w;
                    }
#pragma unroll
                    for (int i = 0; i < BMAX; ++i) {
                        if (i < B) {
                            const float *r = s_r + i * E + k0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 b = *reinterpret_cast<const float4 *>(r + 4 * q);
                                acc[i] = __fmaf_rn(b.x, cv[4 * q], acc[i]);
                                acc[i] = __fmaf_rn(b.y, cv[4 * q + 1], acc[i]);
                                acc[i] = __fmaf_rn(b.z, cv[4 * q + 2], acc[i]);
                                acc[i] = __fmaf_rn(b.w, cv[4 * q + 3], acc[i]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < BMAX; ++i) dist[i] = __fsub_rn(__fadd_rn(s_xx[i < B ? i : 0], ccj), __fmul_rn(2.0f, acc[i]));
            } else {
                // small-batch order (2 <= B and 24·B <= E, i.e. B <= E/24 rows): 16 interleaved chains per product, folded
                constexpr int LB = (E / 24 < BMAX ? E / 24 : BMAX) > 0 ? (E / 24 < BMAX ? E / 24 : BMAX) : 1;
                float a16[LB][16];
#pragma unroll
                for (int i = 0; i < LB; ++i)
#pragma unroll
                    for (int q = 0; q < 16; ++q) a16[i][q] = 0.0f;
#pragma unroll 1
                for (int k0 = 0; k0 < E; k0 += 16) {
                    float cv[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a = *reinterpret_cast<const float4 *>(c + k0 + 4 * q);
                        cv[4 * q] = a.x; cv[4 * q + 1] = a.y; cv[4 * q + 2] = a.z; cv[4 * q + 3] = a.w;
                    }
#pragma unroll
                    for (int i = 0; i < LB; ++i) {
                        if (i < B) {
                            const float *r = s_r + i * E + k0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 b = *reinterpret_cast<const float4 *>(r + 4 * q);
                                a16[i][4 * q] = __fmaf_rn(b.x, cv[4 * q], a16[i][4 * q]);
                                a16[i][4 * q + 1] = __fmaf_rn(b.y, cv[4 * q + 1], a16[i][4 * q + 1]);
                                a16[i][4 * q + 2] = __fmaf_rn(b.z, cv[4 * q + 2], a16[i][4 * q + 2]);
                                a16[i][4 * q + 3] = __fmaf_rn(b.w, cv[4 * q + 3], a16[i][4 * q + 3]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < BMAX; ++i) dist[i] = 0.0f;
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    float s4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        s4[q] = __fadd_rn(__fadd_rn(__fadd_rn(a16[i][q], a16[i][q + 4]), a16[i][q + 8]), a16[i][q + 12]);
                    const float acc = __fadd_rn(__fadd_rn(s4[0], s4[1]), __fadd_rn(s4[2], s4[3]));
                    dist[i] = __fsub_rn(__fadd_rn(s_xx[i < B ? i : 0], ccj), __fmul_rn(2.0f, acc));
                }
            }
            // the column loop is rolled (u is a run-time value here): the thread's matrix is addressed with static indices
            // only, so the distances of this column are routed to their registers by a compile-time switch
#pragma unroll
            for (int uu = 0; uu < CPT; ++uu) {
                if (uu == u) {
#pragma unroll
                    for (int i = 0; i < BMAX; ++i) {
                        if (i < B) {
                            Q[i][uu] = (double)dist[i];
                            mx = fmaxf(mx, dist[i]);
                            mn = fminf(mn, dist[i]);
                        } else {
                            Q[i][uu] = 0.0;                   // rows beyond the group stay exactly zero (see for_rows)
                        }
                    }
                }
            }
        }
        // centre over the whole group (vq.py:51-61), exp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (WARPS > 1) {
            if (lane == 0) { s_mm[w] = mx; s_mm[WARPS + w] = mn; }
            __syncthreads();
            for (int v = 0; v < WARPS; ++v) { mx = fmaxf(mx, s_mm[v]); mn = fminf(mn, s_mm[WARPS + v]); }
        }
        const float middle = __fdiv_rn(__fadd_rn(mx, mn), 2.0f);
        const float amplitude = __fadd_rn(__fsub_rn(mx, middle), 1e-5f);
#pragma unroll
        for (int i = 0; i < BMAX; ++i) {
            if (i < B) {
#pragma unroll
                for (int u = 0; u < CPT; ++u) {
                    const float cd = __fdiv_rn(__fsub_rn((float)Q[i][u], middle), amplitude);
                    Q[i][u] = exp(-((double)cd) / epsilon);
                }
            }
        }

        // Row loops of the hot path.  The rows are walked in blocks of RB: one uniform branch per block, straight-line code
        // inside, so the independent quotients of a block overlap in the fp64 pipe.  Rows of the last block beyond the group
        // ("not live") hold exactly 0.0 and are carried along: 0 / x = 0 through the fma sequence, and a trailing + 0.0 changes
        // no column sum; their window tests are masked.  With EXACT every row is live and all of this folds away.
        constexpr int RB = 4, NBLK = (BMAX + RB - 1) / RB;
        auto for_rows = [&](auto &&f) {
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                if (EXACT || blk * RB < B) {
#pragma unroll
                    for (int ii = 0; ii < RB; ++ii)
                        if (blk * RB + ii < BMAX) f(blk * RB + ii, EXACT || blk * RB + ii < B);
                }
            }
        };
        auto publish = [&]() {                                  // WARPS > 1: every thread stores its columns of every row
            for_rows([&](int i, bool) {
#pragma unroll
                for (int u = 0; u < CPT; ++u) ex[i * K + lane + 32 * (w + WARPS * u)] = Q[i][u];
            });
        };
        auto lane_row_sum = [&](int i) {                        // WARPS = 1
            double rs = 0.0;
#pragma unroll
            for (int u = 0; u < CPT; ++u) rs += Q[i][u];
            return warp_sum(rs);
        };
        auto ex_row_sum = [&](int i) {                          // WARPS > 1, called by the warp that owns row i
            double rs = 0.0;
            for (int j = lane; j < K; j += 32) rs += ex[i * K + j];
            return warp_sum(rs);
        };
        // the thread's elements → scratch, the literal divisions there, and back (see sk_slow_pass)
        auto to_scratch = [&]() {
#pragma unroll
            for (int i = 0; i < BMAX; ++i) {
                if (i < B) {
#pragma unroll
                    for (int u = 0; u < CPT; ++u) el[(size_t)i * EL_IS + (size_t)u * EL_US] = Q[i][u];
                }
            }
        };
        auto from_scratch = [&]() {
#pragma unroll
            for (int i = 0; i < BMAX; ++i) {
                if (i < B) {
#pragma unroll
                    for (int u = 0; u < CPT; ++u) Q[i][u] = el[(size_t)i * EL_IS + (size_t)u * EL_US];
                }
            }
        };
        auto slow_rows = [&](bool pow2, double inv, double dcount) {      // by the row sums in den_row
            to_scratch();
            sk_slow_pass(el, EL_IS, EL_US, den_row, 1, B, CPT, true, pow2, inv, dcount);
            from_scratch();
        };
        // every element of the thread divided by ONE number, literally (the denominator goes through the thread's own
        // column-sum slots, which are free outside the column pass)
        auto slow_all = [&](double d) {
#pragma unroll
            for (int u = 0; u < CPT; ++u) den_col[(size_t)u * TEAM_THREADS] = d;
            to_scratch();
            sk_slow_pass(el, EL_IS, EL_US, den_col, TEAM_THREADS, B, CPT, false, true, 1.0, 1.0);
            from_scratch();
        };
        const DivCtx cOne = make_div(1.0);

        // sum_Q = Q.sum(-1).sum(-2);  Q /= sum_Q      (row sums added per warp in row order, warps in order: as sinkhorn_cta)
        double total;
        if (WARPS == 1) {
            total = 0.0;
#pragma unroll
            for (int i = 0; i < BMAX; ++i)
                if (i < B) total += lane_row_sum(i);
        } else {
            publish();
            __syncthreads();
            for (int i = w; i < B; i += WARPS) {
                const double rs = ex_row_sum(i);
                if (lane == 0) den_row[i] = rs;
            }
            __syncthreads();
            // sinkhorn_cta's order: eight partials (rows i ≡ v mod 8 in ascending order), added in order
            total = 0.0;
            for (int v = 0; v < SK_THREADS / 32; ++v) {
                double part = 0.0;
                for (int i = v; i < B; i += SK_THREADS / 32) part += den_row[i];
                total += part;
            }
            __syncthreads();                                  // den_row is rewritten by the first row pass
        }
        {
            const DivCtx ct = make_div(total);
            bool outside = false;
            for_rows([&](int i, bool live) {
#pragma unroll
                for (int u = 0; u < CPT; ++u) outside |= live && div_outside(Q[i][u], ct);
            });
            if (__builtin_expect(!outside, 1)) {
                for_rows([&](int i, bool) {
#pragma unroll
                    for (int u = 0; u < CPT; ++u) Q[i][u] = div_fast(Q[i][u], ct);
                });
            } else {
                slow_all(total);
            }
        }
        // Every pass below is "divide a batch of numerators by their row / column sum, then by B / K".  With a power-of-two
        // B (K) the second division is an exact scaling, so the pair is ONE quotient by sum·B (an exact product): the same
        // bits as long as that quotient is normal.  A pass first tests all its numerators against the windows of their
        // contexts (two integer instructions each); if all are inside — the rule — the quotients are straight-line fma
        // sequences that overlap in the fp64 pipe; otherwise the pass is redone with the reference's literal two divisions.
        const double dB = (double)B, dK = (double)K;
        const bool b_pow2 = (B & (B - 1)) == 0;
        const double invB = 1.0 / dB, invK = 1.0 / dK;
        const DivCtx cB = make_div(dB);
        for (int it = 0; it < iters; ++it) {
            // Q /= Q.sum(dim=1, keepdim=True);  Q /= B
            if (WARPS > 1) {
                publish();
                __syncthreads();
                // this warp's rows (i = w, w + WARPS, …) are summed TOGETHER — independent chains, one butterfly pass for all —
                // and lane r turns row r's sum into its division context, so the phase costs one division latency, not one
                // per row.  Rows beyond the group read stale tile rows; their sums go nowhere.
                constexpr int RPW = (BMAX + WARPS - 1) / WARPS;
                double rs[RPW];
#pragma unroll
                for (int r = 0; r < RPW; ++r) rs[r] = 0.0;
#pragma unroll
                for (int t = 0; t < K / 32; ++t) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r)
                        if (w + r * WARPS < BMAX) rs[r] += ex[(w + r * WARPS) * K + lane + 32 * t];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], o);
                }
                double mine = rs[0];
#pragma unroll
                for (int r = 1; r < RPW; ++r)
                    if (lane == r) mine = rs[r];
                if (lane < RPW && w + lane * WARPS < B) {
                    const int i = w + lane * WARPS;
                    s_ctx[i] = make_div(b_pow2 ? mine * dB : mine);
                    den_row[i] = mine;
                }
                __syncthreads();
            }
            {
                constexpr int NCR = WARPS == 1 ? BMAX : 1;      // WARPS > 1: the contexts stay in shared memory
                DivCtx cr[NCR];
                bool outside = false;
                for_rows([&](int i, bool live) {
                    if (WARPS == 1) {
                        const double rs = lane_row_sum(i);
                        cr[i % NCR] = make_div(b_pow2 ? rs * dB : rs);
                    }
                    const DivCtx c = WARPS == 1 ? cr[i % NCR] : s_ctx[live ? i : 0];
#pragma unroll
                    for (int u = 0; u < CPT; ++u) outside |= live && div_outside(Q[i][u], c);
                });
                if (__builtin_expect(!outside, 1)) {
                    for_rows([&](int i, bool live) {
                        const DivCtx c = WARPS == 1 ? cr[i % NCR] : s_ctx[live ? i : 0];       // a dead row: 0 / (any row's sum) = 0
#pragma unroll
                        for (int u = 0; u < CPT; ++u) Q[i][u] = div_fast(Q[i][u], c);
                    });
                    if (!b_pow2) {                              // Q /= B as a second batch of quotients
                        bool out2 = false;
                        for_rows([&](int i, bool live) {
#pragma unroll
                            for (int u = 0; u < CPT; ++u) out2 |= live && div_outside(Q[i][u], cB);
                        });
                        if (__builtin_expect(!out2, 1)) {
                            for_rows([&](int i, bool) {
#pragma unroll
                                for (int u = 0; u < CPT; ++u) Q[i][u] = div_fast(Q[i][u], cB);
                            });
                        } else {
                            slow_all(dB);
                        }
                    }
                } else {
                    if (WARPS == 1) {
#pragma unroll
                        for (int i = 0; i < BMAX; ++i)
                            if (i < B) den_row[i] = b_pow2 ? cr[i % NCR].b * invB : cr[i % NCR].b;   // the row sum itself (exact un-scaling)
                    }
                    slow_rows(b_pow2, invB, dB);
                }
            }
            // Q /= Q.sum(dim=0, keepdim=True);  Q /= K      (a column lives in one thread); eight columns at a time
            constexpr int CCH = CPT < 8 ? CPT : 8;
#pragma unroll
            for (int u0 = 0; u0 < CPT; u0 += CCH) {
                DivCtx ccol[CCH];
                bool outside = false;
                double cs[CCH];
#pragma unroll
                for (int v = 0; v < CCH; ++v) cs[v] = 0.0;
                for_rows([&](int i, bool) {
#pragma unroll
                    for (int v = 0; v < CCH; ++v) cs[v] += Q[i][u0 + v];
                });
#pragma unroll
                for (int v = 0; v < CCH; ++v) ccol[v] = make_div(K_POW2 ? cs[v] * dK : cs[v]);
                for_rows([&](int i, bool live) {
#pragma unroll
                    for (int v = 0; v < CCH; ++v) outside |= live && div_outside(Q[i][u0 + v], ccol[v]);
                });
                if (__builtin_expect(!outside && K_POW2, 1)) {
                    for_rows([&](int i, bool) {
#pragma unroll
                        for (int v = 0; v < CCH; ++v) Q[i][u0 + v] = div_fast(Q[i][u0 + v], ccol[v]);
                    });
                } else {
                    // only this chunk goes through the literal divisions: columns are independent in this pass
#pragma unroll
                    for (int v = 0; v < CCH; ++v) den_col[(size_t)v * TEAM_THREADS] = cs[v];
#pragma unroll
                    for (int i = 0; i < BMAX; ++i) {
                        if (i < B) {
#pragma unroll
                            for (int v = 0; v < CCH; ++v) el[(size_t)i * EL_IS + (size_t)v * EL_US] = Q[i][u0 + v];
                        }
                    }
                    sk_slow_pass(el, EL_IS, EL_US, den_col, TEAM_THREADS, B, CCH, false, K_POW2, invK, dK);
#pragma unroll
                    for (int i = 0; i < BMAX; ++i) {
                        if (i < B) {
#pragma unroll
                            for (int v = 0; v < CCH; ++v) Q[i][u0 + v] = el[(size_t)i * EL_IS + (size_t)v * EL_US];
                        }
                    }
                }
            }
        }
        // Q *= B;  arg-max per row (first maximum wins; NaN counts as maximum like torch.argmax)
        if (WARPS == 1) {
#pragma unroll
            for (int i = 0; i < BMAX; ++i) {
                if (i < B) {
                    double best = 0.0;
                    int bj = -1;
#pragma unroll
                    for (int u = 0; u < CPT; ++u) {
                        const double v = Q[i][u] * dB;
                        const bool take = bj < 0 || (!(best != best) && ((v != v) || v > best));
                        if (take) { best = v; bj = lane + 32 * u; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                        const bool onan = ov != ov, bnan = best != best;
                        const bool take = (onan && (!bnan || oj < bj)) || (!onan && !bnan && (ov > best || (ov == best && oj < bj)));
                        if (take) { best = ov; bj = oj; }
                    }
                    if (lane == 0) codes[gi[i] * L + (L - 1)] = bj;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < BMAX; ++i) {
                if (i < B) {
#pragma unroll
                    for (int u = 0; u < CPT; ++u) ex[i * K + lane + 32 * (w + WARPS * u)] = Q[i][u] * dB;
                }
            }
            __syncthreads();
            for (int i = w; i < B; i += WARPS) {
                double best = 0.0;
                int bj = -1;
                for (int j = lane; j < K; j += 32) {
                    const double v = ex[i * K + j];
                    const bool take = bj < 0 || (!(best != best) && ((v != v) || v > best));
                    if (take) { best = v; bj = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                    const bool onan = ov != ov, bnan = best != best;
                    const bool take = (onan && (!bnan || oj < bj)) || (!onan && !bnan && (ov > best || (ov == best && oj < bj)));
                    if (take) { best = ov; bj = oj; }
                }
                if (lane == 0) codes[gi[i] * L + (L - 1)] = bj;
            }
        }
    }
}

// collision groups too large for shared memory: same computation, the fp64 matrix and the row norms live in a global
// scratch slice (offset given per listed group).  One CTA per listed group.
__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_regroup_large_kernel(const float *__restrict__ residual, const int64_t *__restrict__ items,
                              const int64_t *__restrict__ offsets, const int64_t *__restrict__ group_ids,
                              const int64_t *__restrict__ scratch_off, double *__restrict__ scratch, int e,
                              const float *__restrict__ cb, const float *__restrict__ cc, int K, int L, double epsilon,
                              int iters, int64_t *__restrict__ codes) {
    __shared__ double s_red[SK_THREADS / 32];
    __shared__ float s_mm[64];
    const int64_t g = group_ids[blockIdx.x];
    const int64_t g0 = offsets[g], g1 = offsets[g + 1];
    const int B = (int)(g1 - g0);
    double *Q = scratch + scratch_off[blockIdx.x];              // [B][K]
    float *xx = reinterpret_cast<float *>(Q + (size_t)B * K);    // [B]
    const int64_t *gi = items + g0;
    const int tid = threadIdx.x;
    for (int i = tid; i < B; i += SK_THREADS) xx[i] = sumsq_aten_rt(residual + gi[i] * e, e);
    __syncthreads();
    for (int64_t p = tid; p < (int64_t)B * K; p += SK_THREADS) {
        const int i = (int)(p / K), j = (int)(p % K);
        const float *r = residual + gi[i] * e;
        const float *c = cb + (size_t)j * e;
        const float acc = dot_for_batch(r, c, e, B);
        Q[p] = (double)__fsub_rn(__fadd_rn(xx[i], cc[j]), __fmul_rn(2.0f, acc));
    }
    __syncthreads();
    center_and_exp(Q, B * K, epsilon, s_mm);
    sinkhorn_cta(Q, B, K, iters, s_red);
    argmax_rows(Q, B, K, gi, codes, L, nullptr);
}

// ---- general [B,K] matrices in global memory (training-size batches) ----------------------------
// One CTA walks the whole matrix: simple and exact in structure; B*K is at most a few million here.
__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_global_kernel(double *__restrict__ Q, int B, int K, double epsilon, int iters) {
    __shared__ double s_red[SK_THREADS / 32];
    for (int i = threadIdx.x; i < B * K; i += SK_THREADS) Q[i] = exp(-Q[i] / epsilon);   // layers.py:87
    __syncthreads();
    sinkhorn_cta(Q, B, K, iters, s_red);
}

__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_assign_kernel(const float *__restrict__ d, double *__restrict__ Q, int B, int K, double epsilon,
                       int iters, int64_t *__restrict__ idx) {
    __shared__ double s_red[SK_THREADS / 32];
    __shared__ float s_mm[64];
    for (int i = threadIdx.x; i < B * K; i += SK_THREADS) Q[i] = (double)d[i];
    __syncthreads();
    center_and_exp(Q, B * K, epsilon, s_mm);
    sinkhorn_cta(Q, B, K, iters, s_red);
    argmax_rows(Q, B, K, nullptr, nullptr, 0, idx);
}


// ---- training-size batches on the whole GPU ------------------------------------------------------
// vq.py:77-83 on one [B,K] batch: the rows are split over up to 148 co-resident CTAs (cooperative launch), each CTA keeps
// its rows as fp64 in shared memory for all iterations.  Row normalisation is local; column sums are the only global
// quantity: per iteration every CTA publishes its partial column sums, one grid barrier, then every CTA adds the
// partials in CTA order (deterministic).
__device__ __forceinline__ void grid_barrier(unsigned long long *counter, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_assign_grid_kernel(const float *__restrict__ d, double *__restrict__ ws, int B, int K, int rows_per_cta, double epsilon,
                            int iters, int64_t *__restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char sk_smem[];
    __shared__ double s_red[SK_THREADS / 32];
    __shared__ float s_mm[64];
    double *Q = reinterpret_cast<double *>(sk_smem);                 // [rows][K]
    const int G = gridDim.x, cta = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = SK_THREADS / 32;
    const int row0 = cta * rows_per_cta;
    const int rows = max(0, min(B, row0 + rows_per_cta) - row0);
    unsigned long long *counter = reinterpret_cast<unsigned long long *>(ws);
    double *scal = ws + 8;                                           // [G][4]: max, min, total
    double *colp = ws + 8 + (size_t)4 * G;                           // [2][G][K] partial column sums
    double *vglob = colp + (size_t)2 * G * K;                        // [2][K] column factors of the iteration
    const int cpc = (K + G - 1) / G;                                 // columns whose sum this CTA owns
    unsigned long long epoch = 0;
    const int count = rows * K;
    // centre (fp32, over the WHOLE matrix) …
    float mx = -__int_as_float(0x7f800000), mn = __int_as_float(0x7f800000);
    for (int i = tid; i < count; i += SK_THREADS) {
        const float v = d[(size_t)row0 * K + i];
        Q[i] = (double)v;
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) { s_mm[wid] = mx; s_mm[32 + wid] = mn; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NW; ++w) { mx = fmaxf(mx, s_mm[w]); mn = fminf(mn, s_mm[32 + w]); }
        scal[cta * 4 + 0] = (double)mx;
        scal[cta * 4 + 1] = (double)mn;
    }
    grid_barrier(counter, (++epoch) * G);
    mx = -__int_as_float(0x7f800000); mn = __int_as_float(0x7f800000);
    for (int c = 0; c < G; ++c) {
        mx = fmaxf(mx, (float)__ldcg(&scal[c * 4 + 0]));
        mn = fminf(mn, (float)__ldcg(&scal[c * 4 + 1]));
    }
    const float middle = __fdiv_rn(__fadd_rn(mx, mn), 2.0f);
    const float amplitude = __fadd_rn(__fsub_rn(mx, middle), 1e-5f);
    // … exp (fp64), total sum
    double part = 0.0;
    for (int i = tid; i < count; i += SK_THREADS) {
        const float c = __fdiv_rn(__fsub_rn((float)Q[i], middle), amplitude);
        const double q = exp(-((double)c) / epsilon);
        Q[i] = q;
        part += q;
    }
    double tot = block_sum(part, s_red);
    if (tid == 0) scal[cta * 4 + 2] = tot;
    grid_barrier(counter, (++epoch) * G);
    double total = 0.0;
    for (int c = 0; c < G; ++c) total += __ldcg(&scal[c * 4 + 2]);
    for (int i = tid; i < count; i += SK_THREADS) Q[i] = Q[i] / total;
    __syncthreads();
    // The iteration only ever rescales rows and columns, so Q_t = diag(u) Q_0 diag(v): keep Q_0 in shared memory and
    // iterate on the two vectors (2 fp64 FMAs per element and iteration instead of 4 divisions).  Same fixed point and
    // the same arg-max as the element-wise form of layers.py:96-104 up to fp64 rounding.
    double *u = Q + (size_t)rows_per_cta * K;                        // [rows]  this CTA's row factors
    double *v = u + rows_per_cta;                                    // [K]     column factors (identical in every CTA)
    for (int j = tid; j < K; j += SK_THREADS) v[j] = 1.0;
    for (int i = tid; i < rows; i += SK_THREADS) u[i] = 1.0;
    __syncthreads();
    const double dB = (double)B, dK = (double)K;
    for (int it = 0; it < iters; ++it) {
        // Q /= rowsum; Q /= B      →  u_i = 1 / (B * Σ_j Q0_ij v_j)
        for (int i = wid; i < rows; i += NW) {
            double rs = 0.0;
            for (int j = lane; j < K; j += 32) rs = fma(Q[(size_t)i * K + j], v[j], rs);
            rs = warp_sum(rs);
            if (lane == 0) u[i] = (1.0 / rs) / dB;
        }
        __syncthreads();
        // Q /= colsum; Q /= K      →  v_j = 1 / (K * Σ_i u_i Q0_ij), the sum taken over all CTAs
        double *mine = colp + ((size_t)(it & 1) * G + cta) * K;
        for (int j = tid; j < K; j += SK_THREADS) {
            double cs = 0.0;
            for (int i = 0; i < rows; ++i) cs = fma(u[i], Q[(size_t)i * K + j], cs);
            mine[j] = cs;
        }
        grid_barrier(counter, (++epoch) * G);
        const double *all = colp + (size_t)(it & 1) * G * K;
        if (G <= SK_ONE_LEVEL_MAX) {
            // few CTAs: every CTA adds all partials itself (G loads per column, 4 independent chains) — one barrier per
            // iteration
            for (int j = tid; j < K; j += SK_THREADS) {
                double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
                int c = 0;
                for (; c + 3 < G; c += 4) {
                    c0 += __ldcg(&all[(size_t)c * K + j]);
                    c1 += __ldcg(&all[(size_t)(c + 1) * K + j]);
                    c2 += __ldcg(&all[(size_t)(c + 2) * K + j]);
                    c3 += __ldcg(&all[(size_t)(c + 3) * K + j]);
                }
                for (; c < G; ++c) c0 += __ldcg(&all[(size_t)c * K + j]);
                v[j] = (1.0 / ((c0 + c1) + (c2 + c3))) / dK;
            }
            __syncthreads();
            continue;
        }
        // two-level sum: this CTA adds the G partials of the few columns it owns (a warp per column, fixed order) and
        // publishes the new factor; after the second barrier everybody reads the K factors (2 KB instead of G x K)
        double *vnew = vglob + (size_t)(it & 1) * K;
        for (int jj = wid; jj < cpc; jj += NW) {
            const int j = cta * cpc + jj;
            if (j < K) {
                double cs = 0.0;
                for (int c = lane; c < G; c += 32) cs += __ldcg(&all[(size_t)c * K + j]);
                cs = warp_sum(cs);
                if (lane == 0) vnew[j] = (1.0 / cs) / dK;
            }
        }
        grid_barrier(counter, (++epoch) * G);
        for (int j = tid; j < K; j += SK_THREADS) v[j] = __ldcg(&vnew[j]);
        __syncthreads();
    }
    for (int i = tid; i < count; i += SK_THREADS) Q[i] = ((u[i / K] * Q[i]) * v[i % K]) * dB;      // Q *= B
    __syncthreads();
    argmax_rows(Q, rows, K, nullptr, nullptr, 0, idx_out + row0);
}

// div_by against the hardware division on pseudo-random operand pairs: mantissas uniform, exponents spread over ±span binades,
// a fraction of the denominators with all-ones / all-zeros mantissa tails (the hard cases of reciprocal-based division)
__global__ void check_division_kernel(unsigned long long seed, int span, unsigned long long per_thread, unsigned long long *bad) {
    unsigned long long st = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1);
    auto next = [&]() {
        st += 0x9E3779B97F4A7C15ull;
        unsigned long long z = st;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    unsigned long long mism = 0;
    for (unsigned long long t = 0; t < per_thread; ++t) {
        const unsigned long long ra = next(), rb = next(), rc = next();
        unsigned long long mb = rb & 0xFFFFFFFFFFFFFull;
        if ((rc & 7) == 0) mb |= (1ull << (rc >> 58)) - 1;            // trailing ones
        if ((rc & 7) == 1) mb &= ~((1ull << (rc >> 58)) - 1);         // trailing zeros
        const long long expa = 1023 + (long long)((ra >> 52) % (unsigned long long)(2 * span + 1)) - span;
        const long long expb = 1023 + (long long)((rb >> 52) % (unsigned long long)(2 * span + 1)) - span;
        const double a = __longlong_as_double((long long)((unsigned long long)expa << 52 | (ra & 0xFFFFFFFFFFFFFull)));
        const double b = __longlong_as_double((long long)((unsigned long long)expb << 52 | mb));
        const DivCtx c = make_div(b);
        if (__double_as_longlong(div_ctx(a, c)) != __double_as_longlong(a / b)) ++mism;
    }
    if (mism) atomicAdd(bad, mism);
}

}  // namespace
}  // namespace rqb

using namespace rqb;

// diagnostics: number of operand pairs (of `pairs` tried) on which the kernel's reciprocal-based division differs from `a / b`
extern "C" int rqb200_debug_check_division(unsigned long long seed, int exponent_span, unsigned long long pairs,
                                           unsigned long long *mismatches_host, void *stream) {
    RQB_CHECK(mismatches_host != nullptr, "NULL argument");
    RQB_CHECK(exponent_span >= 0 && exponent_span <= 1000, "exponent_span out of range");
    unsigned long long *bad = nullptr;
    RQB_CUDA(cudaMalloc(&bad, sizeof(*bad)));
    RQB_CUDA(cudaMemsetAsync(bad, 0, sizeof(*bad), (cudaStream_t)stream));
    const unsigned blocks = kNumSMs * 8, threads = 256;
    const unsigned long long per_thread = (pairs + (unsigned long long)blocks * threads - 1) / ((unsigned long long)blocks * threads);
    rqb::count_launch();
    check_division_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(seed, exponent_span, per_thread, bad);
    RQB_CUDA(cudaMemcpyAsync(mismatches_host, bad, sizeof(*bad), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RQB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    RQB_CUDA(cudaFree(bad));
    return 0;
}

static int g_sk_variant = 0;      // 0: column-owner register kernels where the shape allows; 1: shared-memory kernels only
extern "C" int rqb200_debug_sinkhorn_variant(int v) {
    g_sk_variant = v;
    return 0;
}

template <int CPT, int WARPS, int BMAX, int E, bool EXACT>
static int launch_own(const rqb200_model *m, const float *residual, const int64_t *items, const int64_t *offsets, int64_t n_groups,
                      const int *glist, const unsigned *gcount, unsigned *gnext, double epsilon, int iters, int64_t *codes,
                      cudaStream_t s) {
    constexpr int K = 32 * WARPS * CPT;
    constexpr int TEAMS = WARPS == 1 ? 4 : 1;
    constexpr int TEAM_THREADS = 32 * WARPS;
    constexpr size_t rows_bytes = (sizeof(float) * BMAX * (E + 1) + 15) / 16 * 16;
    constexpr size_t team_doubles = (size_t)BMAX + (size_t)CPT * TEAM_THREADS + (WARPS == 1 ? (size_t)BMAX * CPT * 32 : 0);
    constexpr size_t team_bytes = (rows_bytes + sizeof(double) * team_doubles + 15) / 16 * 16;
    constexpr size_t smem = WARPS == 1 ? TEAMS * team_bytes
                                       : team_bytes + sizeof(double) * BMAX * K + sizeof(DivCtx) * BMAX + sizeof(double) * WARPS +
                                             sizeof(float) * 2 * WARPS + 16;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static rqb::DeviceOnce once;
    if (once.first())
        RQB_CUDA(cudaFuncSetAttribute(sinkhorn_regroup_own_kernel<CPT, WARPS, BMAX, E, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const int L = m->L;
    int64_t ctas = (n_groups + TEAMS - 1) / TEAMS;
    if (ctas > kNumSMs * 8) ctas = kNumSMs * 8;             // persistent: groups are taken through an atomic ticket
    rqb::count_launch();
    sinkhorn_regroup_own_kernel<CPT, WARPS, BMAX, E, EXACT><<<(unsigned)ctas, WARPS == 1 ? 128 : 32 * WARPS, smem, s>>>(
        residual, items, offsets, glist, gcount, gnext, m->cb[L - 1], m->cc[L - 1], L, epsilon, iters, codes);
    RQB_LAUNCH_CHECK();
    return 0;
}

template <int E>
static int regroup_own(rqb200_model *m, const float *residual, const int64_t *items, const int64_t *offsets, int64_t n_groups,
                       int top, double epsilon, int iters, int64_t *codes, cudaStream_t s, int *covered_up_to) {
    const int K = m->K[m->L - 1];
    SkClassArgs ca;
    ca.n_classes = 0;
    for (int k = 0; k < SK_MAX_CLASSES; ++k) { ca.lo[k] = 0; ca.hi[k] = -1; }
    auto add = [&](int lo, int hi) { ca.lo[ca.n_classes] = lo; ca.hi[ca.n_classes] = hi; ++ca.n_classes; };
    if (K == 256) {
        for (int b = 2; b <= 8; ++b) add(b, b);             // one launch per row count: one warp per group, rows resolved at compile time
        add(9, 16);
        add(17, 32);
        add(33, 64);
    } else {
        add(2, 2);
        add(3, 16);
    }
    RQB_TRY(ws_reserve(m->skws, sizeof(int) * (size_t)ca.n_classes * n_groups + 128));
    unsigned *counts = (unsigned *)m->skws.ptr;              // [16] listed groups per class, [16] tickets
    unsigned *next = counts + 16;
    int *lists = (int *)((char *)m->skws.ptr + 128);
    RQB_CUDA(cudaMemsetAsync(counts, 0, 128, s));
    rqb::count_launch();
    sk_class_lists_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(offsets, n_groups, ca, lists, counts);
    RQB_LAUNCH_CHECK();
#define RQB_OWN(CPT, WARPS, BMAX, EXACT) \
    RQB_TRY((launch_own<CPT, WARPS, BMAX, E, EXACT>(m, residual, items, offsets, n_groups, lists + (size_t)c * n_groups, counts + c, next + c, \
                                                    epsilon, iters, codes, s)))
    for (int c = 0; c < ca.n_classes; ++c) {
        if (ca.lo[c] > top) break;
        if (K == 256) {
            switch (c) {
                case 0: RQB_OWN(8, 1, 2, true); break;
                case 1: RQB_OWN(8, 1, 3, true); break;
                case 2: RQB_OWN(8, 1, 4, true); break;
                case 3: RQB_OWN(8, 1, 5, true); break;
                case 4: RQB_OWN(8, 1, 6, true); break;
                case 5: RQB_OWN(8, 1, 7, true); break;
                case 6: RQB_OWN(8, 1, 8, true); break;
                case 7: RQB_OWN(4, 2, 16, false); break;       // 64 doubles per thread in every class
                case 8: RQB_OWN(2, 4, 32, false); break;
                default: RQB_OWN(1, 8, 64, false); break;
            }
        } else {
            if (c == 0) RQB_OWN(32, 1, 2, true);
            else RQB_OWN(4, 8, 16, false);
        }
    }
#undef RQB_OWN
    *covered_up_to = ca.hi[ca.n_classes - 1];
    return 0;
}

extern "C" int rqb200_sinkhorn_regroup(rqb200_model *m, const float *residual_dev, const int64_t *items_dev,
                                       const int64_t *offsets_dev, int64_t n_groups, int max_group,
                                       double epsilon, int iters, int64_t *codes_dev, void *stream) {
    RQB_CHECK(m != nullptr, "model is NULL");
    if (n_groups == 0) return 0;
    RQB_CHECK(residual_dev && items_dev && offsets_dev && codes_dev, "NULL buffer");
    RQB_CHECK(epsilon > 0.0, "epsilon must be > 0 (the argmin branch needs no re-encode)");
    RQB_CHECK(m->e < 512, "e_dim >= 512 not supported");
    RQB_CUDA(cudaSetDevice(m->device));
    const int L = m->L, K = m->K[L - 1], e = m->e;
    RQB_CHECK(m->cb_set[L - 1], "codebook %d not loaded", L - 1);
    // rows that fit in shared memory next to the fp64 matrix
    const size_t budget = 200 * 1024;
    const size_t per_row = sizeof(double) * K + sizeof(float) * (e + 1);
    const int cap_max = (int)(budget / per_row);
    RQB_CHECK(cap_max >= 2, "K=%d too large for the shared-memory Sinkhorn kernel", K);
    int top = cap_max;
    if (max_group > 0 && max_group < top) top = max_group < 2 ? 2 : max_group;
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(sinkhorn_regroup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(budget + 1024)));
    }
    // Most groups have a handful of members: one launch per size class, shared memory sized for the class, so that many
    // small groups share an SM instead of every CTA reserving room for the largest group.
    ProfScope ps(PROF_SINKHORN, (cudaStream_t)stream);
    const int bounds[4] = {8, 24, 48, top};
    int lo = 2;
    if (g_sk_variant == 0 && (K == 256 || K == 1024) && (e == 32 || e == 64) && n_groups < ((int64_t)1 << 31)) {
        // register-resident kernels for the bulk of the groups; what is larger falls through to the shared-memory kernel
        int covered = 0;
        if (e == 32) RQB_TRY(regroup_own<32>(m, residual_dev, items_dev, offsets_dev, n_groups, top, epsilon, iters, codes_dev,
                                             (cudaStream_t)stream, &covered));
        else RQB_TRY(regroup_own<64>(m, residual_dev, items_dev, offsets_dev, n_groups, top, epsilon, iters, codes_dev,
                                     (cudaStream_t)stream, &covered));
        lo = covered + 1;
    } else {
        // groups of 2 … 8 rows: one warp per group (same bits as the CTA kernel below, no block barriers)
        const size_t per_warp = sizeof(double) * SKW_MAX_ROWS * (size_t)K + sizeof(float) * SKW_MAX_ROWS * (size_t)(e + 1);
        int wpc = (int)((size_t)(100 * 1024) / per_warp);
        wpc = wpc > 4 ? 4 : wpc;
        if (wpc >= 1 && (per_warp % 16) == 0) {
            static rqb::DeviceOnce warp_attr_once;
            if (warp_attr_once.first())
                RQB_CUDA(cudaFuncSetAttribute(sinkhorn_regroup_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 101 * 1024));
            rqb::count_launch();
            sinkhorn_regroup_warp_kernel<<<(unsigned)((n_groups + wpc - 1) / wpc), 128, per_warp * wpc, (cudaStream_t)stream>>>(
                residual_dev, items_dev, offsets_dev, n_groups, e, m->cb[L - 1], m->cc[L - 1], K, L, epsilon, iters, wpc, codes_dev);
            lo = SKW_MAX_ROWS + 1;
        }
    }
    for (int c = 0; c < 4 && lo <= top; ++c) {
        const int hi = bounds[c] < top ? bounds[c] : top;
        if (hi < lo) continue;
        const size_t smem = per_row * hi + 16;
        rqb::count_launch();
        sinkhorn_regroup_kernel<<<(unsigned)n_groups, SK_THREADS, smem, (cudaStream_t)stream>>>(
            residual_dev, items_dev, offsets_dev, e, m->cb[L - 1], m->cc[L - 1], K, L, lo, hi, epsilon, iters, codes_dev);
        lo = hi + 1;
    }
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_sinkhorn_regroup_large(rqb200_model *m, const float *residual_dev, const int64_t *items_dev,
                                             const int64_t *offsets_dev, const int64_t *group_ids_dev,
                                             const int64_t *scratch_off_dev, int64_t n_listed, double *scratch_dev,
                                             double epsilon, int iters, int64_t *codes_dev, void *stream) {
    RQB_CHECK(m != nullptr, "model is NULL");
    if (n_listed == 0) return 0;
    RQB_CHECK(residual_dev && items_dev && offsets_dev && group_ids_dev && scratch_off_dev && scratch_dev && codes_dev, "NULL buffer");
    RQB_CHECK(epsilon > 0.0, "epsilon must be > 0");
    RQB_CHECK(m->e < 512, "e_dim >= 512 not supported");
    RQB_CUDA(cudaSetDevice(m->device));
    const int L = m->L, K = m->K[L - 1], e = m->e;
    RQB_CHECK(m->cb_set[L - 1], "codebook %d not loaded", L - 1);
    rqb::count_launch();
    ProfScope ps(PROF_SINKHORN, (cudaStream_t)stream);
    sinkhorn_regroup_large_kernel<<<(unsigned)n_listed, SK_THREADS, 0, (cudaStream_t)stream>>>(
        residual_dev, items_dev, offsets_dev, group_ids_dev, scratch_off_dev, scratch_dev, e, m->cb[L - 1], m->cc[L - 1], K, L,
        epsilon, iters, codes_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_sinkhorn_group_cap(rqb200_model *m) {
    if (!m) return 0;
    const int L = m->L, K = m->K[L - 1], e = m->e;
    return (int)((200 * 1024) / (sizeof(double) * K + sizeof(float) * (e + 1)));
}

extern "C" int rqb200_sinkhorn(double *Q_dev, int64_t B, int K, double epsilon, int iters, void *stream) {
    if (B == 0) return 0;
    RQB_CHECK(Q_dev != nullptr, "NULL buffer");
    RQB_CHECK(epsilon != 0.0, "epsilon must be non-zero");
    RQB_CHECK(B * (int64_t)K < ((int64_t)1 << 31), "matrix too large");
    rqb::count_launch();
    sinkhorn_global_kernel<<<1, SK_THREADS, 0, (cudaStream_t)stream>>>(Q_dev, (int)B, K, epsilon, iters);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_sinkhorn_assign(const float *d_dev, int64_t B, int K, double epsilon, int iters,
                                      double *scratch_dev, int64_t *idx_dev, void *stream) {
    if (B == 0) return 0;
    RQB_CHECK(d_dev && idx_dev && scratch_dev, "NULL buffer");
    RQB_CHECK(epsilon > 0.0, "epsilon must be > 0");
    RQB_CHECK(B * (int64_t)K < ((int64_t)1 << 31), "matrix too large");
    cudaStream_t s = (cudaStream_t)stream;
    // whole-GPU variant when the rows fit in the shared memory of <= 148 CTAs and the caller's [B,K] scratch holds the
    // exchange area; tiny or huge batches keep the one-CTA kernel
    const size_t row_bytes = sizeof(double) * (size_t)K;
    const int cap_rows = (int)((200 * 1024 - sizeof(double) * (size_t)K) / (row_bytes + sizeof(double)));
    int G = (int)((B + 7) / 8);
    if (G > kNumSMs) G = kNumSMs;
    static int max_ctas = -1;              // RQB200_SINKHORN_CTAS: cap on the CTAs (A/B of barrier cost vs rows per CTA)
    if (max_ctas < 0) {
        const char *e = getenv("RQB200_SINKHORN_CTAS");
        max_ctas = e ? atoi(e) : 0;
    }
    if (max_ctas > 0 && G > max_ctas) {
        G = max_ctas;
        const int need = (int)((B + cap_rows - 1) / cap_rows);
        if (G < need) G = need;
        if (G > kNumSMs) G = kNumSMs;
    }
    const int rows_per_cta = (int)((B + G - 1) / G);
    G = (int)((B + rows_per_cta - 1) / rows_per_cta);
    const size_t ws_doubles = 8 + (size_t)4 * G + (size_t)2 * G * K + (size_t)2 * K;
    static int use_grid = -1;
    if (use_grid < 0) {
        const char *e = getenv("RQB200_SINKHORN_GRID");
        use_grid = (e && e[0] == '0') ? 0 : 1;
    }
    // Below 1024 rows (collision groups, the reference's batch of 64) the one-CTA kernel stays: it keeps every division of
    // layers.py:96-104 and reproduces all of the reference's golden use_sk cases; the scaling-vector form picked another
    // code in 1 of their 840 rows (a 1e-13 near-tie), which is inside the training tolerance but not index parity.
    if (use_grid && B >= 1024 && G >= 2 && rows_per_cta <= cap_rows && ws_doubles <= (size_t)B * K) {
        static rqb::DeviceOnce attr_once;
        if (attr_once.first()) {
            RQB_CUDA(cudaFuncSetAttribute(sinkhorn_assign_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
        }
        RQB_CUDA(cudaMemsetAsync(scratch_dev, 0, 64, s));
        int Bi = (int)B, rpc = rows_per_cta;
        void *args[] = {(void *)&d_dev, (void *)&scratch_dev, (void *)&Bi, (void *)&K, (void *)&rpc, (void *)&epsilon,
                        (void *)&iters, (void *)&idx_dev};
        rqb::count_launch();
        RQB_CUDA(cudaLaunchCooperativeKernel((void *)sinkhorn_assign_grid_kernel, dim3((unsigned)G), dim3(SK_THREADS), args,
                                             (row_bytes + sizeof(double)) * rows_per_cta + sizeof(double) * (size_t)K + 16, s));
        return 0;
    }
    rqb::count_launch();
    sinkhorn_assign_kernel<<<1, SK_THREADS, 0, s>>>(d_dev, scratch_dev, (int)B, K, epsilon, iters, idx_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

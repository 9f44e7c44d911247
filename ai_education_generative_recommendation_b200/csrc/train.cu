// train.cu — the training step of the RQ-VAE (reference RQ-VAE/train.py:97-124):
// forward pieces that only exist in training (dropout, per-level apply with losses), the backward of the MLPs
// (layers.py:18-43), of the straight-through quantizer (vq.py:90-95, rq.py:43-54) and of the reconstruction loss
// (rqvae.py:73-84), and a fused clip_grad_norm_ + AdamW step over all parameters (train.py:116-118, :75-78).
//
// Forward Linear layers reuse linear_exact (the reference's fp32 summation order); the backward contractions have no
// order to reproduce (the reference's autograd runs whatever sgemm the backend picks), so they are plain fp32 FMA
// register-tiled GEMMs with a deterministic split over the batch dimension.  Parity target: 1e-4 relative (north_star).
#include <math.h>

#include "common.cuh"

namespace rqb {
namespace {

// ---------------------------------------------------------------------------------------------- dropout
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// y = x * keep / (1 - p); keep(i) is a pure function of (seed, i) so the backward pass regenerates the mask instead
// of storing it.  One hash yields four 16-bit uniforms → four neighbouring elements.
__global__ void dropout_kernel(const float *__restrict__ x, int64_t count, uint32_t thresh16, float scale, uint64_t seed,
                               const uint64_t *__restrict__ seed_dev, float *__restrict__ y) {
    if (seed_dev) seed ^= mix64(*seed_dev);      // replayed CUDA graphs: the per-step part of the seed lives on the device
    const int64_t q0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = q0; q * 4 < count; q += stride) {
        const uint64_t h = mix64(seed ^ mix64((uint64_t)q));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = q * 4 + j;
            if (i < count) {
                const uint32_t u = (uint32_t)(h >> (16 * j)) & 0xFFFFu;
                y[i] = u >= thresh16 ? x[i] * scale : 0.0f;
            }
        }
    }
}

// dy *= (y > 0)
__global__ void relu_mask_kernel(const float *__restrict__ y, int64_t count, float *__restrict__ dy) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        if (!(y[i] > 0.0f)) dy[i] = 0.0f;
}

// ---------------------------------------------------------------------------------------------- generic fp32 GEMM
// C[m, n] = sum_k A(m, k) * B(k, n) with A(m, k) = A[m * sam + k * sak], B(k, n) = B[k * sbk + n * sbn].
// grid (ceil(N/64), ceil(M/64), S): slice z covers k in [z * klen, (z+1) * klen) and writes its own [M, N] partial.
constexpr int GB = 64, GK = 16, GT = 256;

__global__ void __launch_bounds__(GT)
sgemm_strided_kernel(const float *__restrict__ A, int64_t sam, int64_t sak, const float *__restrict__ B, int64_t sbk,
                     int64_t sbn, float *__restrict__ C, int M, int N, int K, int klen) {
    __shared__ float As[GK][GB + 4];
    __shared__ float Bs[GK][GB + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
    const int kbeg = blockIdx.z * klen;
    const int kend = min(K, kbeg + klen);
    float *Cz = C + (size_t)blockIdx.z * M * N;
    const bool a_kfast = sak == 1, b_nfast = sbn == 1;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
        for (int t = 0; t < GB * GK / GT; ++t) {
            const int eidx = tid + t * GT;
            int mm, kk;
            if (a_kfast) { kk = eidx % GK; mm = eidx / GK; } else { mm = eidx % GB; kk = eidx / GB; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < kend) ? A[(int64_t)m * sam + (int64_t)k * sak] : 0.0f;
            int nn, kb;
            if (b_nfast) { nn = eidx % GB; kb = eidx / GB; } else { kb = eidx % GK; nn = eidx / GK; }
            const int n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < N && k2 < kend) ? B[(int64_t)k2 * sbk + (int64_t)n * sbn] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) Cz[(size_t)m * N + n] = acc[i][j];
        }
    }
}

// out[i] = part[0][i] + part[1][i] + … (ascending slice order: deterministic)
__global__ void reduce_slices_kernel(const float *__restrict__ part, int S, int64_t count, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = part[i];
    for (int z = 1; z < S; ++z) s += part[(size_t)z * count + i];
    out[i] = s;
}

// db[j] = sum_r dy[r, j]; 32 columns per CTA, 32 row lanes per column (1024 threads), fixed combination order
__global__ void __launch_bounds__(1024) colsum_kernel(const float *__restrict__ dy, int64_t n, int out, float *__restrict__ db) {
    __shared__ float part[32][33];
    const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + c;
    float s = 0.0f;
    if (j < out)
        for (int64_t r = rl; r < n; r += 32) s += dy[r * out + j];
    part[rl][c] = s;
    __syncthreads();
    if (rl == 0 && j < out) {
        float t = part[0][c];
#pragma unroll
        for (int q = 1; q < 32; ++q) t += part[q][c];
        db[j] = t;
    }
}

int sgemm(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn, float *C, int M, int N, int K,
          float *scratch, size_t scratch_floats, cudaStream_t s) {
    if (M == 0 || N == 0) return 0;
    const int tiles = ((M + GB - 1) / GB) * ((N + GB - 1) / GB);
    int S = 1;
    if (tiles < 2 * kNumSMs && K > 4 * GK) {
        S = (2 * kNumSMs + tiles - 1) / tiles;
        const int maxS = (K + 4 * GK - 1) / (4 * GK);
        if (S > maxS) S = maxS;
        while (S > 1 && (size_t)S * M * N > scratch_floats) --S;
    }
    int klen = (K + S - 1) / S;
    klen = (klen + GK - 1) / GK * GK;
    S = (K + klen - 1) / klen;
    if (S < 1) S = 1;
    dim3 grid((unsigned)((N + GB - 1) / GB), (unsigned)((M + GB - 1) / GB), (unsigned)S);
    count_launch();
    sgemm_strided_kernel<<<grid, GT, 0, s>>>(A, sam, sak, B, sbk, sbn, S > 1 ? scratch : C, M, N, K, klen);
    RQB_LAUNCH_CHECK();
    if (S > 1) {
        const int64_t count = (int64_t)M * N;
        count_launch();
        reduce_slices_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(scratch, S, count, C);
        RQB_LAUNCH_CHECK();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- quantizer pieces
// One level of rq.py:43-54 given the chosen codes: q = E[idx]; sumsq += Σ (q - r)²; x_res = r + (q - r) (vq.py:95);
// r_next = r - x_res; x_q = (first ? 0 : x_q) + x_res.
__global__ void rq_level_apply_kernel(const float *__restrict__ r, const int64_t *__restrict__ idx, const float *__restrict__ cb,
                                      int64_t n, int e, int first, float *__restrict__ xq, float *__restrict__ r_next,
                                      double *__restrict__ sumsq) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double local = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n * e; p += stride) {
        const int64_t i = p / e;
        const int k = (int)(p - i * e);
        const float q = cb[idx[i] * e + k];
        const float rv = r[p];
        const float diff = __fsub_rn(q, rv);
        const float xres = __fadd_rn(rv, diff);
        r_next[p] = __fsub_rn(rv, xres);
        xq[p] = first ? __fadd_rn(0.0f, xres) : __fadd_rn(xq[p], xres);
        local += (double)__fmul_rn(diff, diff);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = local;
    __syncthreads();
    if (w == 0) {
        double t = lane < (blockDim.x >> 5) ? wsum[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && sumsq) atomicAdd(sumsq, t);
    }
}

// dE[code, k] = coef * g[0] * Σ_{i : idx[i] == code} (E[code, k] - r[i, k]), items in ascending order (deterministic).
// One CTA per code; the index scan is split over the warps of the CTA and combined in warp order.
__global__ void __launch_bounds__(256) vq_codebook_grad_kernel(const float *__restrict__ r, const int64_t *__restrict__ idx,
                                                              const float *__restrict__ cb, int64_t n, int e, float coef,
                                                              const float *__restrict__ g, float *__restrict__ dE) {
    extern __shared__ float sacc[];            // [8 warps][e]
    const int code = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t per = (n + 7) / 8;
    const int64_t i0 = w * per, i1 = min(n, i0 + per);
    for (int k0 = 0; k0 < e; k0 += 32) {
        const int k = k0 + lane;
        float acc = 0.0f;
        const float q = k < e ? cb[(int64_t)code * e + k] : 0.0f;
        for (int64_t i = i0; i < i1; ++i)
            if (idx[i] == code && k < e) acc += q - r[i * e + k];
        if (k < e) sacc[w * e + k] = acc;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < e; k += blockDim.x) {
        float t = sacc[k];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += sacc[q * e + k];
        dE[(int64_t)code * e + k] = coef * g[0] * t;
    }
}

// dz = g_xq + coef * g[0] * (z - E0[idx0])      (commitment term of level 0; deeper levels cancel)
__global__ void rq_latent_grad_kernel(const float *__restrict__ z, const int64_t *__restrict__ idx0, const float *__restrict__ cb0,
                                      const float *__restrict__ gxq, int64_t n, int e, float coef, const float *__restrict__ g,
                                      float *__restrict__ dz) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float c = coef * g[0];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n * e; p += stride) {
        const int64_t i = p / e;
        const int k = (int)(p - i * e);
        dz[p] = (gxq ? gxq[p] : 0.0f) + c * (z[p] - cb0[idx0[i] * e + k]);
    }
}

// d_out = g[0] * scale * (out - x)   (mse, scale = 2 / count)   |   g[0] * scale * sign(out - x)   (l1, scale = 1 / count)
__global__ void recon_grad_kernel(const float *__restrict__ out, const float *__restrict__ x, int64_t count, int l1, float scale,
                                  const float *__restrict__ g, float *__restrict__ d_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float c = g[0] * scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float d = out[i] - x[i];
        d_out[i] = l1 ? (d > 0.0f ? c : (d < 0.0f ? -c : 0.0f)) : c * d;
    }
}

// ---------------------------------------------------------------------------------------------- clip + AdamW
// chunk table row: {param ptr, grad ptr, exp_avg ptr, exp_avg_sq ptr, element count}
constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) gradnorm_chunks_kernel(const int64_t *__restrict__ chunks, float grad_scale,
                                                                      double *__restrict__ partial) {
    const int64_t *c = chunks + (int64_t)blockIdx.x * 5;
    const float *g = reinterpret_cast<const float *>(c[1]);
    const int64_t n = c[4];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += OPT_THREADS) {
        const double v = (double)(g[i] * grad_scale);
        s += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[OPT_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < OPT_THREADS / 32; ++q) t += ws[q];
        partial[blockIdx.x] = t;
    }
}

// stats[0] = total L2 norm of the (scaled) gradients, stats[1] = clip coefficient min(1, max_norm / (norm + 1e-6))
// (torch.nn.utils.clip_grad_norm_, train.py:116); max_norm <= 0 disables clipping.
__global__ void gradnorm_finish_kernel(const double *__restrict__ partial, int n_chunks, float max_norm, float *__restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double t = 0.0;
    for (int i = 0; i < n_chunks; ++i) t += partial[i];
    const float norm = (float)sqrt(t);
    float coef = 1.0f;
    if (max_norm > 0.0f) {
        coef = max_norm / (norm + 1e-6f);
        coef = coef > 1.0f ? 1.0f : coef;
    }
    stats[0] = norm;
    stats[1] = coef;
}

// torch.optim.AdamW (decoupled weight decay, no amsgrad), one step:
//   p *= decay = 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g²;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(OPT_THREADS) adamw_chunks_kernel(const int64_t *__restrict__ chunks, const float *__restrict__ stats,
                                                                   float grad_scale, float decay, float b1, float b2, float eps,
                                                                   float step_size, float bc2_sqrt) {
    const int64_t *c = chunks + (int64_t)blockIdx.x * 5;
    float *p = reinterpret_cast<float *>(c[0]);
    const float *g = reinterpret_cast<const float *>(c[1]);
    float *m = reinterpret_cast<float *>(c[2]);
    float *v = reinterpret_cast<float *>(c[3]);
    const int64_t n = c[4];
    const float coef = stats[1] * grad_scale;
    for (int64_t i = threadIdx.x; i < n; i += OPT_THREADS) {
        const float gi = g[i] * coef;
        float pi = p[i] * decay;
        const float mi = m[i] + (gi - m[i]) * (1.0f - b1);           // lerp, as torch does
        const float vi = v[i] * b2 + gi * gi * (1.0f - b2);
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// the same update with the per-step scalars read from device memory: hyper = {decay, step_size, sqrt(bc2), b1, b2, eps}
// (a captured CUDA graph replays the launch; the host refreshes the six floats before every replay)
__global__ void __launch_bounds__(OPT_THREADS) adamw_chunks_dev_kernel(const int64_t *__restrict__ chunks, const float *__restrict__ stats,
                                                                       float grad_scale, const float *__restrict__ hyper) {
    const int64_t *c = chunks + (int64_t)blockIdx.x * 5;
    float *p = reinterpret_cast<float *>(c[0]);
    const float *g = reinterpret_cast<const float *>(c[1]);
    float *m = reinterpret_cast<float *>(c[2]);
    float *v = reinterpret_cast<float *>(c[3]);
    const int64_t n = c[4];
    const float coef = stats[1] * grad_scale;
    const float decay = hyper[0], step_size = hyper[1], bc2_sqrt = hyper[2], b1 = hyper[3], b2 = hyper[4], eps = hyper[5];
    for (int64_t i = threadIdx.x; i < n; i += OPT_THREADS) {
        const float gi = g[i] * coef;
        float pi = p[i] * decay;
        const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
        const float vi = v[i] * b2 + gi * gi * (1.0f - b2);
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

struct Floats8 { float v[8]; };
// the values travel as launch arguments (copied at launch time), so the host may run many steps ahead of the device
__global__ void set_floats_kernel(float *__restrict__ dst, int n, Floats8 vals) {
    if (threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}

int grid_for(int64_t count) {
    int64_t b = (count + 255) / 256;
    if (b > kNumSMs * 8) b = kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

int default_kblocks_public(int K, int *out);       // api.cu

}  // namespace rqb

using namespace rqb;

extern "C" int rqb200_dropout(const float *x_dev, int64_t count, float p, uint64_t seed, float *y_dev, void *stream) {
    return rqb200_dropout_dev(x_dev, count, p, seed, nullptr, y_dev, stream);
}

extern "C" int rqb200_dropout_dev(const float *x_dev, int64_t count, float p, uint64_t seed, const uint64_t *seed_dev,
                                  float *y_dev, void *stream) {
    if (count == 0) return 0;
    RQB_CHECK(x_dev && y_dev, "NULL buffer");
    RQB_CHECK(p >= 0.0f && p < 1.0f, "dropout probability must be in [0, 1)");
    const uint32_t thresh = (uint32_t)lrintf(p * 65536.0f);
    count_launch();
    dropout_kernel<<<grid_for((count + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x_dev, count, thresh, 1.0f / (1.0f - p), seed, seed_dev, y_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_linear_forward(const float *x_dev, const float *W_dev, const float *b_dev, int64_t n, int in_dim,
                                     int out_dim, int relu, float *y_dev, void *stream) {
    if (n == 0) return 0;
    RQB_CHECK(x_dev && W_dev && b_dev && y_dev, "NULL buffer");
    Linear lin;
    lin.in = in_dim; lin.out = out_dim;
    lin.W = const_cast<float *>(W_dev); lin.b = const_cast<float *>(b_dev);
    lin.nblk = default_kblocks_public(in_dim, lin.kblocks);
    RQB_CHECK(lin.nblk >= 1, "in_features %d not supported", in_dim);
    lin.set = true;
    return linear_exact(lin, x_dev, nullptr, n, y_dev, relu != 0, (cudaStream_t)stream);
}

extern "C" int64_t rqb200_linear_backward_scratch_floats(int64_t n, int in_dim, int out_dim) {
    (void)n;
    return (int64_t)in_dim * out_dim * (2 * kNumSMs);     // upper bound on slices x [out, in]; sgemm() clamps to what it gets
}

extern "C" int rqb200_linear_backward(const float *x_dev, const float *W_dev, const float *y_dev, float *dy_dev, int64_t n,
                                      int in_dim, int out_dim, int relu, float *dx_dev, float *dW_dev, float *db_dev,
                                      float *scratch_dev, int64_t scratch_floats, void *stream) {
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    RQB_CHECK(x_dev && W_dev && dy_dev && dW_dev && db_dev, "NULL buffer");
    RQB_CHECK(n < ((int64_t)1 << 31), "batch too large");
    if (relu) {
        RQB_CHECK(y_dev != nullptr, "the ReLU mask needs the layer output");
        count_launch();
        relu_mask_kernel<<<grid_for(n * out_dim), 256, 0, s>>>(y_dev, n * out_dim, dy_dev);
        RQB_LAUNCH_CHECK();
    }
    // dW[out, in] = dyᵀ · x : A(m = o, k = r) = dy[r * out + o], B(k = r, n = i) = x[r * in + i]
    RQB_TRY(sgemm(dy_dev, 1, out_dim, x_dev, in_dim, 1, dW_dev, out_dim, in_dim, (int)n, scratch_dev,
                  scratch_dev ? (size_t)scratch_floats : 0, s));
    count_launch();
    colsum_kernel<<<(out_dim + 31) / 32, 1024, 0, s>>>(dy_dev, n, out_dim, db_dev);
    RQB_LAUNCH_CHECK();
    // dx[n, in] = dy · W : A(m = r, k = o) = dy[r * out + o], B(k = o, n = i) = W[o * in + i]
    if (dx_dev)
        RQB_TRY(sgemm(dy_dev, out_dim, 1, W_dev, in_dim, 1, dx_dev, (int)n, in_dim, out_dim, nullptr, 0, s));
    return 0;
}

extern "C" int rqb200_rq_level_apply(const float *r_dev, const int64_t *idx_dev, const float *cb_dev, int64_t n, int e,
                                     int first, float *xq_dev, float *r_next_dev, double *sumsq_dev, void *stream) {
    if (n == 0) return 0;
    RQB_CHECK(r_dev && idx_dev && cb_dev && xq_dev && r_next_dev, "NULL buffer");
    count_launch();
    rq_level_apply_kernel<<<grid_for(n * e), 256, 0, (cudaStream_t)stream>>>(r_dev, idx_dev, cb_dev, n, e, first, xq_dev,
                                                                           r_next_dev, sumsq_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_vq_codebook_grad(const float *r_dev, const int64_t *idx_dev, const float *cb_dev, int64_t n, int e, int K,
                                       float coef, const float *g_dev, float *dE_dev, void *stream) {
    RQB_CHECK(r_dev && idx_dev && cb_dev && g_dev && dE_dev, "NULL buffer");
    RQB_CHECK(e <= 1024, "e_dim too large");
    count_launch();
    vq_codebook_grad_kernel<<<K, 256, sizeof(float) * 8 * e, (cudaStream_t)stream>>>(r_dev, idx_dev, cb_dev, n, e, coef, g_dev,
                                                                                    dE_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_rq_latent_grad(const float *z_dev, const int64_t *idx0_dev, const float *cb0_dev, const float *gxq_dev,
                                     int64_t n, int e, float coef, const float *g_dev, float *dz_dev, void *stream) {
    if (n == 0) return 0;
    RQB_CHECK(z_dev && idx0_dev && cb0_dev && g_dev && dz_dev, "NULL buffer");
    count_launch();
    rq_latent_grad_kernel<<<grid_for(n * e), 256, 0, (cudaStream_t)stream>>>(z_dev, idx0_dev, cb0_dev, gxq_dev, n, e, coef, g_dev,
                                                                           dz_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_recon_loss(const float *out_dev, const float *x_dev, int64_t count, double *sums2_dev, void *stream) {
    RQB_CHECK(out_dev && x_dev && sums2_dev, "NULL buffer");
    return recon_error(out_dev, x_dev, count, sums2_dev, (cudaStream_t)stream);
}

extern "C" int rqb200_recon_grad(const float *out_dev, const float *x_dev, int64_t count, int l1, const float *g_dev,
                                 float *d_out_dev, void *stream) {
    if (count == 0) return 0;
    RQB_CHECK(out_dev && x_dev && g_dev && d_out_dev, "NULL buffer");
    count_launch();
    recon_grad_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(out_dev, x_dev, count, l1, (l1 ? 1.0f : 2.0f) / (float)count,
                                                                       g_dev, d_out_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_adamw_clip_step(const int64_t *chunks_dev, int n_chunks, double *partial_dev, float *stats_dev,
                                      float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                                      float weight_decay, int64_t step, void *stream) {
    if (n_chunks == 0) return 0;
    RQB_CHECK(chunks_dev && partial_dev && stats_dev, "NULL buffer");
    RQB_CHECK(step >= 1, "step counts from 1");
    cudaStream_t s = (cudaStream_t)stream;
    count_launch();
    gradnorm_chunks_kernel<<<n_chunks, OPT_THREADS, 0, s>>>(chunks_dev, grad_scale, partial_dev);
    count_launch();
    gradnorm_finish_kernel<<<1, 32, 0, s>>>(partial_dev, n_chunks, max_norm, stats_dev);
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    count_launch();
    adamw_chunks_kernel<<<n_chunks, OPT_THREADS, 0, s>>>(chunks_dev, stats_dev, grad_scale,
                                                        (float)(1.0 - (double)lr * (double)weight_decay), beta1, beta2, eps,
                                                        (float)((double)lr / bc1), (float)sqrt(bc2));
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_adamw_clip_step_dev(const int64_t *chunks_dev, int n_chunks, double *partial_dev, float *stats_dev,
                                          float grad_scale, float max_norm, const float *hyper_dev, void *stream) {
    if (n_chunks == 0) return 0;
    RQB_CHECK(chunks_dev && partial_dev && stats_dev && hyper_dev, "NULL buffer");
    cudaStream_t s = (cudaStream_t)stream;
    count_launch();
    gradnorm_chunks_kernel<<<n_chunks, OPT_THREADS, 0, s>>>(chunks_dev, grad_scale, partial_dev);
    count_launch();
    gradnorm_finish_kernel<<<1, 32, 0, s>>>(partial_dev, n_chunks, max_norm, stats_dev);
    count_launch();
    adamw_chunks_dev_kernel<<<n_chunks, OPT_THREADS, 0, s>>>(chunks_dev, stats_dev, grad_scale, hyper_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_set_floats(float *dst_dev, int n, const float *values_host, void *stream) {
    RQB_CHECK(dst_dev && values_host && n >= 1 && n <= 8, "1..8 floats");
    Floats8 f;
    for (int i = 0; i < 8; ++i) f.v[i] = i < n ? values_host[i] : 0.0f;
    count_launch();
    set_floats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dst_dev, n, f);
    RQB_LAUNCH_CHECK();
    return 0;
}

// kmeans.cu — Lloyd steps of the k-means codebook initialisation, on device.
//
// Replaces `kmeans(samples, num_clusters, num_iters)` (reference RQ-VAE/models/layers.py:69-82, called
// from VectorQuantizer.init_emb, reference RQ-VAE/models/vq.py:40-49), which copies the batch to the
// host and runs scikit-learn there.  Here a Lloyd iteration is: nearest-centre assignment (the
// quantizer's distance/argmin kernel), per-cluster sum/count accumulation (this file) and the centre
// update.  The [K,e] sums and [K] counts are exactly the statistics that are all-reduced over
// NCCL/NVLink when the samples are sharded across GPUs.
// Bound: assignment = fp32 FMA pipe; accumulation = HBM read of the samples (4e B per sample).
#include "common.cuh"

namespace rqb {
namespace {

constexpr int KM_THREADS = 256;

// sums[code[i]][k] += x[i][k]; counts[code[i]] += 1; inertia += ||x_i - c_code||^2   (fp64)
// Shared-memory privatised when K*e doubles fit, flushed once per CTA.
__global__ void __launch_bounds__(KM_THREADS)
kmeans_accumulate_kernel(const float *__restrict__ x, int64_t n, int e, const int64_t *__restrict__ code,
                         const float *__restrict__ centers, int K, double *__restrict__ sums,
                         unsigned long long *__restrict__ counts, double *__restrict__ inertia,
                         int use_smem) {
    extern __shared__ __align__(16) double s_sum[];
    __shared__ double s_in[KM_THREADS / 32];
    const int tid = threadIdx.x;
    if (use_smem) {
        for (int i = tid; i < K * e; i += KM_THREADS) s_sum[i] = 0.0;
        __syncthreads();
    }
    double my_inertia = 0.0;
    const int64_t total = n * (int64_t)e;
    const int64_t per_cta = ((n + gridDim.x - 1) / gridDim.x) * e;      // whole rows per CTA
    const int64_t lo = (int64_t)blockIdx.x * per_cta;
    const int64_t hi = lo + per_cta < total ? lo + per_cta : total;
    for (int64_t p = lo + tid; p < hi; p += KM_THREADS) {
        const int64_t row = p / e;
        const int k = (int)(p - row * e);
        const int c = (int)code[row];
        const float v = x[p];
        const double diff = (double)v - (double)centers[(int64_t)c * e + k];
        my_inertia += diff * diff;
        if (use_smem) atomicAdd(&s_sum[c * e + k], (double)v);
        else atomicAdd(&sums[(int64_t)c * e + k], (double)v);
        if (k == 0) atomicAdd(&counts[c], 1ull);
    }
    for (int o = 16; o > 0; o >>= 1) my_inertia += __shfl_xor_sync(0xffffffffu, my_inertia, o);
    if ((tid & 31) == 0) s_in[tid >> 5] = my_inertia;
    __syncthreads();
    if (tid == 0 && inertia) {
        double t = 0.0;
        for (int w = 0; w < KM_THREADS / 32; ++w) t += s_in[w];
        atomicAdd(inertia, t);
    }
    if (use_smem) {
        for (int i = tid; i < K * e; i += KM_THREADS) {
            double v = s_sum[i];
            if (v != 0.0) atomicAdd(&sums[i], v);
        }
    }
}

__global__ void kmeans_update_kernel(float *__restrict__ centers, int K, int e, const double *__restrict__ sums,
                                     const unsigned long long *__restrict__ counts, double *__restrict__ shift) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double d2 = 0.0;
    if (i < K * e) {
        unsigned long long c = counts[i / e];
        if (c > 0) {                                   // empty clusters keep their centre
            float nv = (float)(sums[i] / (double)c);
            double d = (double)nv - (double)centers[i];
            d2 = d * d;
            centers[i] = nv;
        }
    }
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((threadIdx.x & 31) == 0 && shift && d2 != 0.0) atomicAdd(shift, d2);
}

}  // namespace
}  // namespace rqb

using namespace rqb;

extern "C" int rqb200_kmeans_assign(const float *x_dev, int64_t n, int e, const float *centers_dev, int K,
                                    float *cnorm_scratch_dev, int64_t *assign_dev, void *stream) {
    // nearest centre with the quantizer's exact distance / first-index argmin
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return 0;
    RQB_CHECK(x_dev && centers_dev && cnorm_scratch_dev && assign_dev, "NULL buffer");
    rqb200_model tmp;
    tmp.L = 1; tmp.e = e; tmp.K[0] = K;
    tmp.cb[0] = const_cast<float *>(centers_dev);
    tmp.cc[0] = cnorm_scratch_dev;
    RQB_TRY(codebook_norms(centers_dev, K, e, cnorm_scratch_dev, s));
    return quantize_exact(&tmp, x_dev, n, assign_dev, nullptr, nullptr, nullptr, nullptr, nullptr, s, (int64_t)1 << 30);
}

extern "C" int rqb200_kmeans_distances(const float *x_dev, int64_t n, int e, const float *centers_dev, int K,
                                       float *cnorm_scratch_dev, float *d_dev, void *stream) {
    // d[n, K] = squared distances to K candidate centres (used by the k-means++ seeding), quantizer arithmetic
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return 0;
    RQB_CHECK(x_dev && centers_dev && cnorm_scratch_dev && d_dev, "NULL buffer");
    rqb200_model tmp;
    tmp.L = 1; tmp.e = e; tmp.K[0] = K;
    tmp.cb[0] = const_cast<float *>(centers_dev);
    tmp.cc[0] = cnorm_scratch_dev;
    RQB_TRY(codebook_norms(centers_dev, K, e, cnorm_scratch_dev, s));
    return distances_exact(&tmp, 0, x_dev, n, d_dev, s);
}

extern "C" int rqb200_kmeans_accumulate(const float *x_dev, int64_t n, int e, const int64_t *assign_dev,
                                        const float *centers_dev, int K, double *sums_dev,
                                        int64_t *counts_dev, double *inertia_dev, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return 0;
    RQB_CHECK(x_dev && assign_dev && centers_dev && sums_dev && counts_dev, "NULL buffer");
    const size_t smem = sizeof(double) * (size_t)K * e;
    const int use_smem = smem <= 160 * 1024;
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kmeans_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      160 * 1024));
    }
    int64_t grid = (n + 1023) / 1024;
    if (grid > kNumSMs) grid = kNumSMs;
    rqb::count_launch();
    kmeans_accumulate_kernel<<<(unsigned)grid, KM_THREADS, use_smem ? smem : 0, s>>>(
        x_dev, n, e, assign_dev, centers_dev, K, sums_dev, (unsigned long long *)counts_dev, inertia_dev, use_smem);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_kmeans_update(float *centers_dev, int K, int e, const double *sums_dev,
                                    const int64_t *counts_dev, double *shift_dev, void *stream) {
    RQB_CHECK(centers_dev && sums_dev && counts_dev, "NULL buffer");
    rqb::count_launch();
    kmeans_update_kernel<<<(K * e + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        centers_dev, K, e, sums_dev, (const unsigned long long *)counts_dev, shift_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

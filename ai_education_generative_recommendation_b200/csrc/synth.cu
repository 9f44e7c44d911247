// synth.cu — synthetic item catalogue for bench / parity runs (SURVEY.md §8d "Synthetic inputs").
//
// Clustered BERT-like embeddings from pure integer hashing: every value is a function of
// (seed, row, column) only, built from exact integer sums and single IEEE roundings, so the numpy
// twin (package synth.py) produces identical bytes on the CPU for any row range.  Row 0 is the
// all-zero padding row of the reference catalogue; 1 item in 1000 is an exact duplicate of an
// earlier item (exercises collision groups and suffix codes > 0).
#include "common.cuh"

namespace rqb {
namespace {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
    return splitmix64(splitmix64(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ (b * 0xA0761D6478BD642Full));
}
// Irwin–Hall(4) of 16-bit fields, centred and scaled to unit variance
__device__ __forceinline__ float approx_normal(uint64_t bits) {
    uint32_t s = (uint32_t)(bits & 0xFFFF) + (uint32_t)((bits >> 16) & 0xFFFF) +
                 (uint32_t)((bits >> 32) & 0xFFFF) + (uint32_t)(bits >> 48);
    return __fmul_rn(__fsub_rn((float)s, 131070.0f), 2.6429153e-05f);
}

__global__ void synth_items_kernel(uint64_t seed, int64_t first_row, int64_t n, int dim, int64_t n_total,
                                   float *__restrict__ x) {
    const int64_t total = n * (int64_t)dim;
    const uint64_t n_centres = (uint64_t)(n_total / 1000 > 40 ? n_total / 1000 : 40);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lr = p / dim;
        const uint64_t k = (uint64_t)(p - lr * dim);
        uint64_t row = (uint64_t)(first_row + lr);
        float v = 0.0f;
        if (row != 0) {
            if (hash3(seed, 1, row) % 1000ull == 0) row = 1 + hash3(seed, 2, row) % row;   // duplicate of an earlier item
            if (row >= (uint64_t)1) {
                const uint64_t c = hash3(seed, 3, row) % n_centres;
                const float mu = approx_normal(hash3(seed, 0x100000000ull + c, k));
                const float ep = approx_normal(hash3(seed, 0x200000000ull + row, k));
                v = __fadd_rn(__fmul_rn(0.5f, mu), __fmul_rn(0.1f, ep));
            }
        }
        x[p] = v;
    }
}

}  // namespace
}  // namespace rqb

extern "C" int rqb200_synth_items(uint64_t seed, int64_t first_row, int64_t n, int dim, int64_t n_total,
                                  float *x_dev, void *stream) {
    if (n == 0) return 0;
    RQB_CHECK(x_dev != nullptr && dim > 0 && n_total > 0 && first_row >= 0, "bad argument");
    int64_t blocks = (n * (int64_t)dim + 255) / 256;
    if (blocks > rqb::kNumSMs * 16) blocks = rqb::kNumSMs * 16;
    rqb::count_launch();
    rqb::synth_items_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seed, first_row, n, dim, n_total, x_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

// stream_probe.cu — how fast can a B200 stream a row-major [n, 768] fp32 matrix with the tensor-core producers'
// access pattern (128-row tiles read in K slabs) compared with a linear read?  Diagnostic; build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void linear_read(const float4 *x, long long n4, float *out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(x + i);
        s += v.x + v.y + v.z + v.w;
    }
    if (s == 1.2345f) out[0] = s;
}

// CTA = one 128-row tile at a time (persistent), 256 threads; slab = SLAB floats per row; each thread keeps DEPTH slabs in flight
template <int SLAB, int DEPTH>
__global__ void tile_read(const float *x, long long n, int K, float *out) {
    constexpr int F4_PER_ROW = SLAB / 4;                 // float4 per row per slab
    constexpr int ROWS_PER_PASS = 256 / F4_PER_ROW;      // rows covered by one pass of the CTA
    constexpr int NF4 = 128 / ROWS_PER_PASS;             // float4 per thread per slab
    const int c4 = threadIdx.x % F4_PER_ROW, rbase = threadIdx.x / F4_PER_ROW;
    const long long ntiles = n / 128;
    const int KS = K / SLAB;
    float s = 0.f;
    float4 buf[DEPTH][NF4];
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long steps = my_tiles * KS;
    auto load = [&](long long st, float4 (&dst)[NF4]) {
        const long long tile = blockIdx.x + (st / KS) * (long long)gridDim.x;
        const int k0 = (int)(st % KS) * SLAB + c4 * 4;
#pragma unroll
        for (int i = 0; i < NF4; ++i)
            dst[i] = __ldg(reinterpret_cast<const float4 *>(x + (tile * 128 + rbase + ROWS_PER_PASS * i) * (long long)K + k0));
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
        if (d < steps) load(d, buf[d]);
    for (long long st = 0; st < steps; st += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (st + d < steps) {
#pragma unroll
                for (int i = 0; i < NF4; ++i) s += buf[d][i].x + buf[d][i].y + buf[d][i].z + buf[d][i].w;
                if (st + d + DEPTH < steps) load(st + d + DEPTH, buf[d]);
            }
        }
    }
    if (s == 1.2345f) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const long long n = 1000000 / 128 * 128;
    const int K = 768;
    float *x, *out;
    cudaMalloc(&x, n * K * 4);
    cudaMalloc(&out, 4);
    cudaMemset(x, 0, n * K * 4);
    const double gb = n * (double)K * 4 / 1e9;
    float ms = time_ms([&] { linear_read<<<148 * 8, 256>>>((const float4 *)x, n * K / 4, out); });
    printf("linear read                         %7.3f ms  %6.0f GB/s\n", ms, gb / ms * 1e3);
#define RUN(SLAB, DEPTH, GRID)                                                                                   \
    ms = time_ms([&] { tile_read<SLAB, DEPTH><<<GRID, 256>>>(x, n, K, out); });                                  \
    printf("tile read slab %4d B depth %d grid %4d  %7.3f ms  %6.0f GB/s\n", SLAB * 4, DEPTH, GRID, ms, gb / ms * 1e3);
    RUN(64, 1, 148) RUN(64, 2, 148) RUN(64, 3, 148) RUN(64, 4, 148)
    RUN(128, 2, 148) RUN(256, 1, 148) RUN(256, 2, 148)
    RUN(64, 2, 296) RUN(64, 2, 592) RUN(128, 2, 296) RUN(256, 2, 296)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

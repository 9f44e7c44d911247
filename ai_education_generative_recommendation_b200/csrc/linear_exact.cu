// linear_exact.cu — nn.Linear (+ReLU) in the reference's exact fp32 summation order.
//
// Replaces `MLPLayers.forward` layer by layer (reference RQ-VAE/models/layers.py:23,28-30,42-43).
// On the CPU the reference's addmm is an MKL sgemm with beta=1 on a bias-filled output; every
// output element is  ((b_j + chain(blk0)) + chain(blk1)) + ...  where chain(blk) is ONE sequential
// fp32 FMA chain over k ascending inside the K-block (pinned by
// tests/golden).  A register-tiled SIMT SGEMM has exactly that structure as long as each
// accumulator walks k in ascending order and K-blocks are closed with a separate fp32 add —
// which is what this kernel does.  It is the bit-exact anchor of the library and the rescue
// path of the tensor-core encoder.
//
// Tile: BM x BN outputs per CTA, BK=16 slab, 256 threads, TM x TN outputs per thread laid out as
// strided quads so that every shared-memory read is a conflict-free LDS.128.  Global→register
// prefetch of slab s+1 overlaps the FMAs of slab s.  Roofline: fp32 FMA pipe (128 FMA/clk/SM).
#include "common.cuh"

namespace rqb {

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(NTHREADS, (BN >= 128 ? 2 : 3))
linear_exact_kernel(const float *__restrict__ X, const int64_t *__restrict__ rows, int64_t n,
                    const float *__restrict__ W, const float *__restrict__ bias,
                    float *__restrict__ Y, int K, int N, int relu,
                    const int nblk, const int kb0, const int kb1, const int kb2, const int kb3,
                    const int kb4, const int kb5, const int kb6, const int kb7,
                    const unsigned long long *__restrict__ n_dev) {
    // n_dev (may be NULL): the row count lives on the device (rescue tier: rows the margin gate listed); n is then the
    // capacity and the grid is persistent — every CTA walks row tiles blockIdx.x, blockIdx.x + gridDim.x, …
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    constexpr int NTX = BN / TN;           // threads along n
    constexpr int NTY = BM / TM;           // threads along m
    static_assert(NTX * NTY == NTHREADS, "tile/thread mismatch");
    constexpr int QM = TM / 4, QN = TN / 4;
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int A_F4 = BM * BK / 4 / NTHREADS;   // float4 loads per thread for the X slab
    constexpr int B_F4 = (BN * BK / 4 + NTHREADS - 1) / NTHREADS;

    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];
    extern __shared__ __align__(16) float out_stash[];   // [TM*TN][NTHREADS], only when nblk > 1

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int col0 = blockIdx.y * BN;
    const int kbs[8] = {kb0, kb1, kb2, kb3, kb4, kb5, kb6, kb7};
    const int64_t n_row_tiles = (n + BM - 1) / BM;
    for (int64_t row_tile = blockIdx.x; row_tile < n_row_tiles; row_tile += gridDim.x) {
    if (row_tile != (int64_t)blockIdx.x) __syncthreads();         // the previous tile's shared-memory reads are done
    const int64_t row0 = row_tile * BM;

    // ---- per-thread load coordinates
    const float *a_src[A_F4];
    bool a_ok[A_F4];
    int a_row[A_F4], a_kq[A_F4];
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
        int f = tid + i * NTHREADS;          // float4 index in the [BM][BK/4] slab
        a_row[i] = f / (BK / 4);
        a_kq[i] = f % (BK / 4);
        int64_t r = row0 + a_row[i];
        a_ok[i] = r < n;
        int64_t src_row = a_ok[i] ? (rows ? rows[r] : r) : 0;
        a_src[i] = X + src_row * (int64_t)K + a_kq[i] * 4;
    }
    const float *b_src[B_F4];
    bool b_ok[B_F4];
    int b_row[B_F4], b_kq[B_F4];
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
        int f = tid + i * NTHREADS;
        b_row[i] = f / (BK / 4);
        b_kq[i] = f % (BK / 4);
        b_ok[i] = (f < BN * BK / 4) && (col0 + b_row[i] < N);
        b_src[i] = W + (int64_t)(b_ok[i] ? col0 + b_row[i] : 0) * K + b_kq[i] * 4;
    }

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    float4 a_reg[A_F4], b_reg[B_F4];

    // slab loader: k in [k, kend) valid, zero-filled beyond (fma(0,0,acc) == acc)
    // (float4 loads need k % 4 == 0; K-blocks that start off a 4-boundary take the scalar path)
    auto load_slab = [&](int k, int kend) {
        const bool vec = (k & 3) == 0;
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            int kk = k + a_kq[i] * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[i]) {
                if (vec && kk + 3 < kend) v = *reinterpret_cast<const float4 *>(a_src[i] + k);
                else {
                    if (kk + 0 < kend) v.x = a_src[i][k + 0];
                    if (kk + 1 < kend) v.y = a_src[i][k + 1];
                    if (kk + 2 < kend) v.z = a_src[i][k + 2];
                    if (kk + 3 < kend) v.w = a_src[i][k + 3];
                }
            }
            a_reg[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            int kk = k + b_kq[i] * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_ok[i]) {
                if (vec && kk + 3 < kend) v = *reinterpret_cast<const float4 *>(b_src[i] + k);
                else {
                    if (kk + 0 < kend) v.x = b_src[i][k + 0];
                    if (kk + 1 < kend) v.y = b_src[i][k + 1];
                    if (kk + 2 < kend) v.z = b_src[i][k + 2];
                    if (kk + 3 < kend) v.w = b_src[i][k + 3];
                }
            }
            b_reg[i] = v;
        }
    };
    auto store_slab = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            As[buf][a_kq[i] * 4 + 0][a_row[i]] = a_reg[i].x;
            As[buf][a_kq[i] * 4 + 1][a_row[i]] = a_reg[i].y;
            As[buf][a_kq[i] * 4 + 2][a_row[i]] = a_reg[i].z;
            As[buf][a_kq[i] * 4 + 3][a_row[i]] = a_reg[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            if (tid + i * NTHREADS < BN * BK / 4) {
                Bs[buf][b_kq[i] * 4 + 0][b_row[i]] = b_reg[i].x;
                Bs[buf][b_kq[i] * 4 + 1][b_row[i]] = b_reg[i].y;
                Bs[buf][b_kq[i] * 4 + 2][b_row[i]] = b_reg[i].z;
                Bs[buf][b_kq[i] * 4 + 3][b_row[i]] = b_reg[i].w;
            }
        }
    };

    int kbeg = 0;
    for (int blk = 0; blk < nblk; ++blk) {
        const int kend = kbeg + kbs[blk];
        const int nslab = (kend - kbeg + BK - 1) / BK;
        load_slab(kbeg, kend);
        store_slab(0);
        __syncthreads();
        for (int s = 0; s < nslab; ++s) {
            const int buf = s & 1;
            if (s + 1 < nslab) load_slab(kbeg + (s + 1) * BK, kend);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[TM], b[TN];
#pragma unroll
                for (int q = 0; q < QM; ++q) {
                    float4 v = *reinterpret_cast<const float4 *>(&As[buf][k][q * (NTY * 4) + ty * 4]);
                    a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
                }
#pragma unroll
                for (int q = 0; q < QN; ++q) {
                    float4 v = *reinterpret_cast<const float4 *>(&Bs[buf][k][q * (NTX * 4) + tx * 4]);
                    b[q * 4 + 0] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
            }
            if (s + 1 < nslab) store_slab(buf ^ 1);
            __syncthreads();
        }
        // close the K-block: out = (blk == 0 ? bias : out) + acc
        if (nblk > 1) {
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    float *slot = &out_stash[(i * TN + j) * NTHREADS + tid];
                    if (blk == 0) {
                        int c = col0 + (j / 4) * (NTX * 4) + tx * 4 + (j % 4);
                        float bj = (bias && c < N) ? bias[c] : 0.0f;
                        *slot = __fadd_rn(bj, acc[i][j]);
                    } else if (blk + 1 < nblk) {
                        *slot = __fadd_rn(*slot, acc[i][j]);
                    } else {
                        acc[i][j] = __fadd_rn(*slot, acc[i][j]);
                    }
                    if (blk + 1 < nblk) acc[i][j] = 0.0f;
                }
        } else {
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    int c = col0 + (j / 4) * (NTX * 4) + tx * 4 + (j % 4);
                    float bj = (bias && c < N) ? bias[c] : 0.0f;
                    acc[i][j] = __fadd_rn(bj, acc[i][j]);
                }
        }
        kbeg = kend;
    }

    // ---- epilogue: ReLU (NaN propagates like torch.relu) and float4 stores
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t r = row0 + (i / 4) * (NTY * 4) + ty * 4 + (i % 4);
        if (r >= n) continue;
#pragma unroll
        for (int q = 0; q < QN; ++q) {
            int c = col0 + q * (NTX * 4) + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = acc[i][q * 4 + j];
                if (relu) t = (t != t) ? t : (t > 0.0f ? t : 0.0f);
                v[j] = t;
            }
            float *dst = Y + r * (int64_t)N + c;
            if (c + 3 < N && (N % 4 == 0)) {
                *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < N) dst[j] = v[j];
            }
        }
    }
    }   // row tiles
}

template <int BM, int BN, int TM, int TN>
int launch(const Linear &lin, const float *x, const int64_t *rows, int64_t n, float *y, bool relu,
           cudaStream_t s, const unsigned long long *n_dev = nullptr) {
    auto kern = linear_exact_kernel<BM, BN, TM, TN>;
    size_t stash = lin.nblk > 1 ? sizeof(float) * TM * TN * NTHREADS : 0;
    static rqb::DeviceOnce attr_once;   // per template instantiation
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(float) * TM * TN * NTHREADS)));
    }
    int kb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < lin.nblk; ++i) kb[i] = lin.kblocks[i];
    dim3 grid((unsigned)((n + BM - 1) / BM), (unsigned)((lin.out + BN - 1) / BN));
    if (n_dev) {                                       // persistent over the row tiles: about three CTAs per SM in total
        const unsigned cap = (unsigned)((kNumSMs * 3 + grid.y - 1) / grid.y);
        if (grid.x > cap) grid.x = cap;
    }
    rqb::count_launch();
    kern<<<grid, NTHREADS, stash, s>>>(x, rows, n, lin.W, lin.b, y, lin.in, lin.out, relu ? 1 : 0,
                                       lin.nblk, kb[0], kb[1], kb[2], kb[3], kb[4], kb[5], kb[6], kb[7], n_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int linear_exact(const Linear &lin, const float *x, const int64_t *rows, int64_t n, float *y,
                 bool relu, cudaStream_t s, int64_t batch_rows, const unsigned long long *n_dev, int64_t rows_hint) {
    if (n == 0) return 0;
    if (n_dev) {
        // device-counted rows (rescue tier): n is the capacity; always a subset of a large batch (catalogue order)
        RQB_CHECK(lin.set, "linear layer not loaded");
        RQB_CHECK(lin.in % 4 == 0, "in_features must be a multiple of 4 (got %d)", lin.in);
        RQB_CHECK(lin.nblk >= 1 && lin.nblk <= 8, "at most 8 K-blocks supported (got %d)", lin.nblk);
        RQB_CHECK(!small_batch_lane16(batch_rows, lin.in), "device-counted rows must belong to a batch of 16 or more rows");
        // wide layers: 64 x 64 tiles (4 x 4 per thread) keep every SM busy when only a few thousand rows come through (7 k rows
        // x 256 features = 440 CTAs), but they are bound by shared-memory wavefronts (3-4 per 16 FMAs); from about 12 k rows
        // on 128 x 64 tiles (8 x 4 per thread, half the wavefronts per FMA) fill the GPU as well: measured 133 vs 85 us at
        // 7 k rows, 352 vs 436 us at 21 k rows x 1024 inputs.  Same chains, same bits.
        if (lin.out > 64 && rows_hint >= 12288) return launch<128, 64, 8, 4>(lin, x, rows, n, y, relu, s, n_dev);
        if (lin.out > 64) return launch<64, 64, 4, 4>(lin, x, rows, n, y, relu, s, n_dev);
        if (lin.out > 32) return launch<128, 64, 8, 4>(lin, x, rows, n, y, relu, s, n_dev);
        return launch<128, 32, 4, 4>(lin, x, rows, n, y, relu, s, n_dev);
    }
    RQB_CHECK(lin.set, "linear layer not loaded");
    RQB_CHECK(lin.in % 4 == 0, "in_features must be a multiple of 4 (got %d)", lin.in);
    RQB_CHECK(lin.nblk >= 1 && lin.nblk <= 8, "at most 8 K-blocks supported (got %d)", lin.nblk);
    RQB_CHECK(n <= (int64_t)2147483647 * 128, "too many rows");
    // a batch of 2..15 rows: the reference's CPU GEMM switches to its small-batch summation order (small_batch.cu)
    if (small_batch_lane16(batch_rows < 0 ? n : batch_rows, lin.in)) return linear_small(lin, x, rows, n, y, relu, s);
    // few rows (rescue tier, collision groups): 128 x 128 tiles leave most SMs idle (6.9 k rows x 256 features = 108 CTAs);
    // 64 x 64 tiles give four times the CTAs.  Same chains, same bits.
    if (lin.out > 64 && n <= 24576) return launch<64, 64, 4, 4>(lin, x, rows, n, y, relu, s);
    if (lin.out > 64) return launch<128, 128, 8, 8>(lin, x, rows, n, y, relu, s);
    if (lin.out > 32) return launch<128, 64, 8, 4>(lin, x, rows, n, y, relu, s);
    return launch<128, 32, 4, 4>(lin, x, rows, n, y, relu, s);
}

}  // namespace rqb

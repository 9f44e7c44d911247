// small_batch.cu — the reference's arithmetic on SMALL batches, and the per-group re-encode built on it.
//
// The reference resolves collisions by calling `model.get_indices(data[collision_items], use_sk=True)` once per
// collision group (reference RQ-VAE/infer.py:120-129 ≡ RQ-VAE/generate_code.py:115-124): the encoder
// (RQ-VAE/models/layers.py:42-43) and every quantizer level (RQ-VAE/models/vq.py:71-75) run again on the 2 … few dozen
// rows of the group alone.  Its CPU GEMM (MKL sgemm behind ATen addmm / matmul) uses another kernel for such batches
// than for the catalogue pass, with another summation order, so the latent of an item inside a small group differs in
// the last bit from its latent in the catalogue pass — enough to flip a Sinkhorn arg-max between near-identical items.
// Restated here (derived and verified bit for bit on the live reference stack by oracle/probe_sum_order.py):
//
//   an [M,K]·[K,N] product with 2 <= M <= 15 and 24·M <= K ("lane16"):
//       16 fp32 lanes, lane l = sequential fma chain over k ≡ l (mod 16), k ascending;
//       folded ((p0 + p1) + p2) + p3 with p_q = lanes 4q … 4q+3, then (s0 + s1) + (s2 + s3), bias added last;
//   everything else: the catalogue order of linear_exact.cu / quantize.cu.
//
// Kernels: linear_small_kernel (one Linear over a row list whose rows carry their own batch size M),
// quantize_small_kernel (levels with the arg-min rule; one warp per row), group_sizes_kernel.  SIMT fp32 by necessity
// (bit parity with an fp32 fma chain); the volume is the colliding items only.
#include "common.cuh"

namespace rqb {

namespace {

constexpr int LS_THREADS = 128;
constexpr int LS_ROWS = 16;                 // slots per CTA (rows staged in shared memory)
constexpr int LS_TN = 2;                    // output features per thread: j and j + 128

// slot -> size of the group it belongs to (items are stored group by group, offsets[g] .. offsets[g+1])
__global__ void group_sizes_kernel(const int64_t *__restrict__ offsets, int64_t n_groups, int64_t n_items,
                                   int *__restrict__ msize) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    int64_t lo = 0, hi = n_groups;              // largest g with offsets[g] <= i
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t sz = offsets[lo + 1] - offsets[lo];
    msize[i] = sz > 2147483647 ? 2147483647 : (int)sz;
}

// slots → two compact lists: members of groups with fewer than 16 rows (small-batch arithmetic) and the rest (catalogue
// arithmetic in every layer: the fast exact kernels take those).  Order inside a list is irrelevant: rows are independent.
__global__ void partition_by_size_kernel(const int64_t *__restrict__ items, const int *__restrict__ msize, int64_t n_items,
                                         int64_t *__restrict__ small_items, int *__restrict__ small_msize,
                                         int64_t *__restrict__ big_items, unsigned long long *__restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n_items;
    const bool big = live && msize[i] >= 16;
    const unsigned mb = __ballot_sync(0xffffffffu, big), ms = __ballot_sync(0xffffffffu, live && !big);
    const int lane = threadIdx.x & 31;
    unsigned long long base_b = 0, base_s = 0;
    if (lane == 0) {
        if (mb) base_b = atomicAdd(&counts[1], (unsigned long long)__popc(mb));
        if (ms) base_s = atomicAdd(&counts[0], (unsigned long long)__popc(ms));
    }
    base_b = __shfl_sync(0xffffffffu, base_b, 0);
    base_s = __shfl_sync(0xffffffffu, base_s, 0);
    const unsigned below = (1u << lane) - 1;
    if (big) big_items[base_b + __popc(mb & below)] = items[i];
    else if (live) {
        const unsigned long long p = base_s + __popc(ms & below);
        small_items[p] = items[i];
        small_msize[p] = msize[i];
    }
}


// ---- memo of the per-item part of the re-encode ------------------------------------------------------------------------
// What rqb200_reencode_groups computes for a member — its codes on the first L-1 levels and the residual entering the last
// level — depends on the item and on the SIZE of its group only through which products take the lane16 order, i.e. through
// a handful of thresholds (min(15, K/24) per product).  The sizes between two thresholds form a class; (item, class) → result
// is a pure function, so a member that was already re-encoded inside a group of the same class is answered from the memo.
struct SizeClasses {
    int nb;                 // number of thresholds; classes 0 … nb (class nb: every product in the catalogue order)
    int bound[8];           // ascending; class(M) = #{k : bound[k] < M}
};
__device__ __forceinline__ int size_class(const SizeClasses &sc, int M) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) c += (k < sc.nb && sc.bound[k] < M) ? 1 : 0;
    return c;
}

// slots → (a) members found in the memo: nothing listed (memo_restore_kernel copies them), (b) misses of groups with fewer
// than 16 rows, (c) misses of larger groups.  counts[0] = small misses, counts[1] = big misses.
__global__ void memo_classify_kernel(const int64_t *__restrict__ items, const int *__restrict__ msize, int64_t n_items,
                                     SizeClasses sc, const unsigned *__restrict__ have, int64_t *__restrict__ small_items,
                                     int *__restrict__ small_msize, int64_t *__restrict__ big_items,
                                     unsigned long long *__restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n_items;
    int M = 0;
    int64_t item = 0;
    bool miss = false;
    if (live) {
        M = msize[i];
        item = items[i];
        miss = !((have[item] >> size_class(sc, M)) & 1u);
    }
    const bool big = miss && M >= 16, small = miss && M < 16;
    const unsigned mb = __ballot_sync(0xffffffffu, big), ms = __ballot_sync(0xffffffffu, small);
    const int lane = threadIdx.x & 31;
    unsigned long long base_b = 0, base_s = 0;
    if (lane == 0) {
        if (mb) base_b = atomicAdd(&counts[1], (unsigned long long)__popc(mb));
        if (ms) base_s = atomicAdd(&counts[0], (unsigned long long)__popc(ms));
    }
    base_b = __shfl_sync(0xffffffffu, base_b, 0);
    base_s = __shfl_sync(0xffffffffu, base_s, 0);
    const unsigned below = (1u << lane) - 1;
    if (big) big_items[base_b + __popc(mb & below)] = item;
    else if (small) {
        const unsigned long long p = base_s + __popc(ms & below);
        small_items[p] = item;
        small_msize[p] = M;
    }
}

// members found in the memo: codes[item, 0..L-2] and residual[item, :] from the memo (thread = one float4 of one slot)
__global__ void memo_restore_kernel(const int64_t *__restrict__ items, const int *__restrict__ msize, int64_t n_items,
                                    SizeClasses sc, const unsigned *__restrict__ have, const float *__restrict__ memo_res,
                                    const int *__restrict__ memo_codes, int64_t n_cat, int e, int L,
                                    int64_t *__restrict__ codes, float *__restrict__ residual) {
    const int e4 = e >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slot = t / e4;
    const int q = (int)(t % e4);
    if (slot >= n_items) return;
    const int64_t item = items[slot];
    const int c = size_class(sc, msize[slot]);
    if (!((have[item] >> c) & 1u)) return;
    const size_t row = (size_t)c * n_cat + item;
    reinterpret_cast<float4 *>(residual + item * e)[q] = reinterpret_cast<const float4 *>(memo_res + row * e)[q];
    if (q == 0)
        for (int l = 0; l < L - 1; ++l) codes[item * L + l] = memo_codes[row * (L - 1) + l];
}

// freshly computed members → memo.  msize == nullptr: every listed item is of class sc.nb (groups of 16 or more rows)
__global__ void memo_store_kernel(const int64_t *__restrict__ list, const int *__restrict__ msize, int64_t n_list,
                                  SizeClasses sc, unsigned *__restrict__ have, float *__restrict__ memo_res,
                                  int *__restrict__ memo_codes, int64_t n_cat, int e, int L,
                                  const int64_t *__restrict__ codes, const float *__restrict__ residual) {
    const int e4 = e >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slot = t / e4;
    const int q = (int)(t % e4);
    if (slot >= n_list) return;
    const int64_t item = list[slot];
    const int c = msize ? size_class(sc, msize[slot]) : sc.nb;
    const size_t row = (size_t)c * n_cat + item;
    reinterpret_cast<float4 *>(memo_res + row * e)[q] = reinterpret_cast<const float4 *>(residual + item * e)[q];
    if (q == 0) {
        for (int l = 0; l < L - 1; ++l) memo_codes[row * (L - 1) + l] = (int)codes[item * L + l];
        have[item] |= 1u << c;            // an item is a member of at most one group per round: no other thread touches this word
    }
}

__device__ __forceinline__ float fold_lane16(const float (&acc)[16]) {
    float s[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        s[q] = __fadd_rn(__fadd_rn(__fadd_rn(acc[q], acc[q + 4]), acc[q + 8]), acc[q + 12]);
    return __fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3]));
}

// W[N,K] → Wt[K,N] so that the threads of a warp (consecutive output features) read consecutive addresses
__global__ void transpose_w_kernel(const float *__restrict__ W, int N, int K, float *__restrict__ Wt) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int nn = n0 + i, kk = k0 + threadIdx.x;
        tile[i][threadIdx.x] = (nn < N && kk < K) ? W[(size_t)nn * K + kk] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int kk = k0 + i, nn = n0 + threadIdx.x;
        if (kk < K && nn < N) Wt[(size_t)kk * N + nn] = tile[threadIdx.x][i];
    }
}

// Y[slot, :] = act(X[row(slot), :] · Wᵀ + b) where every slot uses the order of ITS batch size (msize[slot], or m_uniform).
// CTA = LS_ROWS consecutive slots x 256 output features (blockIdx.y selects the 256-column block), thread = features j, j + 128.
// The rows sit in shared memory LANE-MAJOR — element k of a row at [k mod 16][k div 16] — so a lane of the small-batch order
// (k ≡ l mod 16, ascending) is contiguous: one LDS.128 (broadcast) feeds 4 k-steps x LS_TN features.  The lanes are walked
// in the order of the fold, (l = q, q+4, q+8, q+12 for q = 0..3), which needs only four running values per output:
//   s_q = ((P_q + P_{q+4}) + P_{q+8}) + P_{q+12};   result = (s_0 + s_1) + (s_2 + s_3);   + bias last.
// Rows whose batch size takes the catalogue order (K-blocks of sequential chains, bias first) are computed by a second,
// scalar-k pass over the same staged rows — rare (groups of 16 or more members).
__global__ void __launch_bounds__(LS_THREADS)
linear_small_kernel(const float *__restrict__ X, const int64_t *__restrict__ rows, const int *__restrict__ msize,
                    int m_uniform, int64_t n, const float *__restrict__ Wt, const float *__restrict__ bias,
                    float *__restrict__ Y, int K, int N, int relu, int nblk, const int kb0, const int kb1, const int kb2,
                    const int kb3, const int kb4, const int kb5, const int kb6, const int kb7) {
    extern __shared__ __align__(16) float xs[];           // [LS_ROWS][16][KCP]
    __shared__ int s_kind[LS_ROWS];
    const int KC = (K + 15) >> 4;                          // elements per lane
    const int KCP = (KC + 3) & ~3;                         // padded to whole float4s (zero filled: fma(0, w, acc) == acc)
    const int64_t slot0 = (int64_t)blockIdx.x * LS_ROWS;
    const int nr = (int)((n - slot0) < LS_ROWS ? (n - slot0) : LS_ROWS);
    for (int i = threadIdx.x; i < LS_ROWS * 16 * KCP; i += LS_THREADS) xs[i] = 0.0f;
    __syncthreads();
    for (int i = threadIdx.x; i < LS_ROWS * (K / 4); i += LS_THREADS) {
        const int r = i / (K / 4), c4 = i % (K / 4);
        if (r < nr) {
            const int64_t src = rows ? rows[slot0 + r] : slot0 + r;
            const float4 v = *reinterpret_cast<const float4 *>(X + src * K + 4 * c4);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int k = 4 * c4 + t;
                xs[((size_t)r * 16 + (k & 15)) * KCP + (k >> 4)] = e[t];
            }
        }
    }
    if (threadIdx.x < LS_ROWS) {
        const int r = threadIdx.x;
        const int M = r < nr ? (msize ? msize[slot0 + r] : m_uniform) : 0;
        s_kind[r] = r < nr ? (small_batch_lane16(M, K) ? 1 : 0) : -1;
    }
    __syncthreads();
    bool any_lane = false, any_blk = false;
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r) { any_lane |= s_kind[r] == 1; any_blk |= s_kind[r] == 0; }
    const int kb[8] = {kb0, kb1, kb2, kb3, kb4, kb5, kb6, kb7};
    const int j0 = blockIdx.y * (LS_THREADS * LS_TN) + threadIdx.x;
    bool live[LS_TN];
    float bj[LS_TN];
#pragma unroll
    for (int t = 0; t < LS_TN; ++t) {
        live[t] = j0 + t * LS_THREADS < N;
        bj[t] = (live[t] && bias) ? bias[j0 + t * LS_THREADS] : 0.0f;
    }
    const float *w0 = Wt + (live[0] ? j0 : 0);
    const float *w1 = Wt + (live[1] ? j0 + LS_THREADS : 0);

    float out[LS_ROWS][LS_TN];
    if (any_lane) {
        float tt[LS_ROWS][LS_TN], uu[LS_ROWS][LS_TN];
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float ss[LS_ROWS][LS_TN];
#pragma unroll 1
            for (int p = 0; p < 4; ++p) {
                const int l = q + 4 * p;
                float acc[LS_ROWS][LS_TN];
#pragma unroll
                for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
                    for (int t = 0; t < LS_TN; ++t) acc[r][t] = 0.0f;
                for (int c = 0; c < KC; c += 4) {
                    float wa[4], wb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = 16 * (c + i) + l;
                        const bool ok = k < K;
                        wa[i] = ok ? w0[(size_t)k * N] : 0.0f;
                        wb[i] = ok ? w1[(size_t)k * N] : 0.0f;
                    }
#pragma unroll
                    for (int r = 0; r < LS_ROWS; ++r) {
                        const float4 xv = *reinterpret_cast<const float4 *>(xs + ((size_t)r * 16 + l) * KCP + c);
                        const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (c + i < KC) {          // ascending k inside the lane; the padded tail adds nothing
                                acc[r][0] = __fmaf_rn(xe[i], wa[i], acc[r][0]);
                                acc[r][1] = __fmaf_rn(xe[i], wb[i], acc[r][1]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
                    for (int t = 0; t < LS_TN; ++t) ss[r][t] = p == 0 ? acc[r][t] : __fadd_rn(ss[r][t], acc[r][t]);
            }
#pragma unroll
            for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
                for (int t = 0; t < LS_TN; ++t) {
                    if (q == 0) tt[r][t] = ss[r][t];
                    else if (q == 1) tt[r][t] = __fadd_rn(tt[r][t], ss[r][t]);
                    else if (q == 2) uu[r][t] = ss[r][t];
                    else tt[r][t] = __fadd_rn(tt[r][t], __fadd_rn(uu[r][t], ss[r][t]));
                }
        }
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
            for (int t = 0; t < LS_TN; ++t)
                if (s_kind[r] == 1) out[r][t] = __fadd_rn(tt[r][t], bj[t]);
    }
    if (any_blk) {
        float o[LS_ROWS][LS_TN];
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
            for (int t = 0; t < LS_TN; ++t) o[r][t] = bj[t];
        int k0 = 0;
        for (int blk = 0; blk < nblk; ++blk) {
            const int k1 = k0 + kb[blk];
            float acc[LS_ROWS][LS_TN];
#pragma unroll
            for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
                for (int t = 0; t < LS_TN; ++t) acc[r][t] = 0.0f;
            for (int k = k0; k < k1; ++k) {
                const float wa = w0[(size_t)k * N], wb = w1[(size_t)k * N];
                const float *xk = xs + (size_t)(k & 15) * KCP + (k >> 4);
#pragma unroll
                for (int r = 0; r < LS_ROWS; ++r) {
                    const float xv = xk[(size_t)r * 16 * KCP];
                    acc[r][0] = __fmaf_rn(xv, wa, acc[r][0]);
                    acc[r][1] = __fmaf_rn(xv, wb, acc[r][1]);
                }
            }
#pragma unroll
            for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
                for (int t = 0; t < LS_TN; ++t) o[r][t] = __fadd_rn(o[r][t], acc[r][t]);
            k0 = k1;
        }
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
            for (int t = 0; t < LS_TN; ++t)
                if (s_kind[r] == 0) out[r][t] = o[r][t];
    }
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
        for (int t = 0; t < LS_TN; ++t)
            if (r < nr && live[t]) {
                float v = out[r][t];
                if (relu) v = (v != v) ? v : (v > 0.0f ? v : 0.0f);
                Y[(slot0 + r) * N + j0 + t * LS_THREADS] = v;
            }
}

// torch.sum(v*v) for a runtime length (ATen order; twin of sinkhorn.cu's helper, kept local to this unit)
__device__ float sumsq_aten_small(const float *v, int e) {
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
                const float x = v[i * 32 + k * 8 + l];
                part[k][l] = __fadd_rn(part[k][l], __fmul_rn(x, x));
            }
    for (int i = size_ilp * 4; i < vec; ++i)
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float x = v[i * 8 + l];
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

constexpr int QS_WARPS = 4;

struct SmallQuantArgs {
    const float *cb[RQB200_MAX_LEVELS];
    const float *cc[RQB200_MAX_LEVELS];
    int K[RQB200_MAX_LEVELS];
    int L;
};

// ResidualVectorQuantizer.forward with the arg-min rule on levels [0, levels_run) for rows that carry their own batch
// size (rq.py:39-56, vq.py:63-99).  One warp per slot; the row's residual lives in shared memory.
//   codes[item(slot), l] for l < levels_run; residual_out[item(slot), :] = residual entering level `levels_run`
//   (or, with levels_run == L, after the last level); xq_out[slot, :] (optional, levels_run == L);
//   sumsq_out[l] += Σ (q - r)² (optional, fp64 atomics).
__global__ void __launch_bounds__(QS_WARPS * 32)
quantize_small_kernel(const float *__restrict__ z, const int64_t *__restrict__ items, const int *__restrict__ msize,
                      int m_uniform, int64_t n, int e, SmallQuantArgs qa, int levels_run, int64_t *__restrict__ codes,
                      float *__restrict__ residual_out, float *__restrict__ xq_out, double *__restrict__ sumsq_out,
                      float *__restrict__ dist_out, int dist_level) {
    extern __shared__ __align__(16) float sm[];      // [QS_WARPS][2][e]: residual, x_q
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * QS_WARPS + wid;
    if (slot >= n) return;
    float *r = sm + (size_t)wid * 2 * e, *xq = r + e;
    for (int k = lane; k < e; k += 32) { r[k] = z[slot * e + k]; xq[k] = 0.0f; }
    __syncwarp();
    const int64_t item = items ? items[slot] : slot;
    const int M = msize ? msize[slot] : m_uniform;
    const bool lane16 = small_batch_lane16(M, e);
    for (int l0 = 0; l0 < levels_run; ++l0) {
        const int l = dist_out ? dist_level : l0;      // dist_out: distances of ONE level for the rows as given (vq.py:71-73)
        const int K = qa.K[l];
        const float *cb = qa.cb[l], *cc = qa.cc[l];
        const float xx = sumsq_aten_small(r, e);
        int best = -1;
        float bestd = 0.0f;
        for (int j = lane; j < K; j += 32) {
            const float *c = cb + (size_t)j * e;
            float dot;
            if (lane16) {
                float acc[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] = 0.0f;
                int k = 0;
                for (; k + 16 <= e; k += 16) {
                    float cv[16], rv[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {           // e is a multiple of 4 and the rows are 16-byte aligned: 128-bit loads
                        const float4 a = *reinterpret_cast<const float4 *>(c + k + 4 * q), b = *reinterpret_cast<const float4 *>(r + k + 4 * q);
                        cv[4 * q] = a.x; cv[4 * q + 1] = a.y; cv[4 * q + 2] = a.z; cv[4 * q + 3] = a.w;
                        rv[4 * q] = b.x; rv[4 * q + 1] = b.y; rv[4 * q + 2] = b.z; rv[4 * q + 3] = b.w;
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc[q] = __fmaf_rn(rv[q], cv[q], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (k + q < e) acc[q] = __fmaf_rn(r[k + q], c[k + q], acc[q]);
                dot = fold_lane16(acc);
            } else {
                // one sequential chain over k; the operands of the next 8 steps are fetched ahead as two 128-bit loads each
                dot = 0.0f;
                int k = 0;
                for (; k + 8 <= e; k += 8) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(c + k), a1 = *reinterpret_cast<const float4 *>(c + k + 4);
                    const float4 b0 = *reinterpret_cast<const float4 *>(r + k), b1 = *reinterpret_cast<const float4 *>(r + k + 4);
                    dot = __fmaf_rn(b0.x, a0.x, dot); dot = __fmaf_rn(b0.y, a0.y, dot); dot = __fmaf_rn(b0.z, a0.z, dot); dot = __fmaf_rn(b0.w, a0.w, dot);
                    dot = __fmaf_rn(b1.x, a1.x, dot); dot = __fmaf_rn(b1.y, a1.y, dot); dot = __fmaf_rn(b1.z, a1.z, dot); dot = __fmaf_rn(b1.w, a1.w, dot);
                }
                for (; k < e; ++k) dot = __fmaf_rn(r[k], c[k], dot);
            }
            const float d = __fsub_rn(__fadd_rn(xx, cc[j]), __fmul_rn(2.0f, dot));
            if (dist_out) dist_out[slot * K + j] = d;
            // torch.argmin: NaN is the minimum, the first index wins
            if (best < 0 || (!(bestd != bestd) && ((d != d) || d < bestd))) { best = j; bestd = d; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bestd, o);
            const int oj = __shfl_xor_sync(0xffffffffu, best, o);
            if (oj >= 0) {
                const bool onan = od != od, bnan = bestd != bestd;
                const bool take = best < 0 || (onan && (!bnan || oj < best)) ||
                                  (!onan && !bnan && (od < bestd || (od == bestd && oj < best)));
                if (take) { bestd = od; best = oj; }
            }
        }
        if (dist_out) return;
        if (lane == 0) codes[item * qa.L + l] = best;
        const float *q = cb + (size_t)best * e;
        double loss = 0.0;
        for (int k = lane; k < e; k += 32) {
            const float diff = __fsub_rn(q[k], r[k]);
            loss += (double)diff * (double)diff;
            const float xres = __fadd_rn(r[k], diff);          // x + (x_q - x)   (vq.py:95)
            r[k] = __fsub_rn(r[k], xres);                       // rq.py:47
            xq[k] = (l == 0) ? __fadd_rn(0.0f, xres) : __fadd_rn(xq[k], xres);   // rq.py:48
        }
        if (sumsq_out) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
            if (lane == 0) atomicAdd(&sumsq_out[l], loss);
        }
        __syncwarp();
    }
    if (residual_out)
        for (int k = lane; k < e; k += 32) residual_out[item * e + k] = r[k];
    if (xq_out)
        for (int k = lane; k < e; k += 32) xq_out[slot * e + k] = xq[k];
}

int launch_linear_small(const Linear &lin, const float *x, const int64_t *rows, const int *msize, int m_uniform,
                        int64_t n, float *y, bool relu, cudaStream_t s) {
    RQB_CHECK(lin.set, "linear layer not loaded");
    RQB_CHECK(lin.in % 4 == 0, "in_features must be a multiple of 4 (got %d)", lin.in);
    RQB_CHECK(lin.nblk >= 1 && lin.nblk <= 8, "at most 8 K-blocks supported (got %d)", lin.nblk);
    const int KC = (lin.in + 15) / 16, KCP = (KC + 3) & ~3;
    const size_t smem = sizeof(float) * LS_ROWS * 16 * (size_t)KCP;
    RQB_CHECK(smem <= 200 * 1024, "in_features %d too large for the small-batch kernel", lin.in);
    static rqb::DeviceOnce attr_once;
    if (attr_once.first())
        RQB_CUDA(cudaFuncSetAttribute(linear_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (!lin.Wt) {
        float *p = nullptr;
        RQB_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)lin.in * lin.out));
        rqb::count_launch();
        transpose_w_kernel<<<dim3((lin.in + 31) / 32, (lin.out + 31) / 32), dim3(32, 8), 0, s>>>(lin.W, lin.out, lin.in, p);
        RQB_LAUNCH_CHECK();
        lin.Wt = p;
    }
    int kb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < lin.nblk; ++i) kb[i] = lin.kblocks[i];
    rqb::count_launch();
    const dim3 grid((unsigned)((n + LS_ROWS - 1) / LS_ROWS), (unsigned)((lin.out + LS_THREADS * LS_TN - 1) / (LS_THREADS * LS_TN)));
    linear_small_kernel<<<grid, LS_THREADS, smem, s>>>(x, rows, msize, m_uniform, n, lin.Wt, lin.b, y, lin.in, lin.out,
                                                       relu ? 1 : 0, lin.nblk, kb[0], kb[1], kb[2], kb[3], kb[4], kb[5], kb[6],
                                                       kb[7]);
    RQB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// one uniform small batch (module-level get_indices / forward with 2..15 rows)
int linear_small(const Linear &lin, const float *x, const int64_t *rows, int64_t n, float *y, bool relu, cudaStream_t s) {
    return launch_linear_small(lin, x, rows, nullptr, (int)n, n, y, relu, s);
}

int quantize_small(const rqb200_model *m, const float *z, const int64_t *items, const int *msize, int m_uniform, int64_t n,
                   int levels_run, int64_t *codes, float *residual_out, float *xq_out, double *sumsq_out, float *dist_out,
                   int dist_level, cudaStream_t s) {
    if (n == 0) return 0;
    SmallQuantArgs qa;
    qa.L = m->L;
    for (int l = 0; l < m->L; ++l) { qa.cb[l] = m->cb[l]; qa.cc[l] = m->cc[l]; qa.K[l] = m->K[l]; }
    const size_t smem = sizeof(float) * QS_WARPS * 2 * (size_t)m->e;
    RQB_CHECK(smem <= 48 * 1024, "e_dim %d too large for the small-batch quantizer", m->e);
    RQB_CHECK(m->e % 4 == 0, "e_dim must be a multiple of 4 (got %d)", m->e);
    rqb::count_launch();
    quantize_small_kernel<<<(unsigned)((n + QS_WARPS - 1) / QS_WARPS), QS_WARPS * 32, smem, s>>>(
        z, items, msize, m_uniform, n, m->e, qa, dist_out ? 1 : levels_run, codes, residual_out, xq_out, sumsq_out, dist_out,
        dist_level);
    RQB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rqb

using namespace rqb;

// shared body: every listed row is re-encoded up to (not including) the last level in the order of ITS batch size
static int reencode_rows_impl(rqb200_model *m, const float *x_dev, int x_is_gathered, const int64_t *items_dev, const int *msize,
                              int64_t n_items, float *buf0, float *buf1, int64_t *codes_dev, float *residual_dev, cudaStream_t s) {
    float *buf[2] = {buf0, buf1};
    const float *cur = x_dev;
    for (int i = 0; i < m->n_layers; ++i) {
        const bool last = i == m->n_layers - 1;
        float *dst = buf[i & 1];
        RQB_TRY(launch_linear_small(m->enc[i], cur, (i == 0 && !x_is_gathered) ? items_dev : nullptr, msize, 0, n_items, dst,
                                    !last, s));
        cur = dst;
    }
    return quantize_small(m, cur, items_dev, msize, 0, n_items, m->L - 1, codes_dev, residual_dev, nullptr, nullptr, nullptr, 0, s);
}

static int reencode_check(rqb200_model *m) {
    RQB_CHECK(m != nullptr, "model is NULL");
    for (int i = 0; i < m->n_layers; ++i) RQB_CHECK(m->enc[i].set, "encoder layer %d not loaded", i);
    for (int l = 0; l < m->L; ++l) RQB_CHECK(m->cb_set[l], "codebook %d not loaded", l);
    return 0;
}

// batch-size classes of this model (see SizeClasses): one threshold min(15, K/24) >= 2 per product on the way to the last level
static SizeClasses size_classes_of(const rqb200_model *m) {
    int t[RQB200_MAX_LAYERS + 1], nt = 0;
    for (int i = 0; i < m->n_layers; ++i) t[nt++] = m->enc[i].in / 24;
    t[nt++] = m->e / 24;
    SizeClasses sc;
    sc.nb = 0;
    for (int k = 0; k < 8; ++k) sc.bound[k] = 0;
    for (int v = 2; v <= 15; ++v) {                 // ascending, distinct
        bool present = false;
        for (int i = 0; i < nt; ++i) present |= (t[i] > 15 ? 15 : t[i]) == v;
        if (present && sc.nb < 8) sc.bound[sc.nb++] = v;
    }
    return sc;
}

struct ReencodeMemo {
    unsigned *have = nullptr;     // [n_cat] bit c: (item, class c) is in the memo
    float *res = nullptr;         // [n_classes][n_cat][e]
    int *codes = nullptr;         // [n_classes][n_cat][L-1]
    int64_t n_cat = 0;
};

// infer.py:120-122 for ALL collision groups of one round, up to (not including) the last level:
//   for every group g and every member i:  z = encoder(x_i) in the order of a batch of |g| rows,
//   codes[i, 0..L-2] = arg-min codes of the first L-1 levels (distance product in the order of that batch size),
//   residual[i, :]   = residual entering the last level.
// The last level (Sinkhorn over the group's distance matrix) is rqb200_sinkhorn_regroup on `residual`.
static int reencode_groups_impl(rqb200_model *m, const float *x_dev, int x_is_gathered, const int64_t *items_dev,
                                const int64_t *offsets_dev, int64_t n_groups, int64_t n_items,
                                int64_t *codes_dev, float *residual_dev, const ReencodeMemo *memo, cudaStream_t s) {
    RQB_TRY(reencode_check(m));
    if (n_groups == 0 || n_items == 0) return 0;
    RQB_CHECK(x_dev && items_dev && offsets_dev && codes_dev && residual_dev, "NULL buffer");
    RQB_CHECK(!memo || !x_is_gathered, "the memo needs the catalogue on the device (x_is_gathered = 0)");
    RQB_CUDA(cudaSetDevice(m->device));
    int maxdim = m->e;
    for (int i = 0; i < m->n_layers; ++i) maxdim = m->enc[i].out > maxdim ? m->enc[i].out : maxdim;
    const size_t isz = ((sizeof(int64_t) * (size_t)n_items + 255) / 256) * 256;
    const size_t msz = ((sizeof(int) * (size_t)n_items + 255) / 256) * 256;
    // msize | small msize | small items | big items | counts | two activation buffers (sized once the split is known)
    const size_t head = 2 * msz + 2 * isz + 256;
    RQB_TRY(ws_reserve(m->groupws, head));
    char *base = (char *)m->groupws.ptr;
    int *msize = (int *)base;
    int *small_msize = (int *)(base + msz);
    int64_t *small_items = (int64_t *)(base + 2 * msz);
    int64_t *big_items = (int64_t *)(base + 2 * msz + isz);
    unsigned long long *counts = (unsigned long long *)(base + 2 * msz + 2 * isz);
    const SizeClasses sc = size_classes_of(m);
    const int e4 = m->e / 4;
    ProfScope ps(PROF_REENCODE, s);
    rqb::count_launch();
    group_sizes_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, s>>>(offsets_dev, n_groups, n_items, msize);
    unsigned long long h[2] = {(unsigned long long)n_items, 0};
    if (!x_is_gathered) {
        // groups of 16 or more rows use the catalogue arithmetic in every layer: hand those rows to the fast exact kernels
        RQB_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), s));
        rqb::count_launch();
        if (memo) {
            memo_classify_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, s>>>(items_dev, msize, n_items, sc, memo->have, small_items,
                                                                                 small_msize, big_items, counts);
            rqb::count_launch();
            memo_restore_kernel<<<(unsigned)((n_items * e4 + 255) / 256), 256, 0, s>>>(items_dev, msize, n_items, sc, memo->have, memo->res,
                                                                                      memo->codes, memo->n_cat, m->e, m->L, codes_dev,
                                                                                      residual_dev);
        } else {
            partition_by_size_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, s>>>(items_dev, msize, n_items, small_items, small_msize,
                                                                                     big_items, counts);
        }
        RQB_LAUNCH_CHECK();
        RQB_CUDA(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, s));
        RQB_CUDA(cudaStreamSynchronize(s));
    }
    const int64_t n_small = (int64_t)h[0], n_big = (int64_t)h[1];
    if (n_small > 0) {
        const size_t act = sizeof(float) * (size_t)n_small * maxdim;
        RQB_TRY(ws_reserve(m->rescue_act[0], act));
        RQB_TRY(ws_reserve(m->rescue_act[1], act));
        RQB_TRY(reencode_rows_impl(m, x_dev, x_is_gathered, x_is_gathered ? items_dev : small_items, x_is_gathered ? msize : small_msize,
                                   n_small, (float *)m->rescue_act[0].ptr, (float *)m->rescue_act[1].ptr, codes_dev, residual_dev, s));
        if (memo) {
            rqb::count_launch();
            memo_store_kernel<<<(unsigned)((n_small * e4 + 255) / 256), 256, 0, s>>>(small_items, small_msize, n_small, sc, memo->have,
                                                                                    memo->res, memo->codes, memo->n_cat, m->e, m->L,
                                                                                    codes_dev, residual_dev);
            RQB_LAUNCH_CHECK();
        }
    }
    if (n_big > 0) {
        const int64_t batch = (int64_t)1 << 30;               // "a large batch": catalogue order in every layer
        RQB_TRY(ws_reserve(m->rescue_act[0], sizeof(float) * (size_t)n_big * maxdim));
        RQB_TRY(ws_reserve(m->rescue_act[1], sizeof(float) * (size_t)n_big * maxdim));
        RQB_TRY(ws_reserve(m->rescue, sizeof(float) * (size_t)n_big * m->e));
        const float *cur = x_dev;
        for (int i = 0; i < m->n_layers; ++i) {
            float *dst = (float *)m->rescue_act[i & 1].ptr;
            RQB_TRY(linear_exact(m->enc[i], cur, i == 0 ? big_items : nullptr, n_big, dst, i != m->n_layers - 1, s, batch));
            cur = dst;
        }
        float *res_compact = (float *)m->rescue.ptr;
        // all L arg-min levels are written; the Sinkhorn pass then overwrites the last one
        RQB_TRY(quantize_exact(m, cur, n_big, codes_dev, big_items, nullptr, nullptr, res_compact, nullptr, s, batch));
        RQB_TRY(scatter_rows(res_compact, big_items, n_big, m->e, residual_dev, s));
        if (memo) {
            rqb::count_launch();
            memo_store_kernel<<<(unsigned)((n_big * e4 + 255) / 256), 256, 0, s>>>(big_items, nullptr, n_big, sc, memo->have, memo->res,
                                                                                  memo->codes, memo->n_cat, m->e, m->L, codes_dev,
                                                                                  residual_dev);
            RQB_LAUNCH_CHECK();
        }
    }
    return 0;
}

extern "C" int rqb200_reencode_groups(rqb200_model *m, const float *x_dev, int x_is_gathered, const int64_t *items_dev,
                                      const int64_t *offsets_dev, int64_t n_groups, int64_t n_items,
                                      int64_t *codes_dev, float *residual_dev, void *stream) {
    return reencode_groups_impl(m, x_dev, x_is_gathered, items_dev, offsets_dev, n_groups, n_items, codes_dev, residual_dev, nullptr,
                                (cudaStream_t)stream);
}

extern "C" int rqb200_reencode_classes(const rqb200_model *m) {
    if (!m) return 0;
    return size_classes_of(m).nb + 1;
}

extern "C" int rqb200_reencode_groups_memo(rqb200_model *m, const float *x_dev, const int64_t *items_dev,
                                           const int64_t *offsets_dev, int64_t n_groups, int64_t n_items, int64_t *codes_dev,
                                           float *residual_dev, int64_t n_catalogue, unsigned *memo_have_dev, float *memo_residual_dev,
                                           int *memo_codes_dev, void *stream) {
    RQB_CHECK(m != nullptr, "model is NULL");
    RQB_CHECK(memo_have_dev && memo_residual_dev && (memo_codes_dev || m->L == 1), "NULL memo buffer");
    RQB_CHECK(n_catalogue > 0, "n_catalogue must be positive");
    ReencodeMemo memo;
    memo.have = memo_have_dev;
    memo.res = memo_residual_dev;
    memo.codes = memo_codes_dev;
    memo.n_cat = n_catalogue;
    return reencode_groups_impl(m, x_dev, 0, items_dev, offsets_dev, n_groups, n_items, codes_dev, residual_dev, &memo,
                                (cudaStream_t)stream);
}

// The same for rows that are members of groups whose other members live elsewhere (sharded catalogue): the caller
// supplies the size of every row's group (msize_dev, int32 [n_rows]) instead of the group lists.
extern "C" int rqb200_reencode_rows(rqb200_model *m, const float *x_dev, int x_is_gathered, const int64_t *rows_dev,
                                    const int *msize_dev, int64_t n_rows, int64_t *codes_dev, float *residual_dev, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RQB_TRY(reencode_check(m));
    if (n_rows == 0) return 0;
    RQB_CHECK(x_dev && rows_dev && msize_dev && codes_dev && residual_dev, "NULL buffer");
    RQB_CUDA(cudaSetDevice(m->device));
    int maxdim = m->e;
    for (int i = 0; i < m->n_layers; ++i) maxdim = m->enc[i].out > maxdim ? m->enc[i].out : maxdim;
    const size_t act = sizeof(float) * (size_t)n_rows * maxdim;
    RQB_TRY(ws_reserve(m->groupws, 2 * act));
    ProfScope ps(PROF_REENCODE, s);
    return reencode_rows_impl(m, x_dev, x_is_gathered, rows_dev, msize_dev, n_rows, (float *)m->groupws.ptr,
                              (float *)((char *)m->groupws.ptr + act), codes_dev, residual_dev, s);
}

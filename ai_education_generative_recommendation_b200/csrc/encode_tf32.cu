// encode_tf32.cu — the wide first encoder layer (768 / 1024 → 256, reference RQ-VAE/models/layers.py:23) as ONE
// TF32 tensor pass fed by TMA: the screening tier of the fast route.
//
// Why: the three-pass split-fp16 kernel (encode_tc2.cu) is bound by its operand path — 16 converter warps pull fp32 X
// through registers (LDG → cvt → STS) in lock-step with a 3-stage ring and a cross-CTA relay per K slab (ablation:
// profiles/r2_ablation_linear_tc.txt: the empty barrier skeleton alone costs 0.34-0.41 ms of the 1.09 ms).  kind::tf32
// consumes the fp32 bits as they lie in HBM (the tensor core reads the upper 19 bits), so the X tile needs no
// conversion at all: cp.async.bulk.tensor.2d drops 128 x 32 fp32 boxes straight into SWIZZLE_128B shared memory and
// the only resident threads are one TMA lane, one MMA lane and the epilogue warps.
//
//   CTA pair (cta_group::2), M = 256 rows per pair, N = 256, accumulator fp32 in TMEM (2 x 256 columns, double buffered)
//   stage = A box 128 rows x 32 k (16 KB) + this CTA's half of W, 128 features x 32 k (16 KB); TF_STAGES stages
//   both CTAs' TMA loads complete on the LEADER's full barrier (expect_tx = 64 KB), tcgen05.commit multicasts "stage
//   free" / "accumulator full" to both CTAs; 4 MMAs (K = 8 each) per stage
//   epilogue: tcgen05.ld → per-warp transpose patch → bias + ReLU → split-fp16 UMMA tiles for mlp23_tc_kernel (or fp32 rows)
//
// Precision: operands carry 10 explicit mantissa bits (X truncated by the hardware, W rounded to nearest when packed), so
// this pass only SCREENS: every row it keeps is certified by the quantizer's margin gate at the screening bound gamma1,
// the rest is re-run by the three-pass kernels and, if still uncertified, by the exact SIMT kernels.  Algorithmic
// traffic: 4·K bytes per row read once (HBM-bound: 2·K·256 flop per row at the TF32 rate take about as long).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace rqb {

namespace {

constexpr int TF_TM = 128;                  // rows per CTA (256 per pair)
constexpr int TF_BK = 32;                   // fp32 per K slab = 128 bytes = one SWIZZLE_128B row
constexpr int TF_N = 256;
constexpr int TF_NH = TF_N / 2;
#ifndef TF_STAGES_N
#define TF_STAGES_N 6
#endif
#ifndef TF_EPI_WARPS_N
#define TF_EPI_WARPS_N 4
#endif
constexpr int TF_STAGES = TF_STAGES_N;
constexpr int TF_EPI = TF_EPI_WARPS_N;      // 4 (one warp per TMEM lane quarter) or 8 (two per quarter, 128 columns each)
constexpr int TF_A_TILE = TF_TM * TF_BK * 4;    // 16 KB
constexpr int TF_W_TILE = TF_NH * TF_BK * 4;    // 16 KB
constexpr int TF_STAGE = TF_A_TILE + TF_W_TILE;
constexpr int TF_THREADS = (TF_EPI + 2) * 32;
constexpr int TF_TMA_WARP = TF_EPI, TF_MMA_WARP = TF_EPI + 1;
constexpr int TF_SMEM = TF_STAGES * TF_STAGE + 1024 + 256 + TF_EPI * 32 * EPI_LD * 4;
constexpr int TF_TMEM_COLS = 512;
static_assert(TF_SMEM <= 227 * 1024, "shared memory budget");

// instruction descriptor: D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// address of `bar` in the LEADER CTA (rank 0) of the pair, as a shared::cluster address
__device__ __forceinline__ uint32_t leader_addr(const void *bar) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(0u));
    return ra;
}

// 2-D tiled TMA load into THIS CTA's shared memory; completion bytes are counted on a barrier that may live in the
// leader CTA (cta_group::2 form)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster_addr,
                                            uint64_t cache_hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(cache_hint)
        : "memory");
}

constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;    // X: streamed once
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;     // W: re-read by every tile

// split-fp16 tile epilogue for the column range [c_lo, c_hi) (see epilogue_rows_split in tc_common.cuh)
template <int N>
__device__ __forceinline__ void tf_epilogue_split(uint32_t taddr, float *patch, int lane, const float *__restrict__ bias, int relu,
                                                  unsigned char *__restrict__ tiled, int64_t tile, int row_in_tile0, int c_lo,
                                                  int c_hi) {
    const int rsub = lane >> 2, ch = lane & 3;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4 *>(patch + lane * EPI_LD + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const float4 ba = __ldg(reinterpret_cast<const float4 *>(bias + c + 8 * ch));
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + c + 8 * ch + 4));
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        unsigned char *hi_t = tiled + ((size_t)tile * (N / 64) + (c >> 6)) * 32768;
        const int chunk = ((c & 63) >> 3) + ch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + rsub;
            const float4 a0 = *reinterpret_cast<const float4 *>(patch + rr * EPI_LD + 8 * ch);
            const float4 a1 = *reinterpret_cast<const float4 *>(patch + rr * EPI_LD + 8 * ch + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float h[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float y = av[t] + bv[t];
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
        __syncwarp();
    }
}

// fp32 row epilogue for the column range [c_lo, c_hi)
template <int N>
__device__ __forceinline__ void tf_epilogue_rows(uint32_t taddr, float *patch, int lane, const float *__restrict__ bias, int relu,
                                                 float *__restrict__ Y, int64_t row0, int64_t n, int c_lo, int c_hi) {
    const int sub = lane >> 3, q = lane & 7;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4 *>(patch + lane * EPI_LD + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + c) + q);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + sub;
            const float4 a = *reinterpret_cast<const float4 *>(patch + r * EPI_LD + 4 * q);
            float4 o = make_float4(a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w);
            if (relu) {
                o.x = (o.x != o.x) ? o.x : fmaxf(o.x, 0.0f); o.y = (o.y != o.y) ? o.y : fmaxf(o.y, 0.0f);
                o.z = (o.z != o.z) ? o.z : fmaxf(o.z, 0.0f); o.w = (o.w != o.w) ? o.w : fmaxf(o.w, 0.0f);
            }
            if (row0 + r < n) *reinterpret_cast<float4 *>(Y + (row0 + r) * (int64_t)N + c + 4 * q) = o;
        }
        __syncwarp();
    }
}

// dbg: ablation switches (tools/ablate_tf32.py; 0 in production) — bit0 no epilogue stores (drain only), bit1 no MMA,
// bit3 no W loads, bit4 no X loads
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF_THREADS, 1)
linear_tf32_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, int64_t n, int K,
                   const float *__restrict__ bias, int relu, float *__restrict__ Y, int tiled_out, int dbg) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TF_STAGES * TF_STAGE);
    uint64_t *full = bars;                               // [STAGES] (used in the leader: both CTAs' TMA bytes land here)
    uint64_t *empty = bars + TF_STAGES;                  // [STAGES] (both CTAs: multicast commit)
    uint64_t *tmem_full = bars + 2 * TF_STAGES;          // [2]      (both CTAs: multicast commit)
    uint64_t *tmem_empty = tmem_full + 2;                // [2]      (leader: one arrival per epilogue warp of the pair)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int KS = (K + TF_BK - 1) / TF_BK;
    const int64_t npt = (n + 2 * TF_TM - 1) / (2 * TF_TM);
    const int64_t pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TF_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 2 * TF_EPI);
        }
        fence_barrier_init();
    }
    if (warp == TF_TMA_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == TF_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TF_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TF_EPI) {
        // ===================== epilogue: each CTA drains its own 128 rows =====================
        float *patch = reinterpret_cast<float *>(smem + TF_STAGES * TF_STAGE + 256) + warp * (32 * EPI_LD);
        const int quarter = warp & 3;                              // TMEM lanes 32*quarter .. +31
        const int c_lo = (TF_EPI == 8) ? (warp >> 2) * (TF_N / 2) : 0;
        const int c_hi = (TF_EPI == 8) ? c_lo + TF_N / 2 : TF_N;
        int64_t it = 0;
        for (int64_t pt = pair0; pt < npt; pt += npairs, ++it) {
            const int buf = (int)(it & 1);
            mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TF_N);
            if (!(dbg & 1)) {
                if (tiled_out)
                    tf_epilogue_split<TF_N>(taddr, patch, lane, bias, relu, reinterpret_cast<unsigned char *>(Y), pt * 2 + rank,
                                            quarter * 32, c_lo, c_hi);
                else
                    tf_epilogue_rows<TF_N>(taddr, patch, lane, bias, relu, Y, pt * (2 * TF_TM) + rank * TF_TM + quarter * 32, n,
                                           c_lo, c_hi);
            } else {
                for (int c = c_lo; c < c_hi; c += 32) { uint32_t v[32]; tmem_ld32(taddr + (uint32_t)c, v); }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tmem_empty[buf]);
                else mbar_arrive_remote(&tmem_empty[buf], 0);
            }
        }
    } else if (warp == TF_TMA_WARP) {
        // ===================== TMA producer (one lane): this CTA's 128 rows of X and its half of W =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t pt = pair0; pt < npt; pt += npairs) {
                const int row0 = (int)(pt * (2 * TF_TM) + rank * TF_TM);
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *a_dst = smem + stage * TF_STAGE;
                    const uint32_t per_cta = ((dbg & 16) ? 0 : TF_A_TILE) + ((dbg & 8) ? 0 : TF_W_TILE);
                    if (rank == 0) {                                                        // both CTAs' bytes
                        if (per_cta) mbar_arrive_expect_tx(&full[stage], 2 * per_cta);
                        else mbar_arrive(&full[stage]);
                    }
                    const uint32_t bar = leader_addr(&full[stage]);
                    if (!(dbg & 16)) tma_load_2d(a_dst, &map_x, slab * TF_BK, row0, bar, HINT_EVICT_FIRST);
                    if (!(dbg & 8)) tma_load_2d(a_dst + TF_A_TILE, &map_w, slab * TF_BK, (int)rank * TF_NH, bar, HINT_EVICT_LAST);
                    if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== MMA issuer (leader CTA, one lane) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = umma_idesc_tf32(2 * TF_TM, TF_N);
            int stage = 0;
            uint32_t phase = 0;
            int64_t it = 0;
            for (int64_t pt = pair0; pt < npt; pt += npairs, ++it) {
                const int buf = (int)(it & 1);
                mbar_wait_cluster(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * TF_N);
                for (int slab = 0; slab < KS; ++slab) {
                    mbar_wait_cluster(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a = smem_u32(smem + stage * TF_STAGE);
                    const uint32_t w = a + TF_A_TILE;
                    if (!(dbg & 2)) {
#pragma unroll
                        for (int kk = 0; kk < TF_BK / 8; ++kk)
                            umma_tf32_2cta(d_tmem, umma_desc(a + kk * 32), umma_desc(w + kk * 32), idesc, (slab | kk) != 0);
                    }
                    umma_commit_2cta(&empty[stage]);
                    if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(&tmem_full[buf]);
            }
        }
    }
    // ---- teardown: neither CTA may leave while the pair can still touch its smem / TMEM
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == TF_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TF_TMEM_COLS));
    }
}

// W[256,K] fp32 → W·comp with every element rounded to nearest-even at the TF32 mantissa (10 bits).
// comp > 1 compensates the hardware's TRUNCATION of the X operand to 10 mantissa bits: trunc(x) = x·(1 - u), u in
// [0, 2^-10) with mean ≈ 0.72·2^-11 for log-uniform mantissas — a systematic shrink of every product that would otherwise
// dominate the error of the pass (measured: mean relative error of the latent 2^-11.1 without, see tools/calibrate_gate.py).
__global__ void round_w_tf32_kernel(const float *__restrict__ W, int64_t count, float comp, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t u = __float_as_uint(W[i] * comp);
        u += 0x0FFFu + ((u >> 13) & 1u);
        out[i] = __uint_as_float(u & 0xFFFFE000u);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [rows, K] fp32 row-major → boxes of 128 rows x 32 columns, SWIZZLE_128B, out-of-range rows / columns read as zero
int make_map(CUtensorMap *map, const float *base, int64_t rows, int K) {
    EncodeTiledFn fn = encode_tiled_fn();
    RQB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TF_BK, (cuuint32_t)TF_TM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RQB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a [%lld, %d] fp32 matrix at %p", (int)r, (long long)rows, K,
              (const void *)base);
    return 0;
}

}  // namespace

int tc_debug_flags();     // encode_tc.cu

bool linear_tf32_supported(const Linear &l) { return l.out == TF_N && l.in % 4 == 0 && l.in >= TF_BK; }

// y = act(x·Wᵀ + b) with TF32 operands (screening precision).  x: [n, in] fp32 row-major, 16-byte aligned.
// tiled_out: y receives split-fp16 UMMA tiles for mlp23_tc_kernel (whole 256-row pair tiles), else fp32 rows.
int linear_tf32(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, bool tiled_out) {
    if (n == 0) return 0;
    RQB_CHECK(linear_tf32_supported(l), "TF32 screening kernel needs out_features 256 and in_features %% 4 == 0");
    RQB_CHECK(((uintptr_t)x & 15) == 0, "x must be 16-byte aligned for TMA");
    RQB_CHECK(n < ((int64_t)1 << 31), "too many rows for one launch");
    if (!l.W_tf32) {
        void *p = nullptr;
        RQB_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)l.out * l.in));
        count_launch();
        float comp = 1.0f + 0.72f * 4.8828125e-4f;                 // 1 + 0.72·2^-11
        if (const char *ev = getenv("RQB200_TF32_COMP")) comp = 1.0f + (float)atof(ev) * 4.8828125e-4f;   // diagnostics: in units of 2^-11
        round_w_tf32_kernel<<<kNumSMs, 256, 0, s>>>(l.W, (int64_t)l.out * l.in, comp, (float *)p);
        RQB_LAUNCH_CHECK();
        l.W_tf32 = (float *)p;
    }
    CUtensorMap map_x, map_w;
    RQB_TRY(make_map(&map_x, x, n, l.in));
    RQB_TRY(make_map(&map_w, l.W_tf32, l.out, l.in));
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM));
    }
    const int64_t npt = (n + 2 * TF_TM - 1) / (2 * TF_TM);
    const int64_t pairs = npt < kNumSMs / 2 ? npt : kNumSMs / 2;
    count_launch();
    linear_tf32_kernel<<<(unsigned)(pairs * 2), TF_THREADS, TF_SMEM, s>>>(map_x, map_w, n, l.in, l.b, relu ? 1 : 0, y,
                                                                         tiled_out ? 1 : 0, tc_debug_flags());
    RQB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rqb

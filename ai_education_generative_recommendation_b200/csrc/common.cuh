// common.cuh — shared declarations of the rqvae_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rqvae_b200.h"

namespace rqb {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char *fmt, ...);
void count_launch();          // bumps the process-wide kernel-launch counter (rqb200_launch_count)

// per-kernel device timing (cudaEvent pairs on the launching stream), enabled by rqb200_profile_enable
enum ProfSlot { PROF_LINEAR0 = 0, PROF_LINEAR_REST = 1, PROF_QUANTIZE = 2, PROF_DEDUP = 3, PROF_TC_ENCODER = 4,
                PROF_SINKHORN = 5, PROF_TC_REST = 6, PROF_TIER2 = 7, PROF_RESCUE = 8, PROF_REENCODE = 9, PROF_NSLOTS = 12 };
void prof_begin(int slot, cudaStream_t s);
void prof_end(int slot, cudaStream_t s);
struct ProfScope {
    int slot; cudaStream_t s;
    ProfScope(int slot_, cudaStream_t s_) : slot(slot_), s(s_) { prof_begin(slot, s); }
    ~ProfScope() { prof_end(slot, s); }
};

#define RQB_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            rqb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return RQB200_ECUDA;                                                            \
        }                                                                                   \
    } while (0)

#define RQB_CHECK(cond, ...)                                                                \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            rqb::set_error(__VA_ARGS__);                                                    \
            return RQB200_EINVAL;                                                           \
        }                                                                                   \
    } while (0)

#define RQB_TRY(expr)                                                                       \
    do {                                                                                    \
        int _rc = (expr);                                                                   \
        if (_rc != 0) return _rc;                                                           \
    } while (0)

#define RQB_LAUNCH_CHECK() RQB_CUDA(cudaGetLastError())

// Which [M,K]·[K,N] products the reference's CPU GEMM computes in its small-batch ("lane16") order — see small_batch.cu.
__host__ __device__ __forceinline__ bool small_batch_lane16(long long M, int K) {
    return M >= 2 && M <= 15 && 24 * M <= (long long)K;
}

struct Linear {
    int in = 0, out = 0;
    float *W = nullptr;      // [out, in] row-major (device)
    float *b = nullptr;      // [out]
    int kblocks[16];
    int nblk = 0;
    bool set = false;
    // tensor-core path: split-fp16 (hi | lo) images of W, packed per K-slab in UMMA SW128 layout
    void *W_tc = nullptr;
    size_t W_tc_bytes = 0;
    float *absW_rowmax = nullptr;
    int tc_scale_exp = 0;    // W_tc holds W * 2^tc_scale_exp
    void *W_tc2 = nullptr;   // image for the 2-CTA kernel (each CTA of a pair stages half of the output features)
    float *W_tf32 = nullptr; // W rounded to the TF32 mantissa (screening kernel, encode_tf32.cu)
    mutable float *Wt = nullptr;   // [in, out] transposed copy (small-batch kernel, small_batch.cu), built on first use
};

// "Do this once per device": cudaFuncSetAttribute is a per-device setting, and one process may drive several devices
// through the C ABI (rqb200_model_create takes a device index).  first() is true the first time it is called while
// a given device is current.
struct DeviceOnce {
    unsigned long long mask = 0;
    bool first() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long bit = 1ull << (d & 63);
        const unsigned long long old = __atomic_fetch_or(&mask, bit, __ATOMIC_RELAXED);
        return !(old & bit);
    }
};

struct Workspace {
    void *ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace rqb

struct rqb200_model {
    int device = 0;
    int n_layers = 0;
    int dims[RQB200_MAX_LAYERS + 1];
    rqb::Linear enc[RQB200_MAX_LAYERS];
    rqb::Linear dec[RQB200_MAX_LAYERS];
    int L = 0;
    int e = 0;
    int K[RQB200_MAX_LEVELS];
    float *cb[RQB200_MAX_LEVELS];   // [K, e]
    float *cc[RQB200_MAX_LEVELS];   // [K]  sum of squares in the reference's order
    bool cb_set[RQB200_MAX_LEVELS];
    void *cb_tc[RQB200_MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // packed fp16 hi/lo chunks
    int cb_tc_scale_exp[RQB200_MAX_LEVELS] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cb_tc_aug_exp[RQB200_MAX_LEVELS] = {0, 0, 0, 0, 0, 0, 0, 0};
    float *ccs_tc[RQB200_MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // e == 64: scaled norms
    rqb::Workspace act[2];          // ping-pong activations for the MLP
    rqb::Workspace sortws;          // radix sort scratch
    rqb::Workspace misc;            // rescue lists, counters
    rqb::Workspace hostpipe[2];     // device chunks of the host-buffer pipeline
    rqb::Workspace rescue;          // exact latent of gated rows
    rqb::Workspace groupws;         // per-group re-encode: group sizes + activations of the colliding items
    rqb::Workspace skws;            // Sinkhorn re-encode: per-size-class group lists + tickets
    rqb::Workspace rescue_act[2];
    float gate_gamma = 3.0517578125e-05f;   // 2^-15: bound on |z~ - z| / |z| of the tensor-core encoder
    int use_2cta = -1;                      // -1: decide from RQB200_TC2 env (default on), 0/1: forced
    bool force_simt_quantizer = false;      // diagnostics: keep the SIMT quantizer behind the tensor-core encoder
    float gate_floor = 1.0e-3f;             // absolute floor added to |z| in that bound
    bool screen_enabled = false;            // tier 1: one reduced-precision pass over every row, only gated rows get the 3-pass run
    bool screen_auto = true;                // until rqb200_model_set_screen is called: TF32 screening where it pays (see get_indices_fast)
    int screen_kind = 1;                    // 1: TF32 first layer fed by TMA + three-pass tail (encode_tf32.cu); 0: one fp16 pass through all layers
    float screen_gamma = 4.8828125e-04f;    // 2^-11: calibrated gate of the screening tier (tools/calibrate_gate.py)
    int64_t last_tier_rows[2] = {0, 0};     // rows re-run by tier 2 / tier 3 in the last fast get_indices
    unsigned long long *tier_counts_dev = nullptr;   // the same two counts on the device (the fast route does not wait for them)
    cudaStream_t tier_counts_stream = nullptr;
    bool tier_counts_pending = false;       // last_tier_rows is stale until rqb200_model_last_tier_rows reads tier_counts_dev
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace rqb {

int ws_reserve(Workspace &w, size_t bytes);

// linear_exact.cu
// batch_rows: how many rows the reference would have in the batch these n rows belong to (decides the summation order,
// small_batch.cu); -1 = n.  Internal callers that recompute a SUBSET of a large batch (rescue tier) pass the batch size.
// n_dev (may be NULL): the row count lives on the device (n = capacity; persistent grid) — the exact tier of the fast route;
// rows_hint: about how many rows that count will be (picks the tile shape only).
int linear_exact(const Linear &lin, const float *x, const int64_t *rows, int64_t n, float *y,
                 bool relu, cudaStream_t s, int64_t batch_rows = -1, const unsigned long long *n_dev = nullptr,
                 int64_t rows_hint = -1);
// small_batch.cu: the reference's order for batches of 2..15 rows (and per-row batch sizes for the group re-encode)
int linear_small(const Linear &lin, const float *x, const int64_t *rows, int64_t n, float *y, bool relu, cudaStream_t s);
int quantize_small(const rqb200_model *m, const float *z, const int64_t *items, const int *msize, int m_uniform, int64_t n,
                   int levels_run, int64_t *codes, float *residual_out, float *xq_out, double *sumsq_out, float *dist_out,
                   int dist_level, cudaStream_t s);
// encode_tc2.cu
bool linear_tc2_supported(const Linear &l);
int linear_tc2(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, int passes = 3,
               const int64_t *rows = nullptr, const unsigned long long *n_dev = nullptr, bool tiled_out = false);
// encode_tf32.cu: TMA-fed one-pass TF32 first layer (screening tier)
bool linear_tf32_supported(const Linear &l);
int linear_tf32(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, bool tiled_out);
// encode_tc.cu
int linear_tc(Linear &l, const float *x, int64_t n, float *y, bool relu, cudaStream_t s, int passes = 3,
              const int64_t *rows = nullptr, const unsigned long long *n_dev = nullptr, bool tiled_out = false);
int mlp_tc(rqb200_model *m, int which, const float *x, int64_t n, float *y, cudaStream_t s, int passes = 3,
           const int64_t *rows = nullptr, const unsigned long long *n_dev = nullptr, bool profile = true);
int scatter_rows(const float *src, const int64_t *rows, int64_t nr, int e, float *dst, cudaStream_t s);   // encode_tc.cu: dst[rows[i]] = src[i]
// quantize_tc.cu
bool quantize_tc_supported(const rqb200_model *m);
// rows (may be NULL): row i of z is item rows[i] (codes / list entries use the item index); n_dev: device-resident
// row count (n = upper bound); gamma: relative bound on |z~ - z| used by the margin gate of this tier.
int quantize_tc(rqb200_model *m, const float *z, int64_t n, int64_t *codes, int64_t *list, unsigned long long *count,
                cudaStream_t s, float gamma, const int64_t *rows = nullptr, const unsigned long long *n_dev = nullptr);
// quantize.cu
int codebook_norms(const float *cb, int K, int e, float *cc, cudaStream_t s);
int quantize_exact(const rqb200_model *m, const float *z, int64_t n, int64_t *codes,
                   const int64_t *rows_out, float *xq, double *sumsq, float *last_residual,
                   float *margin_out, cudaStream_t s, int64_t batch_rows = -1, const unsigned long long *n_dev = nullptr,
                   int64_t n_hint = 0);
int distances_exact(const rqb200_model *m, int level, const float *r, int64_t n, float *d,
                    cudaStream_t s);
int recon_error(const float *out, const float *x, int64_t count, double *recon_sum, cudaStream_t s);

}  // namespace rqb

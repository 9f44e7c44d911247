/*
 * rqvae_oracle.c — CPU restatement of the reference RQ-VAE semantic-ID encode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product path (the CUDA C-ABI
 * library) never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the
 * reference's own PyTorch CPU modules (imported from /root/reference/RQ-VAE in the
 * build container by oracle/make_golden.py) and against the golden vectors that
 * script committed under tests/golden/.
 *
 * The reference is pure PyTorch; its fp32 results are whatever ATen/MKL compute on
 * the CPU.  The summation orders restated here were derived from the reference's
 * call sites and verified against its outputs:
 *
 *   - nn.Linear (reference RQ-VAE/models/layers.py:23, forward at layers.py:42-43)
 *     → addmm → MKL sgemm with beta=1 on a bias-filled C:
 *       y_j = ((b_j + chain(blk0)) + chain(blk1)) + ...,
 *       chain(blk) = acc=0; for k ascending in blk: acc = fma(x_k, W_jk, acc).
 *   - torch.sum(v**2, dim=1) (reference RQ-VAE/models/vq.py:71-72) → ATen
 *     vectorized_inner_sum / row_sum / multi_row_sum with 8 fp32 lanes and
 *     4-way ILP, cascade levels included.
 *   - torch.matmul(latent, E.t()) (vq.py:73) → one FMA chain per (row, code).
 *   - d = (xx + cc_j) - 2*dot_j, argmin first-index, NaN wins (vq.py:71-75).
 *   - x_res = r + (q - r); r = r - x_res; x_q = x_q + x_res
 *     (vq.py:95, reference RQ-VAE/models/rq.py:47-48).
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * fma() calls are explicit; -ffp-contract=off keeps every other a*b+c unfused.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_THREADS 64

/* ------------------------------------------------------------------ threading */

typedef void (*range_fn)(void *ctx, int64_t lo, int64_t hi);
typedef struct { range_fn fn; void *ctx; int64_t lo, hi; } job_t;

static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->ctx, j->lo, j->hi); return NULL; }

static void parallel_rows(range_fn fn, void *ctx, int64_t n, int threads) {
    if (threads < 1) threads = 1;
    if (threads > ORACLE_MAX_THREADS) threads = ORACLE_MAX_THREADS;
    if (n < 64 || threads == 1) { fn(ctx, 0, n); return; }
    pthread_t tid[ORACLE_MAX_THREADS];
    job_t jobs[ORACLE_MAX_THREADS];
    int64_t per = (n + threads - 1) / threads;
    int started = 0;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = t * per, hi = lo + per;
        if (lo >= n) break;
        if (hi > n) hi = n;
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].lo = lo; jobs[t].hi = hi;
        pthread_create(&tid[t], NULL, job_main, &jobs[t]);
        ++started;
    }
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
}

/* ------------------------------------------------------------------ Linear */

typedef struct {
    const float *x; const float *wt; const float *b; float *y;
    int in_dim, out_dim, relu; const int *kblocks; int nblk;
} linear_ctx;

#define ROWS_PER_STEP 4
#define COLS_PER_STEP 64

/* Each (row, out) accumulator is its own sequential fma chain over k; blocking over
 * rows and columns only changes which chains run side by side, never their order. */
static void linear_rows(void *p, int64_t lo, int64_t hi) {
    linear_ctx *c = (linear_ctx *)p;
    const int K = c->in_dim, N = c->out_dim;
    for (int64_t r0 = lo; r0 < hi; r0 += ROWS_PER_STEP) {
        int nr = (int)((hi - r0) < ROWS_PER_STEP ? (hi - r0) : ROWS_PER_STEP);
        for (int j0 = 0; j0 < N; j0 += COLS_PER_STEP) {
            int nc = (N - j0) < COLS_PER_STEP ? (N - j0) : COLS_PER_STEP;
            float out[ROWS_PER_STEP][COLS_PER_STEP];
            for (int r = 0; r < nr; ++r)
                for (int j = 0; j < nc; ++j) out[r][j] = c->b ? c->b[j0 + j] : 0.0f;
            int k0 = 0;
            for (int blk = 0; blk < c->nblk; ++blk) {
                int k1 = k0 + c->kblocks[blk];
                float acc[ROWS_PER_STEP][COLS_PER_STEP];
                memset(acc, 0, sizeof(acc));
                if (nr == ROWS_PER_STEP && nc == COLS_PER_STEP) {
                    for (int k = k0; k < k1; ++k) {
                        const float *w = c->wt + (int64_t)k * N + j0;
                        for (int r = 0; r < ROWS_PER_STEP; ++r) {
                            float xv = c->x[(r0 + r) * K + k];
                            for (int j = 0; j < COLS_PER_STEP; ++j)
                                acc[r][j] = __builtin_fmaf(xv, w[j], acc[r][j]);
                        }
                    }
                } else {
                    for (int k = k0; k < k1; ++k) {
                        const float *w = c->wt + (int64_t)k * N + j0;
                        for (int r = 0; r < nr; ++r) {
                            float xv = c->x[(r0 + r) * K + k];
                            for (int j = 0; j < nc; ++j)
                                acc[r][j] = __builtin_fmaf(xv, w[j], acc[r][j]);
                        }
                    }
                }
                for (int r = 0; r < nr; ++r)
                    for (int j = 0; j < nc; ++j) out[r][j] = out[r][j] + acc[r][j];
                k0 = k1;
            }
            for (int r = 0; r < nr; ++r)
                for (int j = 0; j < nc; ++j) {
                    float v = out[r][j];
                    /* torch relu: NaN propagates, -0 stays -0 only if v > 0 fails → 0 */
                    if (c->relu) v = (v != v) ? v : (v > 0.0f ? v : 0.0f);
                    c->y[(r0 + r) * N + j0 + j] = v;
                }
        }
    }
}

/* y[n,out] = act(x[n,in] · W[out,in]^T + b).  kblocks sums to in_dim. */
int rq_oracle_linear(const float *x, int64_t n, int in_dim, const float *W, const float *b,
                     int out_dim, int relu, const int *kblocks, int nblk, float *y, int threads) {
    int tot = 0;
    for (int i = 0; i < nblk; ++i) tot += kblocks[i];
    if (tot != in_dim) return -1;
    float *wt = (float *)malloc(sizeof(float) * (size_t)in_dim * out_dim);
    if (!wt) return -2;
    for (int j = 0; j < out_dim; ++j)
        for (int k = 0; k < in_dim; ++k) wt[(int64_t)k * out_dim + j] = W[(int64_t)j * in_dim + k];
    linear_ctx c = { x, wt, b, y, in_dim, out_dim, relu, kblocks, nblk };
    parallel_rows(linear_rows, &c, n, threads);
    free(wt);
    return 0;
}

/* ------------------------------------------------------------------ Linear, small batches
 *
 * The reference re-runs its whole model on every collision group (reference RQ-VAE/infer.py:120-122 →
 * rqvae.py:67-71 → layers.py:42-43), i.e. on batches of 2 … a few dozen rows, and its CPU GEMM (MKL sgemm
 * behind ATen addmm / matmul) picks another kernel for such small M: a dot-product kernel that keeps 16 fp32
 * lanes (one AVX-512 register), lane l accumulating k ≡ l (mod 16) as a sequential fma chain, then folds the
 * register 512 → 128 bits as ((p0 + p1) + p2) + p3 (p_q = lanes 4q … 4q+3), the four survivors as
 * (s0 + s1) + (s2 + s3), and adds the bias (beta = 1) last.  Derived with oracle/probe_sum_order.py (mask
 * probe of the summation tree on the live reference stack) and verified bit for bit on random data there.
 * Which M use it depends on the layer shape; the table lives in oracle.py (small_batch_plan). */
static float dot_lane16(const float *a, const float *b, int K) {
    float acc[16];
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int k = 0;
    for (; k + 16 <= K; k += 16)
        for (int l = 0; l < 16; ++l) acc[l] = __builtin_fmaf(a[k + l], b[k + l], acc[l]);
    for (int l = 0; k + l < K; ++l) acc[l] = __builtin_fmaf(a[k + l], b[k + l], acc[l]);   /* masked tail (unprobed) */
    float s[4];
    for (int m = 0; m < 4; ++m) s[m] = ((acc[m] + acc[m + 4]) + acc[m + 8]) + acc[m + 12];
    return (s[0] + s[1]) + (s[2] + s[3]);
}

/* kind 1: lane16 (above);  kind 2: four K-blocks folded pairwise, ((c0 + b) + c1) + (c2 + c3) — what the
 * reference stack does for 1024 → 256 with 16 ≤ M < 176 on the 8-thread build container. */
int rq_oracle_linear_small(const float *x, int64_t n, int in_dim, const float *W, const float *b,
                           int out_dim, int relu, int kind, float *y) {
    if (kind != 1 && kind != 2) return -1;
    if (kind == 2 && (in_dim % 4)) return -1;
    for (int64_t r = 0; r < n; ++r)
        for (int j = 0; j < out_dim; ++j) {
            const float *xr = x + r * in_dim, *wj = W + (int64_t)j * in_dim;
            float v;
            if (kind == 1) {
                v = dot_lane16(xr, wj, in_dim) + (b ? b[j] : 0.0f);
            } else {
                int q = in_dim / 4; float c[4];
                for (int blk = 0; blk < 4; ++blk) {
                    float acc = 0.0f;
                    for (int k = blk * q; k < (blk + 1) * q; ++k) acc = __builtin_fmaf(xr[k], wj[k], acc);
                    c[blk] = acc;
                }
                v = ((c[0] + (b ? b[j] : 0.0f)) + c[1]) + (c[2] + c[3]);
            }
            if (relu) v = (v != v) ? v : (v > 0.0f ? v : 0.0f);
            y[r * out_dim + j] = v;
        }
    return 0;
}

/* ------------------------------------------------------------------ sum of squares */

#define LANES 8
#define ILP 4

static int ceil_log2_i64(int64_t x) {
    if (x <= 2) return 1;
    int l = 0; int64_t v = x - 1;
    while (v > 0) { v >>= 1; ++l; }
    return l;
}

/* ATen multi_row_sum over `size` steps of ILP vectors (each LANES wide), base = sq,
 * step stride ILP*LANES floats, row k at offset k*LANES. */
static void multi_row_sum_vec(const float *sq, int64_t size, float acc0[ILP][LANES]) {
    enum { NUM_LEVELS = 4 };
    int level_power = ceil_log2_i64(size) / NUM_LEVELS;
    if (level_power < 4) level_power = 4;
    const int64_t level_step = (int64_t)1 << level_power;
    const int64_t level_mask = level_step - 1;
    float acc[NUM_LEVELS][ILP][LANES];
    memset(acc, 0, sizeof(acc));
    int64_t i = 0;
    for (; i + level_step <= size;) {
        for (int64_t j = 0; j < level_step; ++j, ++i) {
            const float *base = sq + i * ILP * LANES;
            for (int k = 0; k < ILP; ++k)
                for (int l = 0; l < LANES; ++l) acc[0][k][l] = acc[0][k][l] + base[k * LANES + l];
        }
        for (int j = 1; j < NUM_LEVELS; ++j) {
            for (int k = 0; k < ILP; ++k)
                for (int l = 0; l < LANES; ++l) {
                    acc[j][k][l] = acc[j][k][l] + acc[j - 1][k][l];
                    acc[j - 1][k][l] = 0.0f;
                }
            const int64_t mask = level_mask << (j * level_power);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size; ++i) {
        const float *base = sq + i * ILP * LANES;
        for (int k = 0; k < ILP; ++k)
            for (int l = 0; l < LANES; ++l) acc[0][k][l] = acc[0][k][l] + base[k * LANES + l];
    }
    for (int j = 1; j < NUM_LEVELS; ++j)
        for (int k = 0; k < ILP; ++k)
            for (int l = 0; l < LANES; ++l) acc[0][k][l] = acc[0][k][l] + acc[j][k][l];
    memcpy(acc0, acc[0], sizeof(float) * ILP * LANES);
}

/* torch.sum(v**2, dim=1) for one contiguous row of length e (fp32). */
float rq_oracle_sumsq_row(const float *v, int e) {
    float sqbuf[4096];
    float *sq = sqbuf;
    if (e > 4096) sq = (float *)malloc(sizeof(float) * (size_t)e);
    for (int i = 0; i < e; ++i) sq[i] = v[i] * v[i];
    const int64_t vec_size = e / LANES;          /* number of LANES-wide vectors */
    const int64_t size_ilp = vec_size / ILP;
    float part[ILP][LANES];
    multi_row_sum_vec(sq, size_ilp, part);
    for (int64_t i = size_ilp * ILP; i < vec_size; ++i)
        for (int l = 0; l < LANES; ++l) part[0][l] = part[0][l] + sq[i * LANES + l];
    for (int k = 1; k < ILP; ++k)
        for (int l = 0; l < LANES; ++l) part[0][l] = part[0][l] + part[k][l];
    float fin = 0.0f;
    for (int64_t k = vec_size * LANES; k < e; ++k) fin = fin + sq[k];
    for (int l = 0; l < LANES; ++l) fin = fin + part[0][l];
    if (sq != sqbuf) free(sq);
    return fin;
}

void rq_oracle_sumsq(const float *v, int64_t n, int e, float *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = rq_oracle_sumsq_row(v + i * e, e);
}

/* ------------------------------------------------------------------ residual quantizer */

typedef struct {
    const float *z; int e; int L; const int *K; const float *const *cb; const float *const *cc;
    int64_t *idx; float *xq; double *loss_sq; /* per level, per thread slot */ float *dist0; int dist_level;
    double *loss_slots; int nslots; int dot_kind;
} quant_ctx;

static float dot_chain(const float *a, const float *b, int e) {
    float acc = 0.0f;
    for (int k = 0; k < e; ++k) acc = __builtin_fmaf(a[k], b[k], acc);
    return acc;
}

static void quant_rows_impl(quant_ctx *c, int64_t lo, int64_t hi, double *loss /*[L]*/) {
    const int e = c->e;
    float r[1024], xq[1024];
    for (int64_t i = lo; i < hi; ++i) {
        memcpy(r, c->z + i * e, sizeof(float) * e);
        for (int l = 0; l < c->L; ++l) {
            const float *cb = c->cb[l];
            const float xx = rq_oracle_sumsq_row(r, e);
            int best = 0; float bestd = 0.0f; int have = 0;
            for (int j = 0; j < c->K[l]; ++j) {
                float dot = c->dot_kind == 1 ? dot_lane16(r, cb + (int64_t)j * e, e) : dot_chain(r, cb + (int64_t)j * e, e);
                float d = (xx + c->cc[l][j]) - (2.0f * dot);
                if (c->dist0 && l == c->dist_level) c->dist0[i * c->K[l] + j] = d;
                if (!have) { best = j; bestd = d; have = 1; }
                else if (!(bestd != bestd) && ((d != d) || d < bestd)) { best = j; bestd = d; }
            }
            c->idx[i * c->L + l] = best;
            const float *q = cb + (int64_t)best * e;
            for (int k = 0; k < e; ++k) {
                float diff = q[k] - r[k];
                loss[l] += (double)diff * (double)diff;
                float xres = r[k] + diff;         /* x + (x_q - x)  (vq.py:95) */
                r[k] = r[k] - xres;               /* rq.py:47 */
                xq[k] = (l == 0) ? (0.0f + xres) : (xq[k] + xres);   /* rq.py:48 */
            }
        }
        if (c->xq) memcpy(c->xq + i * e, xq, sizeof(float) * e);
    }
}

typedef struct { quant_ctx *c; } quant_job;

static void quant_rows(void *p, int64_t lo, int64_t hi) {
    quant_ctx *c = (quant_ctx *)p;
    double loss[16] = {0};
    quant_rows_impl(c, lo, hi, loss);
    /* slot chosen by lo so that concurrent ranges never share one */
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_lock(&mu);
    for (int l = 0; l < c->L; ++l) c->loss_sq[l] += loss[l];
    pthread_mutex_unlock(&mu);
}

/* z[n,e] → idx[n,L] (int64), xq[n,e] (may be NULL), loss_sq[L] = Σ (q-r)^2 per level in fp64
 * (may be NULL), dist_out[n,K[dist_level]] optional full distance matrix of one level.
 * codebooks: concatenated [ΣK, e]; K[l] rows per level.  e ≤ 1024, L ≤ 16. */
int rq_oracle_quantize_ex(const float *z, int64_t n, int e, const float *codebooks, const int *K, int L,
                          int64_t *idx, float *xq, double *loss_sq, float *dist_out, int dist_level,
                          int threads, int dot_kind) {
    if (e > 1024 || L > 16) return -1;
    const float *cb[16]; float *cc[16];
    int64_t off = 0;
    for (int l = 0; l < L; ++l) {
        cb[l] = codebooks + off * e;
        cc[l] = (float *)malloc(sizeof(float) * (size_t)K[l]);
        rq_oracle_sumsq(cb[l], K[l], e, cc[l]);
        off += K[l];
    }
    double lsum[16] = {0};
    quant_ctx c;
    memset(&c, 0, sizeof(c));
    c.z = z; c.e = e; c.L = L; c.K = K; c.cb = cb; c.cc = (const float *const *)cc;
    c.idx = idx; c.xq = xq; c.loss_sq = lsum; c.dist0 = dist_out; c.dist_level = dist_level; c.dot_kind = dot_kind;
    parallel_rows(quant_rows, &c, n, dist_out ? 1 : threads);
    if (loss_sq) for (int l = 0; l < L; ++l) loss_sq[l] = lsum[l];
    for (int l = 0; l < L; ++l) free(cc[l]);
    return 0;
}

int rq_oracle_quantize(const float *z, int64_t n, int e, const float *codebooks, const int *K, int L,
                       int64_t *idx, float *xq, double *loss_sq, float *dist_out, int dist_level,
                       int threads) {
    return rq_oracle_quantize_ex(z, n, e, codebooks, K, L, idx, xq, loss_sq, dist_out, dist_level, threads, 0);
}

/* ------------------------------------------------------------------ suffix dedup */

/* suffix[i] = #{ j < i : codes[j,:] == codes[i,:] }   (reference RQ-VAE/infer.py:152-163).
 * Restated with an open-addressing hash table instead of the reference's O(N·G) scan;
 * the result is defined purely by the formula above. */
static uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
    return x;
}

int rq_oracle_suffix(const int64_t *codes, int64_t n, int L, int64_t *out /*[n, L+1]*/) {
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    int64_t *slot_first = (int64_t *)malloc(sizeof(int64_t) * cap);   /* first item index with this code, -1 empty */
    int64_t *slot_count = (int64_t *)calloc(cap, sizeof(int64_t));
    if (!slot_first || !slot_count) { free(slot_first); free(slot_count); return -2; }
    for (uint64_t i = 0; i < cap; ++i) slot_first[i] = -1;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t h = 0x9e3779b97f4a7c15ULL;
        for (int l = 0; l < L; ++l) h = mix64(h ^ (uint64_t)codes[i * L + l]);
        uint64_t s = h & (cap - 1);
        for (;;) {
            if (slot_first[s] < 0) { slot_first[s] = i; break; }
            if (memcmp(codes + slot_first[s] * L, codes + i * L, sizeof(int64_t) * L) == 0) break;
            s = (s + 1) & (cap - 1);
        }
        for (int l = 0; l < L; ++l) out[i * (L + 1) + l] = codes[i * L + l];
        out[i * (L + 1) + L] = slot_count[s]++;
    }
    free(slot_first); free(slot_count);
    return 0;
}

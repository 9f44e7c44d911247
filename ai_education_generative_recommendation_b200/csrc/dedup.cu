// dedup.cu — collision grouping and the suffix column, as integer CUDA kernels.
//
// Replaces (bit-exactly) the tail of the reference's encode driver:
//   * get_collision_item / check_collision (reference RQ-VAE/infer.py:18-42): groups of items that
//     share a full code;
//   * the suffix-code dedup (infer.py:152-163): out[i, L] = #{ j < i : codes[j] == codes[i] }.
// The reference does this with Python dicts of strings and an O(N * #groups) numpy scan.  Here:
// codes are bit-packed into one u64 key per item, (key, item) pairs go through a stable LSD radix
// sort (8-bit digits, only as many passes as the key has bits), and because the sort is stable the
// position of an item inside its run of equal keys IS the number of equal codes with a smaller
// item index.  Everything is HBM-bound integer work: coalesced loads, shared-memory ranking, grids
// sized from the SM count.
#include <string.h>

#include <new>

#include "common.cuh"

namespace rqb {

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;                                // keys per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;         // 2048 keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int WARP_CHUNK = 32 * SORT_ITEMS;                  // contiguous keys owned by one warp
constexpr int RADIX = 256;

// ---------------------------------------------------------------- key packing

struct PackArgs {
    int shift[RQB200_MAX_LEVELS];
    int bits[RQB200_MAX_LEVELS];     // width of the field of level l (the key is unpacked again by seg_rank_kernel)
    int L;
};

__global__ void pack_keys_kernel(const int64_t *__restrict__ codes, int64_t n, PackArgs pa,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0;
    for (int l = 0; l < pa.L; ++l) k |= (uint64_t)codes[i * pa.L + l] << pa.shift[l];
    keys[i] = k;
    if (vals) vals[i] = (uint32_t)i;
}

__global__ void column_minmax_kernel(const int64_t *__restrict__ codes, int64_t n, int L,
                                     long long *__restrict__ mn, long long *__restrict__ mx) {
    // one thread per row; per-column warp reduction then atomics
    long long lo[RQB200_MAX_LEVELS], hi[RQB200_MAX_LEVELS];
    for (int l = 0; l < L; ++l) { lo[l] = 0x7fffffffffffffffLL; hi[l] = -0x7fffffffffffffffLL - 1; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        for (int l = 0; l < L; ++l) {
            long long v = codes[i * L + l];
            lo[l] = v < lo[l] ? v : lo[l];
            hi[l] = v > hi[l] ? v : hi[l];
        }
    for (int l = 0; l < L; ++l) {
        for (int o = 16; o > 0; o >>= 1) {
            long long a = __shfl_xor_sync(0xffffffffu, lo[l], o);
            long long b = __shfl_xor_sync(0xffffffffu, hi[l], o);
            lo[l] = a < lo[l] ? a : lo[l];
            hi[l] = b > hi[l] ? b : hi[l];
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&mn[l], lo[l]);
            atomicMax(&mx[l], hi[l]);
        }
    }
}

// ---------------------------------------------------------------- radix sort (stable, LSD)

__device__ __forceinline__ uint32_t digit_of(uint64_t k, int shift) { return (uint32_t)(k >> shift) & (RADIX - 1); }

// One read of the input for ALL digit passes: pack the codes (PACK) or take the keys as they are, and add every
// 8-bit digit of every key to the global histograms ghist[pass][digit] (shared-memory partials first).  The
// histogram of a digit does not depend on the order of the keys, so the passes below never count again.
constexpr int MAX_PASSES = 8;

template <bool PACK>
__global__ void __launch_bounds__(256)
pack_hist_kernel(const int64_t *__restrict__ codes, int64_t n, PackArgs pa, uint64_t *__restrict__ keys, int npasses,
                 uint32_t *__restrict__ ghist, const unsigned long long *__restrict__ n_dev) {
    __shared__ uint32_t s_h[MAX_PASSES][RADIX];
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    __shared__ int64_t s_codes[PACK ? 256 * RQB200_MAX_LEVELS : 1];      // 256 rows, read with coalesced loads
    for (int p = 0; p < npasses; ++p) s_h[p][threadIdx.x] = 0;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
        const int cnt = n - base < 256 ? (int)(n - base) : 256;
        uint64_t k = 0;
        if (PACK) {
            __syncthreads();
            for (int e = threadIdx.x; e < cnt * pa.L; e += 256) s_codes[e] = codes[base * pa.L + e];
            __syncthreads();
            if ((int)threadIdx.x < cnt) {
#pragma unroll
                for (int l = 0; l < RQB200_MAX_LEVELS; ++l)
                    if (l < pa.L) k |= (uint64_t)s_codes[threadIdx.x * pa.L + l] << pa.shift[l];
                keys[base + threadIdx.x] = k;
            }
        } else {
            if (base == (int64_t)blockIdx.x * 256) __syncthreads();     // histograms cleared
            if ((int)threadIdx.x < cnt) k = keys[base + threadIdx.x];
        }
        if ((int)threadIdx.x < cnt)
            for (int p = 0; p < npasses; ++p) atomicAdd(&s_h[p][(uint32_t)(k >> (8 * p)) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    for (int p = 0; p < npasses; ++p) {
        const uint32_t c = s_h[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * RADIX + threadIdx.x], c);
    }
}

// per-digit exclusive scan over blocks (one CTA per digit) + digit totals; then a 256-wide scan of the totals
__global__ void __launch_bounds__(256) radix_scan_rows_kernel(uint32_t *__restrict__ hist, int nblocks,
                                                              uint32_t *__restrict__ digit_total) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_carry;
    uint32_t *row = hist + (int64_t)blockIdx.x * nblocks;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 256 * 4) {
        const int i0 = base + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < nblocks) ? row[i0 + j] : 0u;
        const uint32_t tsum = v[0] + v[1] + v[2] + v[3];
        uint32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < wid; ++w) woff += s_warp[w];
        uint32_t excl = s_carry + woff + inc - tsum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < nblocks) row[i0 + j] = excl;
            excl += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = s_carry;
}

// ---- one digit pass in ONE launch (decoupled look-back)
//
// A tile = OS_TILE consecutive keys of the current order; tiles are handed out by an atomic ticket, so a CTA
// only ever waits for tiles whose CTAs already run (no dead-lock whatever the scheduling order).  Per digit d,
// thread d publishes the tile's count ("aggregate"), walks back over the earlier tiles adding aggregates until it
// meets a tile whose inclusive prefix is known, then publishes its own prefix.  One 64-bit word per (tile, digit):
// tag << 32 | count, tag = 2·pass + 1 (aggregate) / 2·pass + 2 (prefix) — words left by an earlier pass carry a
// smaller tag and read as "not there yet", so the table is cleared once per sort, not per pass.
// Stability: warp w owns keys [w·OS_WARP_CHUNK, +OS_WARP_CHUNK) of the tile in rounds of 32 lanes, so the order
// (tile, warp, round, lane) is the memory order; equal digits keep it.
constexpr int OS_THREADS = 256;
#ifndef RQB_OS_ITEMS
#define RQB_OS_ITEMS 16
#endif
constexpr int OS_ITEMS = RQB_OS_ITEMS;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;          // 4096 keys per CTA
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_WARP_CHUNK = 32 * OS_ITEMS;
constexpr int OS_LOOKBACK = 4;                          // earlier tiles polled per round trip
constexpr unsigned long long OS_SPIN_TIMEOUT_NS = 4ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// lanes of this warp holding the same digit as the caller (d in [0, RADIX]; RADIX = "no element")
// (measured the same with nine ballots, one per digit bit, in place of match.any, and with 8 keys per thread: a pass is
// a chain of short phases — load 3 us, count 2-4, offsets 1.5, stage 2-5, look back 2-3, store 2 — not one bottleneck;
// tools/time_dedup.py with a -DRQB_OS_TRACE build prints the phases per tile)
__device__ __forceinline__ uint32_t os_peers(uint32_t d) { return __match_any_sync(0xffffffffu, d); }
__device__ __forceinline__ unsigned long long os_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifdef RQB_OS_TRACE
__device__ unsigned long long g_os_trace[3][1024][8];
#define OS_TRACE(slot) do { if (threadIdx.x == 0 && tile < 1024 && pass < 3) g_os_trace[pass][tile][slot] = os_globaltimer_ns(); } while (0)
#else
#define OS_TRACE(slot) do { } while (0)
#endif

// FIRST: the values are the positions themselves (no value buffer is read).
// Order of work inside a tile: count → publish the aggregate → rank the tile locally and stage it SORTED in shared
// memory → look back (by now the earlier tiles have published) → copy out: consecutive threads store consecutive
// addresses of a digit's run, so a warp store touches two or three lines instead of 32.
constexpr int OS_SMEM = OS_TILE * 12 + OS_WARPS * RADIX * 4 + RADIX * 4;

template <bool FIRST>
__global__ void __launch_bounds__(OS_THREADS)
onesweep_pass_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int pass,
                     const uint32_t *__restrict__ ghist, uint32_t *__restrict__ ticket,
                     unsigned long long *__restrict__ status, const unsigned long long *__restrict__ n_dev) {
    extern __shared__ __align__(16) unsigned char os_smem[];
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    if ((int64_t)blockIdx.x * OS_TILE >= n) return;          // grid sized for the capacity: as many CTAs as tiles take a ticket
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(os_smem);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(os_smem + OS_TILE * 8);
    uint32_t(*s_cnt)[RADIX] = reinterpret_cast<uint32_t(*)[RADIX]>(os_smem + OS_TILE * 12);
    uint32_t *s_goff = reinterpret_cast<uint32_t *>(os_smem + OS_TILE * 12 + OS_WARPS * RADIX * 4);
    __shared__ uint32_t s_warp[OS_WARPS], s_warp2[OS_WARPS];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int shift = 8 * pass;
    if (tid == 0) s_tile = atomicAdd(ticket + pass, 1u);
    for (int i = tid; i < OS_WARPS * RADIX; i += OS_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    OS_TRACE(0);
    const int64_t tbase = (int64_t)tile * OS_TILE;
    const int64_t wbase = tbase + wid * OS_WARP_CHUNK;
    uint64_t k[OS_ITEMS];
    uint32_t v[OS_ITEMS];
#pragma unroll
    for (int i = 0; i < OS_ITEMS; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        const bool ok = idx < n;
        k[i] = ok ? keys_in[idx] : 0;
        v[i] = FIRST ? (uint32_t)idx : (ok ? vals_in[idx] : 0u);
    }
    // phase A: per-warp digit counts
#pragma unroll
    for (int i = 0; i < OS_ITEMS; ++i) {
        const bool ok = wbase + i * 32 + lane < n;
        const uint32_t d = ok ? digit_of(k[i], shift) : RADIX;     // RADIX = "no element"
        const uint32_t peers = os_peers(d);
        if (ok && lane == (__ffs(peers) - 1)) s_cnt[wid][d] += __popc(peers);
        __syncwarp();
    }
    // start of digit `tid` in the output = exclusive scan of the global histogram
    const uint32_t g = ghist[pass * RADIX + tid];
    uint32_t inc = g;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    OS_TRACE(1);
    uint32_t base = inc - g;
    for (int w = 0; w < wid; ++w) base += s_warp[w];
    // phase B (thread d): tile count of digit d, published at once
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) total += s_cnt[w][tid];
    const unsigned long long tag_agg = (unsigned long long)(2 * pass + 1) << 32;
    const unsigned long long tag_pre = (unsigned long long)(2 * pass + 2) << 32;
    unsigned long long *mine = status + (size_t)tile * RADIX + tid;
    st_relaxed_gpu(mine, (tile == 0 ? tag_pre : tag_agg) | total);
    // start of digit d inside the sorted tile, then per warp
    uint32_t linc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, linc, o);
        if (lane >= o) linc += t;
    }
    if (lane == 31) s_warp2[wid] = linc;
    __syncthreads();
    uint32_t lstart = linc - total;
    for (int w = 0; w < wid; ++w) lstart += s_warp2[w];
    {
        uint32_t run = lstart;
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) {
            const uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    OS_TRACE(2);
    // phase C: rank inside the tile, stage sorted
#pragma unroll
    for (int i = 0; i < OS_ITEMS; ++i) {
        const bool ok = wbase + i * 32 + lane < n;
        const uint32_t d = ok ? digit_of(k[i], shift) : RADIX;
        const uint32_t peers = os_peers(d);
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (ok) pos = s_cnt[wid][d] + below;
        __syncwarp();
        if (ok && lane == (__ffs(peers) - 1)) s_cnt[wid][d] += __popc(peers);
        __syncwarp();
        if (ok) {
            s_keys[pos] = k[i];
            s_vals[pos] = v[i];
        }
    }
    // phase D (thread d): look back over the earlier tiles
    OS_TRACE(3);
    {
        uint32_t excl = 0;
        if (tile != 0) {
            long long t = (long long)tile - 1;
            unsigned long long t0 = 0;
            bool timing = false;
            while (true) {
                unsigned long long w[OS_LOOKBACK];
#pragma unroll
                for (int j = 0; j < OS_LOOKBACK; ++j)
                    w[j] = (t - j >= 0) ? ld_relaxed_gpu(status + (size_t)(t - j) * RADIX + tid) : tag_pre;
                int adv = 0;
                bool fin = false;
#pragma unroll
                for (int j = 0; j < OS_LOOKBACK; ++j) {
                    if (fin || adv != j) continue;                 // stop at the first word that is not there yet
                    const unsigned long long tag = w[j] & 0xffffffff00000000ull;
                    if (tag == tag_pre) { excl += (uint32_t)w[j]; fin = true; ++adv; }
                    else if (tag == tag_agg) { excl += (uint32_t)w[j]; ++adv; }
                }
                if (fin) break;
                t -= adv;
                if (adv == 0) {
                    if (!timing) { t0 = os_globaltimer_ns(); timing = true; }
                    else if (os_globaltimer_ns() - t0 > OS_SPIN_TIMEOUT_NS) __trap();   // never hang the GPU
                    __nanosleep(20);
                } else {
                    timing = false;
                }
            }
            st_relaxed_gpu(mine, tag_pre | (unsigned long long)(excl + total));
        }
        s_goff[tid] = base + excl - lstart;       // sorted tile position p of digit d goes to s_goff[d] + p (mod 2^32)
    }
    __syncthreads();
    OS_TRACE(4);
    // phase E: copy out
    const int64_t left = n - tbase;
    const uint32_t cnt_tile = left < OS_TILE ? (uint32_t)left : (uint32_t)OS_TILE;
#pragma unroll
    for (int i = 0; i < OS_ITEMS; ++i) {
        const uint32_t p = i * OS_THREADS + tid;
        if (p < cnt_tile) {
            const uint64_t key = s_keys[p];
            const uint32_t gpos = s_goff[digit_of(key, shift)] + p;
            keys_out[gpos] = key;
            vals_out[gpos] = s_vals[p];
        }
    }
    OS_TRACE(5);
}

// ---------------------------------------------------------------- segmented rank over sorted keys

constexpr int SEG_THREADS = 256;
constexpr int SEG_ITEMS = 8;
constexpr int SEG_TILE = SEG_THREADS * SEG_ITEMS;

// start of the run that contains position b - 1 (b > 0, ks[b - 1] == ks[b]): the keys are sorted, so positions
// with ks[j] == key form the interval [start, b]; gallop backwards until the key changes, then bisect.
__device__ __forceinline__ long long run_start_before(const uint64_t *__restrict__ ks, int64_t b) {
    const uint64_t key = ks[b];
    long long hi = b - 1, lo = -1;                    // ks[hi] == key; ks[lo] != key (or lo == -1)
    for (long long step = 1;; step <<= 1) {
        const long long j = b - 1 - step;
        if (j < 0) break;
        if (ks[j] != key) { lo = j; break; }
        hi = j;
    }
    while (hi - lo > 1) {
        const long long mid = lo + (hi - lo) / 2;
        if (ks[mid] == key) hi = mid; else lo = mid;
    }
    return hi;
}

// rank[i] = i - (start of the run containing i), in sorted order.  Optionally also:
//   out rows:  out[item, 0..L) = codes[item], out[item, L] = rank (+ base[item])   (item = perm[i])
//   statistics: number of runs, longest run
__global__ void __launch_bounds__(SEG_THREADS)
seg_rank_kernel(const uint64_t *__restrict__ ks, const uint32_t *__restrict__ perm, int64_t n,
                int64_t *__restrict__ rank_out, PackArgs pa, int64_t *__restrict__ out,
                unsigned long long *__restrict__ stats /* [0]=runs [1]=max run */,
                uint32_t *__restrict__ rank_by_item /* rank_by_item[perm[i]] = rank, may be NULL */,
                const unsigned long long *__restrict__ n_dev) {
    __shared__ long long s_warp[SEG_THREADS / 32];
    if (n_dev) { const int64_t nd = (int64_t)*n_dev; n = nd < n ? nd : n; }
    if ((int64_t)blockIdx.x * SEG_TILE >= n) return;
    __shared__ unsigned long long s_runs, s_maxrun;
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_runs = 0; s_maxrun = 0;
        // the run that reaches into this tile from the left starts where the tile's first key starts
        const int64_t b = (int64_t)blockIdx.x * SEG_TILE;
        s_carry = (b > 0 && b < n && ks[b - 1] == ks[b]) ? run_start_before(ks, b) : -1;
    }
    // blocked arrangement: thread t owns items [base + t*SEG_ITEMS, +SEG_ITEMS)
    const int64_t base = (int64_t)blockIdx.x * SEG_TILE + (int64_t)tid * SEG_ITEMS;
    long long start[SEG_ITEMS];
    long long run = -1;
    unsigned heads = 0;
#pragma unroll
    for (int i = 0; i < SEG_ITEMS; ++i) {
        int64_t idx = base + i;
        bool head = false;
        if (idx < n) head = idx == 0 || ks[idx] != ks[idx - 1];
        if (head) { run = idx; ++heads; }
        start[i] = run;
    }
    // inclusive max-scan of `run` across threads
    long long inc = run;
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = t > inc ? t : inc;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    long long before = s_carry;
    for (int w = 0; w < wid; ++w) before = s_warp[w] > before ? s_warp[w] : before;
    long long up = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane > 0) before = up > before ? up : before;
    unsigned long long maxrun = 0;
#pragma unroll
    for (int i = 0; i < SEG_ITEMS; ++i) {
        int64_t idx = base + i;
        if (idx >= n) break;
        long long st = start[i] >= 0 ? start[i] : before;
        long long rk = idx - st;
        if (rank_out) rank_out[idx] = rk;
        if (rank_by_item) rank_by_item[perm[idx]] = (uint32_t)rk;
        if (out) {
            // the row is rebuilt from the key (the packing is lossless): no second, scattered read of the codes
            const int64_t item = perm[idx];
            const uint64_t key = ks[idx];
            int64_t *row = out + item * (pa.L + 1);
            if (pa.L == 3 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
                const longlong2 r01 = make_longlong2((long long)((key >> pa.shift[0]) & ((1ull << pa.bits[0]) - 1ull)),
                                                     (long long)((key >> pa.shift[1]) & ((1ull << pa.bits[1]) - 1ull)));
                const longlong2 r23 = make_longlong2((long long)((key >> pa.shift[2]) & ((1ull << pa.bits[2]) - 1ull)), rk);
                *reinterpret_cast<longlong2 *>(row) = r01;          // rows of 32 bytes: 16-byte aligned
                *reinterpret_cast<longlong2 *>(row + 2) = r23;
            } else {
#pragma unroll
                for (int l = 0; l < RQB200_MAX_LEVELS; ++l)
                    if (l < pa.L) row[l] = (int64_t)((key >> pa.shift[l]) & ((1ull << pa.bits[l]) - 1ull));
                row[pa.L] = rk;
            }
        }
        maxrun = (unsigned long long)(rk + 1) > maxrun ? (unsigned long long)(rk + 1) : maxrun;
    }
    if (stats) {
        for (int o = 16; o > 0; o >>= 1) {
            heads += __shfl_xor_sync(0xffffffffu, heads, o);
            unsigned long long t = __shfl_xor_sync(0xffffffffu, maxrun, o);
            maxrun = t > maxrun ? t : maxrun;
        }
        if (lane == 0) { atomicAdd(&s_runs, (unsigned long long)heads); atomicMax(&s_maxrun, maxrun); }
        __syncthreads();
        if (tid == 0) { atomicAdd(&stats[0], s_runs); atomicMax(&stats[1], s_maxrun); }
    }
}

// ---------------------------------------------------------------- collision groups (compaction)

// flags per sorted position: low 32 bits = 1 if the item belongs to a run longer than 1,
// high 32 bits = 1 if it is the head of such a run.  Exclusive u64 sum scan gives both the
// compacted item position and the group number.
__global__ void group_flags_kernel(const uint64_t *__restrict__ ks, int64_t n, uint64_t *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool same_prev = i > 0 && ks[i] == ks[i - 1];
    bool same_next = i + 1 < n && ks[i] == ks[i + 1];
    uint64_t f = 0;
    if (same_prev || same_next) f |= 1ull;
    if (!same_prev && same_next) f |= 1ull << 32;
    flags[i] = f;
}

// Rounds of the collision loop after the first: a group whose member set is EXACTLY a group that was re-encoded in the
// previous round is a fixed point — its members' codes are the output of re-encoding that very set, and the re-encode is
// a pure function of the member rows, so running it again changes nothing (typical: exact duplicates that no Sinkhorn
// pass can separate keep colliding for all 30 rounds of infer.py:112-130).  Such groups are dropped from the work list.
// Record per item: prev_first = smallest member of the group it was last re-encoded in, prev_meta = (round << 32) | size.
// Groups partition the items, so "every member carries (first, size, round - 1) of this group" ⇔ same member set.
// One thread per run head walks its run (sorted positions); runs longer than GROUP_WALK_CAP are always kept (and not
// recorded, which makes them count as changed next round as well).
constexpr int GROUP_WALK_CAP = 4096;
__global__ void group_unchanged_kernel(const uint64_t *__restrict__ ks, const uint32_t *__restrict__ perm, int64_t n,
                                       uint64_t *__restrict__ flags, int64_t *__restrict__ prev_first,
                                       int64_t *__restrict__ prev_meta, int round, unsigned long long *__restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !(flags[i] >> 32)) return;
    atomicAdd(&counts[0], 1ull);                               // every multi-member group
    const uint64_t key = ks[i];
    int64_t len = 1;
    while (i + len < n && len <= GROUP_WALK_CAP && ks[i + len] == key) ++len;
    if (len > GROUP_WALK_CAP) return;
    const int64_t first = (int64_t)perm[i];                    // stable sort: ascending item index inside a run
    const int64_t want = ((int64_t)(round - 1) << 32) | len;
    bool same = round > 0;
    for (int64_t t = 0; t < len && same; ++t) {
        const int64_t it = (int64_t)perm[i + t];
        same = prev_first[it] == first && prev_meta[it] == want;
    }
    const int64_t now = ((int64_t)round << 32) | len;
    for (int64_t t = 0; t < len; ++t) {
        const int64_t it = (int64_t)perm[i + t];
        prev_first[it] = first;
        prev_meta[it] = now;
        if (same) flags[i + t] = 0;                             // dropped from the compaction below
    }
    if (same) atomicAdd(&counts[1], 1ull);
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan64_tile_sums_kernel(const uint64_t *__restrict__ in, int64_t n, uint64_t *__restrict__ tile_sums) {
    __shared__ uint64_t s_w[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s_w[w];
        tile_sums[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of tile sums in place; total written to tile_sums[ntiles]
__global__ void __launch_bounds__(1024) scan64_tiles_kernel(uint64_t *__restrict__ tile_sums, int ntiles) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + threadIdx.x;
        uint64_t v = i < ntiles ? tile_sums[i] : 0;
        uint64_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = s_warp[lane], winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        uint64_t excl = s_carry + s_warp[wid] + inc - v;
        if (i < ntiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = s_carry;
}

// compaction: items of multi-member runs → items_out (sorted order), offsets_out[group] = start
__global__ void __launch_bounds__(SCAN_THREADS)
group_compact_kernel(const uint64_t *__restrict__ flags, const uint32_t *__restrict__ perm, int64_t n,
                     const uint64_t *__restrict__ tile_excl, int64_t *__restrict__ items_out,
                     int64_t *__restrict__ offsets_out) {
    __shared__ uint64_t s_w[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;
    uint64_t f[SCAN_ITEMS];
    uint64_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        f[i] = (base + i < n) ? flags[base + i] : 0;
        tsum += f[i];
    }
    uint64_t inc = tsum;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint64_t excl = tile_excl[blockIdx.x] + inc - tsum;
    for (int w = 0; w < wid; ++w) excl += s_w[w];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i >= n) break;
        if (f[i] & 1ull) {
            uint32_t pos = (uint32_t)(excl & 0xffffffffull);
            items_out[pos] = perm[base + i];
            if (f[i] >> 32) offsets_out[(uint32_t)(excl >> 32)] = pos;
        }
        excl += f[i];
    }
}

__global__ void write_total_offset_kernel(const uint64_t *__restrict__ total, int64_t *__restrict__ offsets) {
    uint64_t t = *total;
    offsets[(uint32_t)(t >> 32)] = (int64_t)(t & 0xffffffffull);
}

// out[i, 0..L) = codes[i], out[i, L] = rank[i]: one thread per output element, fully coalesced (the route for row widths whose
// scattered stores cannot be 16-byte vectors: L != 3)
__global__ void rows_from_rank_kernel(const int64_t *__restrict__ codes, int64_t n, int L, const uint32_t *__restrict__ rank,
                                      int64_t *__restrict__ out) {
    const int64_t total = n * (L + 1);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / (L + 1);
        const int c = (int)(e - i * (L + 1));
        out[e] = c < L ? codes[i * L + c] : (int64_t)rank[i];
    }
}
__global__ void gather_i64_kernel(const int64_t *__restrict__ in, const uint32_t *__restrict__ perm, int64_t n,
                                  int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct SortScratch {
    uint64_t *keys[2];
    uint32_t *vals[2];
    uint32_t *hist;         // owner histograms of the sharded partition (rows of nblocks)
    long long *tile_last;   // u64 tile sums of the group compaction
    uint64_t *flags;
    unsigned long long *stats;   // [0]=runs [1]=max run [2..] column min/max
    // one-sweep control block, cleared once per sort: digit histograms of all passes, tile tickets, look-back table
    uint32_t *ghist;             // [MAX_PASSES][RADIX]
    uint32_t *ticket;            // [MAX_PASSES]
    unsigned long long *status;  // [tiles][RADIX]
    size_t control_bytes;        // ghist .. end of status (for this n)
    int nblocks;
};

int carve(rqb200_model *m, int64_t n, SortScratch &sc) {
    const int nblocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    const size_t os_tiles = (size_t)((n + OS_TILE - 1) / OS_TILE);
    size_t need = 0;
    size_t o_k0 = need; need += align256(sizeof(uint64_t) * n);
    size_t o_k1 = need; need += align256(sizeof(uint64_t) * n);
    size_t o_v0 = need; need += align256(sizeof(uint32_t) * n);
    size_t o_v1 = need; need += align256(sizeof(uint32_t) * n);
    size_t o_h = need; need += align256(sizeof(uint32_t) * ((size_t)RADIX * nblocks + 2 * RADIX));
    size_t o_tl = need; need += align256(sizeof(long long) * (ntiles + 2));
    size_t o_f = need; need += align256(sizeof(uint64_t) * n);
    size_t o_s = need; need += 256;
    size_t o_g = need; need += align256(sizeof(uint32_t) * MAX_PASSES * RADIX);
    size_t o_t = need; need += 256;
    size_t o_st = need; need += align256(sizeof(unsigned long long) * os_tiles * RADIX);
    RQB_TRY(ws_reserve(m->sortws, need));
    char *p = (char *)m->sortws.ptr;
    sc.keys[0] = (uint64_t *)(p + o_k0); sc.keys[1] = (uint64_t *)(p + o_k1);
    sc.vals[0] = (uint32_t *)(p + o_v0); sc.vals[1] = (uint32_t *)(p + o_v1);
    sc.hist = (uint32_t *)(p + o_h);
    sc.tile_last = (long long *)(p + o_tl);
    sc.flags = (uint64_t *)(p + o_f);
    sc.stats = (unsigned long long *)(p + o_s);
    sc.ghist = (uint32_t *)(p + o_g);
    sc.ticket = (uint32_t *)(p + o_t);
    sc.status = (unsigned long long *)(p + o_st);
    sc.control_bytes = need - o_g;
    sc.nblocks = nblocks;
    return 0;
}

// Stable LSD radix sort of sc.keys[0] (values = positions) with 8-bit digits: one launch that counts every digit
// of every key (`codes` given: it also packs the keys), then ONE launch per digit.  Returns the buffer index
// holding the result.  `n` may be smaller than what carve() was sized for.
// `keys_src`: the first pass reads the keys from there instead of sc.keys[0] (no copy); `n_dev`: the number of keys is on the
// device (n = its upper bound: grids and the look-back table are sized for it, surplus CTAs leave at once).
int radix_sort(SortScratch &sc, int64_t n, int key_bits, cudaStream_t s, int *result_buf,
               const int64_t *codes = nullptr, const PackArgs *pa = nullptr, uint64_t *keys_src = nullptr,
               const unsigned long long *n_dev = nullptr) {
    const int npasses = (key_bits + 7) / 8;
    RQB_CHECK(npasses >= 1 && npasses <= MAX_PASSES, "key_bits=%d out of range", key_bits);
    const int os_tiles = (int)((n + OS_TILE - 1) / OS_TILE);
    const size_t control = (size_t)((char *)sc.status - (char *)sc.ghist) + sizeof(unsigned long long) * (size_t)os_tiles * RADIX;
    RQB_CUDA(cudaMemsetAsync(sc.ghist, 0, control < sc.control_bytes ? control : sc.control_bytes, s));
    int blocks = (int)((n + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    rqb::count_launch();
    uint64_t *first_in = keys_src ? keys_src : sc.keys[0];
    if (codes) pack_hist_kernel<true><<<blocks, 256, 0, s>>>(codes, n, *pa, first_in, npasses, sc.ghist, n_dev);
    else pack_hist_kernel<false><<<blocks, 256, 0, s>>>(nullptr, n, PackArgs(), first_in, npasses, sc.ghist, n_dev);
    static rqb::DeviceOnce attr_once;
    if (attr_once.first()) {
        RQB_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_SMEM));
        RQB_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_SMEM));
    }
    int cur = 0;
    for (int pass = 0; pass < npasses; ++pass) {
        rqb::count_launch();
        if (pass == 0)
            onesweep_pass_kernel<true><<<os_tiles, OS_THREADS, OS_SMEM, s>>>(first_in, nullptr, sc.keys[cur ^ 1], sc.vals[cur ^ 1], n,
                                                                             pass, sc.ghist, sc.ticket, sc.status, n_dev);
        else
            onesweep_pass_kernel<false><<<os_tiles, OS_THREADS, OS_SMEM, s>>>(sc.keys[cur], sc.vals[cur], sc.keys[cur ^ 1],
                                                                        sc.vals[cur ^ 1], n, pass, sc.ghist, sc.ticket, sc.status, n_dev);
        cur ^= 1;
    }
    RQB_LAUNCH_CHECK();
    *result_buf = cur;
    return 0;
}

int bits_for(long long maxval) {
    int b = 1;
    while (b < 63 && (maxval >> b) != 0) ++b;
    return b;
}

// decide the bit layout of the packed key: from K_host if given, else from a device min/max scan
int plan_pack(SortScratch &sc, const int64_t *codes, int64_t n, int L, const int *K_host, PackArgs &pa,
              int *key_bits, cudaStream_t s) {
    RQB_CHECK(L >= 1 && L <= RQB200_MAX_LEVELS, "L=%d out of range", L);
    long long mx[RQB200_MAX_LEVELS];
    if (K_host) {
        for (int l = 0; l < L; ++l) mx[l] = K_host[l] > 1 ? K_host[l] - 1 : 1;
    } else {
        long long init[2 * RQB200_MAX_LEVELS];
        for (int l = 0; l < RQB200_MAX_LEVELS; ++l) { init[l] = 0x7fffffffffffffffLL; init[RQB200_MAX_LEVELS + l] = -0x7fffffffffffffffLL - 1; }
        long long *dmm = (long long *)(sc.stats + 4);
        RQB_CUDA(cudaMemcpyAsync(dmm, init, sizeof(init), cudaMemcpyHostToDevice, s));
        int blocks = (int)((n + 255) / 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        rqb::count_launch();
        column_minmax_kernel<<<blocks, 256, 0, s>>>(codes, n, L, dmm, dmm + RQB200_MAX_LEVELS);
        RQB_LAUNCH_CHECK();
        long long got[2 * RQB200_MAX_LEVELS];
        RQB_CUDA(cudaMemcpyAsync(got, dmm, sizeof(got), cudaMemcpyDeviceToHost, s));
        RQB_CUDA(cudaStreamSynchronize(s));
        for (int l = 0; l < L; ++l) {
            RQB_CHECK(got[l] >= 0, "negative code in column %d", l);
            mx[l] = got[RQB200_MAX_LEVELS + l];
        }
    }
    int total = 0;
    // level 0 is the most significant field so that key order == lexicographic row order
    for (int l = L - 1; l >= 0; --l) {
        pa.shift[l] = total;
        pa.bits[l] = bits_for(mx[l]);
        total += pa.bits[l];
    }
    RQB_CHECK(total <= 64, "codes need %d key bits (> 64)", total);
    pa.L = L;
    *key_bits = total;
    return 0;
}

int sort_codes(rqb200_model *m, const int64_t *codes, int64_t n, int L, const int *K_host,
               SortScratch &sc, int *buf, cudaStream_t s, PackArgs *pa_out = nullptr) {
    RQB_CHECK(n < ((int64_t)1 << 32), "n too large for 32-bit item indices");
    RQB_TRY(carve(m, n, sc));
    PackArgs pa;
    int key_bits = 0;
    RQB_TRY(plan_pack(sc, codes, n, L, K_host, pa, &key_bits, s));
    if (pa_out) *pa_out = pa;
    return radix_sort(sc, n, key_bits, s, buf, codes, &pa);
}

// pa = nullptr: no output rows
int run_seg_rank(SortScratch &sc, int buf, int64_t n, int64_t *rank_out, const PackArgs *pa,
                 int64_t *out, bool want_stats, cudaStream_t s, uint32_t *rank_by_item = nullptr,
                 const unsigned long long *n_dev = nullptr) {
    const int ntiles = (int)((n + SEG_TILE - 1) / SEG_TILE);
    if (want_stats) RQB_CUDA(cudaMemsetAsync(sc.stats, 0, 2 * sizeof(unsigned long long), s));
    rqb::count_launch();
    seg_rank_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sc.keys[buf], sc.vals[buf], n, rank_out, pa ? *pa : PackArgs(),
                                                   pa ? out : nullptr, want_stats ? sc.stats : nullptr, rank_by_item, n_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}


// ---------------------------------------------------------------- sharded suffix dedup over NVLink peer memory
//
// New design (the reference is single-process): the catalogue is sharded by contiguous
// item ranges, one process per GPU.  The suffix of an item is its rank among ALL items with the same code in
// ascending GLOBAL item order, so equal keys have to meet: every key is routed to owner = hash(key) mod G.
// The exchange is done by the kernels themselves over peer memory (cudaIpc mappings of one symmetric block
// per rank, NVLink/NVSwitch underneath) instead of NCCL calls with host-side split sizes:
//   1. histogram of owners per CTA, scanned (the radix machinery above with the owner as the digit);
//   2. every rank stores its count row into every peer's count matrix, system-scope fence, signal, spin
//      until all rows are in → each rank knows where its keys start in every owner's receive buffer;
//   3. stable partition scatter that writes each key DIRECTLY into the owner's receive buffer (P2P stores);
//      arrival order = (source rank, local index) = ascending global item index;
//   4. signal/spin; the owner stable-sorts (key, slot) and ranks every slot inside its run of equal keys;
//   5. the owner writes the ranks back as contiguous blocks into each source's return buffer (P2P stores);
//   6. signal/spin; out[i] = (codes[i], ret[slot position of i]).
// The result is bit-identical to rqb200_suffix_dedup on the concatenated catalogue.

constexpr int SHARD_MAX_WORLD = 16;
constexpr unsigned long long SHARD_SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct ShardPeers { unsigned char *base[SHARD_MAX_WORLD]; };

struct ShardPlan {                         // device-resident, written by shard_exchange_counts_kernel
    uint32_t dest_off[SHARD_MAX_WORLD];    // where my keys start in owner d's receive buffer
    uint32_t send_off[SHARD_MAX_WORLD];    // where the block for owner d starts in my own return buffer
    uint32_t src_off[SHARD_MAX_WORLD + 1]; // receive buffer: slots [src_off[r], src_off[r+1]) came from rank r
    uint32_t ret_off[SHARD_MAX_WORLD];     // where my block starts in source r's return buffer
    unsigned long long recv_total;
    unsigned long long sort_n;             // keys the owner sorts: recv_total, or 0 when the exchange failed (shard_barrier_kernel)
    int status;                            // 0 ok, 1 receive capacity exceeded, 2 peer timeout
};

__device__ __forceinline__ uint32_t shard_owner(uint64_t key, int world) {
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    return (uint32_t)((h >> 33) % (uint64_t)world);
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// signal every peer with `epoch` and wait until every peer has signalled at least `epoch` (thread r <-> rank r)
__device__ __forceinline__ void shard_signal_and_wait(ShardPeers peers, size_t sig_off, int rank, int world,
                                                      unsigned long long epoch, int *status) {
    const int r = threadIdx.x;
    __threadfence_system();
    if (r < world) {
        st_release_sys(reinterpret_cast<unsigned long long *>(peers.base[r] + sig_off) + rank, epoch);
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(peers.base[rank] + sig_off) + r;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys(mine) < epoch) {
            if (globaltimer_ns() - t0 > SHARD_SPIN_TIMEOUT_NS) { atomicMax(status, 2); break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
}

__global__ void __launch_bounds__(SORT_THREADS)
shard_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int world, uint32_t *__restrict__ hist, int nblocks) {
    __shared__ uint32_t s_hist[SHARD_MAX_WORLD];
    if (threadIdx.x < SHARD_MAX_WORLD) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    uint32_t local[SHARD_MAX_WORLD];
#pragma unroll
    for (int d = 0; d < SHARD_MAX_WORLD; ++d) local[d] = 0;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int64_t idx = base + i * SORT_THREADS + threadIdx.x;
        if (idx < n) {
            const uint32_t o = shard_owner(keys[idx], world);
#pragma unroll
            for (int d = 0; d < SHARD_MAX_WORLD; ++d) local[d] += (o == (uint32_t)d);
        }
    }
#pragma unroll
    for (int d = 0; d < SHARD_MAX_WORLD; ++d) {
        uint32_t v = local[d];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_hist[d], v);
    }
    __syncthreads();
    if (threadIdx.x < world) hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
}

// one CTA of 32 threads: publish my counts row, barrier, derive the plan
__global__ void __launch_bounds__(32)
shard_exchange_counts_kernel(ShardPeers peers, size_t sig_off, size_t counts_off, int rank, int world,
                             unsigned long long epoch, const uint32_t *__restrict__ my_counts /* digit totals */,
                             unsigned long long cap_recv, ShardPlan *__restrict__ plan) {
    __shared__ uint32_t M[SHARD_MAX_WORLD][SHARD_MAX_WORLD];
    const int t = threadIdx.x;
    if (t == 0) plan->status = 0;
    __syncthreads();
    if (t < world) {
        uint32_t *row = reinterpret_cast<uint32_t *>(peers.base[t] + counts_off) + rank * SHARD_MAX_WORLD;
        for (int d = 0; d < world; ++d) row[d] = my_counts[d];
    }
    shard_signal_and_wait(peers, sig_off, rank, world, epoch, &plan->status);
    const volatile uint32_t *mat = reinterpret_cast<const volatile uint32_t *>(peers.base[rank] + counts_off);
    for (int i = t; i < world * world; i += 32) M[i / world][i % world] = mat[(i / world) * SHARD_MAX_WORLD + (i % world)];
    __syncthreads();
    if (t < world) {
        uint32_t a = 0;
        for (int r = 0; r < rank; ++r) a += M[r][t];
        plan->dest_off[t] = a;                                   // my keys in owner t's buffer
        uint32_t b = 0;
        for (int d = 0; d < t; ++d) b += M[rank][d];
        plan->send_off[t] = b;                                   // block for owner t in my return buffer
        uint32_t c = 0;
        for (int r = 0; r < t; ++r) c += M[r][rank];
        plan->src_off[t] = c;                                    // slots of source t in my receive buffer
        uint32_t e = 0;
        for (int d = 0; d < rank; ++d) e += M[t][d];
        plan->ret_off[t] = e;                                    // my block in source t's return buffer
        unsigned long long col = 0;
        for (int r = 0; r < world; ++r) col += M[r][t];
        if (col > cap_recv) atomicMax(&plan->status, 1);         // same verdict on every rank (same matrix)
    }
    if (t == 0) {
        unsigned long long tot = 0;
        for (int r = 0; r < world; ++r) tot += M[r][rank];
        plan->src_off[world] = (uint32_t)tot;
        plan->recv_total = tot;
    }
}

// stable partition by owner, written straight into the owners' receive buffers (peer memory)
__global__ void __launch_bounds__(SORT_THREADS)
shard_scatter_kernel(const uint64_t *__restrict__ keys_in, int64_t n, int world, ShardPeers peers, size_t recv_off,
                     const uint32_t *__restrict__ hist, int nblocks, const ShardPlan *__restrict__ plan,
                     uint32_t *__restrict__ slotpos) {
    __shared__ uint32_t s_cnt[SORT_WARPS][SHARD_MAX_WORLD];
    __shared__ uint32_t s_dest[SHARD_MAX_WORLD], s_send[SHARD_MAX_WORLD];
    if (plan->status != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < SORT_WARPS * SHARD_MAX_WORLD) (&s_cnt[0][0])[tid] = 0;
    if (tid < SHARD_MAX_WORLD) { s_dest[tid] = tid < world ? plan->dest_off[tid] : 0; s_send[tid] = tid < world ? plan->send_off[tid] : 0; }
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * SORT_TILE + wid * WARP_CHUNK;
    uint64_t k[SORT_ITEMS];
    uint32_t own[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int64_t idx = wbase + i * 32 + lane;
        const bool ok = idx < n;
        k[i] = ok ? keys_in[idx] : 0;
        own[i] = ok ? shard_owner(k[i], world) : 0xffu;
        uint32_t peers_m = __match_any_sync(0xffffffffu, own[i]);
        if (ok && lane == (__ffs(peers_m) - 1)) s_cnt[wid][own[i]] += __popc(peers_m);
        __syncwarp();
    }
    __syncthreads();
    if (tid < world) {
        uint32_t run = hist[(int64_t)tid * nblocks + blockIdx.x];       // exclusive scan over CTAs of this owner
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int64_t idx = wbase + i * 32 + lane;
        const bool ok = idx < n;
        uint32_t peers_m = __match_any_sync(0xffffffffu, own[i]);
        uint32_t below = __popc(peers_m & ((1u << lane) - 1u));
        uint32_t rel = 0;
        if (ok) rel = s_cnt[wid][own[i]] + below;
        __syncwarp();
        if (ok && lane == (__ffs(peers_m) - 1)) s_cnt[wid][own[i]] += __popc(peers_m);
        __syncwarp();
        if (ok) {
            uint64_t *dst = reinterpret_cast<uint64_t *>(peers.base[own[i]] + recv_off);
            dst[s_dest[own[i]] + rel] = k[i];
            slotpos[idx] = s_send[own[i]] + rel;
        }
    }
}

__global__ void __launch_bounds__(32)
shard_barrier_kernel(ShardPeers peers, size_t sig_off, int rank, int world, unsigned long long epoch,
                     ShardPlan *__restrict__ plan) {
    shard_signal_and_wait(peers, sig_off, rank, world, epoch, &plan->status);
    if (threadIdx.x == 0) plan->sort_n = plan->status == 0 ? plan->recv_total : 0ull;
}

// the owner returns the ranks: slot s of source r → r's return buffer at ret_off[r] + (s - src_off[r])
__global__ void shard_return_kernel(const uint32_t *__restrict__ rank_slot, int world, ShardPeers peers, size_t ret_off_bytes,
                                    const ShardPlan *__restrict__ plan) {
    if (plan->status != 0) return;
    const unsigned long long total = plan->recv_total;
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < total;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        int r = 0;
        while (r + 1 < world && s >= plan->src_off[r + 1]) ++r;
        uint32_t *dst = reinterpret_cast<uint32_t *>(peers.base[r] + ret_off_bytes);
        dst[plan->ret_off[r] + (uint32_t)(s - plan->src_off[r])] = rank_slot[s];
    }
}

__global__ void shard_finalize_kernel(const int64_t *__restrict__ codes, int64_t n, int L, const uint32_t *__restrict__ ret,
                                      const uint32_t *__restrict__ slotpos, const ShardPlan *__restrict__ plan,
                                      int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int l = 0; l < L; ++l) out[i * (L + 1) + l] = codes[i * L + l];
    out[i * (L + 1) + L] = plan->status == 0 ? (int64_t)ret[slotpos[i]] : -1;
}

}  // namespace

}  // namespace rqb

using namespace rqb;

#ifdef RQB_OS_TRACE
extern "C" __attribute__((visibility("default"))) int rqb200_debug_os_trace(unsigned long long *host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, rqb::g_os_trace, sizeof(rqb::g_os_trace));
}
#endif

extern "C" int rqb200_suffix_dedup(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L,
                                   const int *K_host, int64_t *out_dev, int64_t *n_distinct_host,
                                   int64_t *max_group_host, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RQB_CHECK(m != nullptr, "model is NULL");
    if (n == 0) {
        if (n_distinct_host) *n_distinct_host = 0;
        if (max_group_host) *max_group_host = 0;
        return 0;
    }
    RQB_CUDA(cudaSetDevice(m->device));
    SortScratch sc;
    int buf = 0;
    ProfScope ps(PROF_DEDUP, s);
    const bool stats = n_distinct_host || max_group_host;
    PackArgs pa;
    RQB_TRY(sort_codes(m, codes_dev, n, L, K_host, sc, &buf, s, &pa));
    if (L == 3 && (reinterpret_cast<uintptr_t>(out_dev) & 15) == 0) {
        // rows of 32 bytes: the rank kernel scatters them itself as two 16-byte stores per row
        RQB_TRY(run_seg_rank(sc, buf, n, nullptr, &pa, out_dev, stats, s));
    } else {
        // other widths: scatter only the 4-byte rank by item, then write the rows in item order, coalesced (40-byte rows as five
        // scalar stores each took 67 us at 1 M items against 22 us for the vector form)
        uint32_t *rank_by_item = reinterpret_cast<uint32_t *>(sc.flags);
        RQB_TRY(run_seg_rank(sc, buf, n, nullptr, nullptr, nullptr, stats, s, rank_by_item));
        int blocks = (int)((n * (L + 1) + 255) / 256);
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        rqb::count_launch();
        rows_from_rank_kernel<<<blocks, 256, 0, s>>>(codes_dev, n, L, rank_by_item, out_dev);
        RQB_LAUNCH_CHECK();
    }
    if (stats) {
        unsigned long long h[2];
        RQB_CUDA(cudaMemcpyAsync(h, sc.stats, sizeof(h), cudaMemcpyDeviceToHost, s));
        RQB_CUDA(cudaStreamSynchronize(s));
        if (n_distinct_host) *n_distinct_host = (int64_t)h[0];
        if (max_group_host) *max_group_host = (int64_t)h[1];
    }
    return 0;
}

static int collision_groups_impl(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L, const int *K_host,
                                 int64_t *items_dev, int64_t *offsets_dev, int64_t *n_groups_host, int64_t *n_items_host,
                                 int64_t *max_group_host, int64_t *prev_first_dev, int64_t *prev_meta_dev, int round,
                                 int64_t *n_groups_total_host, cudaStream_t s) {
    RQB_CHECK(m != nullptr, "model is NULL");
    RQB_CHECK(n_groups_host && n_items_host, "count outputs are required");
    *n_groups_host = 0; *n_items_host = 0;
    if (max_group_host) *max_group_host = 0;
    if (n_groups_total_host) *n_groups_total_host = 0;
    if (n == 0) return 0;
    RQB_CUDA(cudaSetDevice(m->device));
    SortScratch sc;
    int buf = 0;
    RQB_TRY(sort_codes(m, codes_dev, n, L, K_host, sc, &buf, s));
    if (max_group_host) RQB_TRY(run_seg_rank(sc, buf, n, nullptr, nullptr, nullptr, true, s));
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    uint64_t *tile_sums = (uint64_t *)sc.tile_last;
    rqb::count_launch();
    group_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc.keys[buf], n, sc.flags);
    unsigned long long *gcounts = sc.stats + 24;      // stats: [0]=runs [1]=max run [4..19]=column min/max [24..25]=these
    if (prev_first_dev) {
        RQB_CUDA(cudaMemsetAsync(gcounts, 0, 2 * sizeof(unsigned long long), s));
        rqb::count_launch();
        group_unchanged_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc.keys[buf], sc.vals[buf], n, sc.flags, prev_first_dev,
                                                                           prev_meta_dev, round, gcounts);
    }
    rqb::count_launch();
    scan64_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(sc.flags, n, tile_sums);
    rqb::count_launch();
    scan64_tiles_kernel<<<1, 1024, 0, s>>>(tile_sums, ntiles);
    rqb::count_launch();
    group_compact_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(sc.flags, sc.vals[buf], n, tile_sums, items_dev, offsets_dev);
    rqb::count_launch();
    write_total_offset_kernel<<<1, 1, 0, s>>>(tile_sums + ntiles, offsets_dev);
    RQB_LAUNCH_CHECK();
    uint64_t total = 0;
    unsigned long long h[2] = {0, 0}, gc[2] = {0, 0};
    RQB_CUDA(cudaMemcpyAsync(&total, tile_sums + ntiles, sizeof(total), cudaMemcpyDeviceToHost, s));
    if (max_group_host) RQB_CUDA(cudaMemcpyAsync(h, sc.stats, sizeof(h), cudaMemcpyDeviceToHost, s));
    if (prev_first_dev) RQB_CUDA(cudaMemcpyAsync(gc, gcounts, sizeof(gc), cudaMemcpyDeviceToHost, s));
    RQB_CUDA(cudaStreamSynchronize(s));
    *n_items_host = (int64_t)(total & 0xffffffffull);
    *n_groups_host = (int64_t)(total >> 32);
    if (max_group_host) *max_group_host = (int64_t)h[1];
    if (n_groups_total_host) *n_groups_total_host = prev_first_dev ? (int64_t)gc[0] : *n_groups_host;
    return 0;
}

extern "C" int rqb200_collision_groups(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L,
                                       const int *K_host, int64_t *items_dev, int64_t *offsets_dev,
                                       int64_t *n_groups_host, int64_t *n_items_host,
                                       int64_t *max_group_host, void *stream) {
    return collision_groups_impl(m, codes_dev, n, L, K_host, items_dev, offsets_dev, n_groups_host, n_items_host, max_group_host,
                                 nullptr, nullptr, 0, nullptr, (cudaStream_t)stream);
}

extern "C" int rqb200_collision_groups_changed(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L, const int *K_host,
                                               int64_t *prev_first_dev, int64_t *prev_meta_dev, int round,
                                               int64_t *items_dev, int64_t *offsets_dev, int64_t *n_groups_host,
                                               int64_t *n_items_host, int64_t *max_group_host, int64_t *n_groups_total_host,
                                               void *stream) {
    RQB_CHECK(prev_first_dev && prev_meta_dev, "the per-item group records are required");
    RQB_CHECK(round >= 0, "round < 0");
    return collision_groups_impl(m, codes_dev, n, L, K_host, items_dev, offsets_dev, n_groups_host, n_items_host, max_group_host,
                                 prev_first_dev, prev_meta_dev, round, n_groups_total_host, (cudaStream_t)stream);
}

extern "C" int rqb200_pack_keys(const int64_t *codes_dev, int64_t n, int L, const int *K_host,
                                uint64_t *keys_dev, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RQB_CHECK(K_host != nullptr, "K_host is required");
    RQB_CHECK(L >= 1 && L <= RQB200_MAX_LEVELS, "L=%d out of range", L);
    if (n == 0) return 0;
    PackArgs pa;
    int total = 0;
    for (int l = L - 1; l >= 0; --l) {
        pa.shift[l] = total;
        pa.bits[l] = bits_for(K_host[l] > 1 ? K_host[l] - 1 : 1);
        total += pa.bits[l];
    }
    RQB_CHECK(total <= 64, "codes need %d key bits (> 64)", total);
    pa.L = L;
    rqb::count_launch();
    pack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(codes_dev, n, pa, keys_dev, nullptr);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_sort_pairs(rqb200_model *m, uint64_t *keys_dev, int64_t *vals_dev, int64_t n,
                                 int key_bits, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RQB_CHECK(m != nullptr, "model is NULL");
    RQB_CHECK(key_bits >= 1 && key_bits <= 64, "key_bits=%d out of range", key_bits);
    RQB_CHECK(n < ((int64_t)1 << 32), "n too large");
    if (n == 0) return 0;
    RQB_CUDA(cudaSetDevice(m->device));
    SortScratch sc;
    RQB_TRY(carve(m, n, sc));
    RQB_CUDA(cudaMemcpyAsync(sc.keys[0], keys_dev, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s));
    // values travel as 32-bit positions (the first digit pass makes them); the caller's int64 payload is permuted at the end
    int buf = 0;
    RQB_TRY(radix_sort(sc, n, key_bits, s, &buf));
    RQB_CUDA(cudaMemcpyAsync(keys_dev, sc.keys[buf], sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s));
    // permute payload: tmp = vals_dev[perm]
    int64_t *tmp = (int64_t *)sc.flags;
    rqb::count_launch();
    gather_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(vals_dev, sc.vals[buf], n, tmp);
    RQB_LAUNCH_CHECK();
    RQB_CUDA(cudaMemcpyAsync(vals_dev, tmp, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, s));
    return 0;
}

extern "C" int rqb200_segment_rank(rqb200_model *m, const uint64_t *sorted_keys_dev, int64_t n,
                                   int64_t *rank_dev, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RQB_CHECK(m != nullptr, "model is NULL");
    if (n == 0) return 0;
    RQB_CUDA(cudaSetDevice(m->device));
    const int ntiles = (int)((n + SEG_TILE - 1) / SEG_TILE);
    rqb::count_launch();
    seg_rank_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sorted_keys_dev, nullptr, n, rank_dev, PackArgs(), nullptr, nullptr, nullptr,
                                                   nullptr);
    RQB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- sharded dedup: C ABI

struct rqb200_shard {
    rqb200_model *m = nullptr;
    int rank = 0, world = 1;
    int64_t cap_local = 0, cap_recv = 0;
    unsigned char *block = nullptr;          // this rank's symmetric block (cudaMalloc, exported by cudaIpc)
    size_t block_bytes = 0;
    size_t sig_off = 0, counts_off = 0, recv_off = 0, ret_off = 0;
    rqb::ShardPeers peers;
    bool connected = false;
    unsigned long long epoch = 0;
    rqb::ShardPlan *plan = nullptr;          // device
};

extern "C" int rqb200_shard_create(rqb200_shard **out, rqb200_model *m, int rank, int world,
                                   int64_t max_local_items, int64_t max_recv_items) {
    RQB_CHECK(out != nullptr && m != nullptr, "NULL argument");
    RQB_CHECK(world >= 1 && world <= SHARD_MAX_WORLD, "world=%d out of range (1..%d)", world, SHARD_MAX_WORLD);
    RQB_CHECK(rank >= 0 && rank < world, "rank %d out of range", rank);
    RQB_CHECK(max_local_items >= 0 && max_local_items < ((int64_t)1 << 32), "max_local_items out of range");
    if (max_recv_items <= 0) max_recv_items = 2 * max_local_items + 65536;
    RQB_CHECK(max_recv_items < ((int64_t)1 << 32), "max_recv_items out of range");
    RQB_CUDA(cudaSetDevice(m->device));
    rqb200_shard *sh = new (std::nothrow) rqb200_shard();
    if (!sh) { set_error("out of host memory"); return RQB200_ENOMEM; }
    sh->m = m; sh->rank = rank; sh->world = world;
    sh->cap_local = max_local_items; sh->cap_recv = max_recv_items;
    sh->sig_off = 0;
    sh->counts_off = 256;
    sh->recv_off = 4096;
    sh->ret_off = sh->recv_off + align256(sizeof(uint64_t) * (size_t)max_recv_items);
    sh->block_bytes = sh->ret_off + align256(sizeof(uint32_t) * (size_t)max_local_items) + 256;
    for (int r = 0; r < SHARD_MAX_WORLD; ++r) sh->peers.base[r] = nullptr;
    if (cudaMalloc(&sh->block, sh->block_bytes) != cudaSuccess || cudaMalloc(&sh->plan, sizeof(rqb::ShardPlan)) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaMalloc(%zu) for the symmetric block failed", sh->block_bytes);
        if (sh->block) cudaFree(sh->block);
        delete sh;
        return RQB200_ENOMEM;
    }
    RQB_CUDA(cudaMemset(sh->block, 0, sh->block_bytes));
    RQB_CUDA(cudaMemset(sh->plan, 0, sizeof(rqb::ShardPlan)));
    RQB_CUDA(cudaDeviceSynchronize());
    sh->peers.base[rank] = sh->block;
    sh->connected = world == 1;
    *out = sh;
    return 0;
}

extern "C" int rqb200_shard_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int rqb200_shard_get_handle(rqb200_shard *sh, void *handle_out) {
    RQB_CHECK(sh != nullptr && handle_out != nullptr, "NULL argument");
    RQB_CUDA(cudaSetDevice(sh->m->device));
    cudaIpcMemHandle_t h;
    RQB_CUDA(cudaIpcGetMemHandle(&h, sh->block));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

extern "C" int rqb200_shard_connect(rqb200_shard *sh, const void *all_handles) {
    RQB_CHECK(sh != nullptr && all_handles != nullptr, "NULL argument");
    RQB_CUDA(cudaSetDevice(sh->m->device));
    for (int r = 0; r < sh->world; ++r) {
        if (r == sh->rank || sh->peers.base[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)all_handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        RQB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        sh->peers.base[r] = (unsigned char *)p;
    }
    sh->connected = true;
    return 0;
}

extern "C" void rqb200_shard_destroy(rqb200_shard *sh) {
    if (!sh) return;
    cudaSetDevice(sh->m->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < sh->world; ++r)
        if (r != sh->rank && sh->peers.base[r]) cudaIpcCloseMemHandle(sh->peers.base[r]);
    if (sh->block) cudaFree(sh->block);
    if (sh->plan) cudaFree(sh->plan);
    (void)cudaGetLastError();
    delete sh;
}

extern "C" int rqb200_shard_suffix_dedup(rqb200_shard *sh, const int64_t *codes_dev, int64_t n, int L,
                                         const int *K_host, int64_t *out_dev, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RQB_CHECK(sh != nullptr && K_host != nullptr, "NULL argument");
    RQB_CHECK(sh->connected, "rqb200_shard_connect has not been called");
    RQB_CHECK(n >= 0 && n <= sh->cap_local, "n=%lld exceeds the shard capacity %lld", (long long)n, (long long)sh->cap_local);
    RQB_CHECK(n == 0 || (codes_dev != nullptr && out_dev != nullptr), "NULL buffer");
    rqb200_model *m = sh->m;
    RQB_CUDA(cudaSetDevice(m->device));
    ProfScope ps(PROF_DEDUP, s);
    SortScratch sc;
    const int64_t nmax = (n > sh->cap_recv ? n : sh->cap_recv) + 1;
    RQB_TRY(carve(m, nmax, sc));
    const int nblocks_max = sc.nblocks;
    PackArgs pa;
    int key_bits = 0;
    RQB_TRY(plan_pack(sc, codes_dev, n, L, K_host, pa, &key_bits, s));
    uint32_t *digit_total = sc.hist + (size_t)RADIX * nblocks_max;
    uint32_t *slotpos = reinterpret_cast<uint32_t *>(sc.flags);
    uint32_t *rank_slot = slotpos + (size_t)nmax;
    const int world = sh->world, rank = sh->rank;
    const int nb1 = (int)((n + SORT_TILE - 1) / SORT_TILE);
    RQB_CUDA(cudaMemsetAsync(digit_total, 0, sizeof(uint32_t) * RADIX, s));
    if (n > 0) {
        rqb::count_launch();
        pack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(codes_dev, n, pa, sc.keys[0], nullptr);
        rqb::count_launch();
        shard_hist_kernel<<<nb1, SORT_THREADS, 0, s>>>(sc.keys[0], n, world, sc.hist, nb1);
        rqb::count_launch();
        radix_scan_rows_kernel<<<world, 256, 0, s>>>(sc.hist, nb1, digit_total);
    }
    rqb::count_launch();
    shard_exchange_counts_kernel<<<1, 32, 0, s>>>(sh->peers, sh->sig_off, sh->counts_off, rank, world, ++sh->epoch, digit_total,
                                                  (unsigned long long)sh->cap_recv, sh->plan);
    if (n > 0) {
        rqb::count_launch();
        shard_scatter_kernel<<<nb1, SORT_THREADS, 0, s>>>(sc.keys[0], n, world, sh->peers, sh->recv_off, sc.hist, nb1, sh->plan,
                                                         slotpos);
    }
    rqb::count_launch();
    shard_barrier_kernel<<<1, 32, 0, s>>>(sh->peers, sh->sig_off, rank, world, ++sh->epoch, sh->plan);
    RQB_LAUNCH_CHECK();
    // The owner's part runs on a count that stays on the device (plan->sort_n): grids and the look-back table are sized for the
    // receive capacity, the keys are sorted straight out of the receive buffer, nothing waits for the host in mid-step.
    {
        const int64_t Rcap = sh->cap_recv;
        const unsigned long long *n_dev = &sh->plan->sort_n;
        int buf = 0;
        RQB_TRY(radix_sort(sc, Rcap, key_bits, s, &buf, nullptr, nullptr, reinterpret_cast<uint64_t *>(sh->block + sh->recv_off), n_dev));
        RQB_TRY(run_seg_rank(sc, buf, Rcap, nullptr, nullptr, nullptr, false, s, rank_slot, n_dev));
        int blocks = (int)((Rcap + 255) / 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        rqb::count_launch();
        shard_return_kernel<<<blocks, 256, 0, s>>>(rank_slot, world, sh->peers, sh->ret_off, sh->plan);
    }
    rqb::count_launch();
    shard_barrier_kernel<<<1, 32, 0, s>>>(sh->peers, sh->sig_off, rank, world, ++sh->epoch, sh->plan);
    if (n > 0) {
        rqb::count_launch();
        shard_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(codes_dev, n, L, reinterpret_cast<const uint32_t *>(sh->block + sh->ret_off),
                                                                         slotpos, sh->plan, out_dev);
    }
    RQB_LAUNCH_CHECK();
    int status = 0;
    RQB_CUDA(cudaMemcpyAsync(&status, &sh->plan->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    RQB_CUDA(cudaStreamSynchronize(s));
    if (status == 1) { set_error("sharded dedup: a key owner would receive more than max_recv_items=%lld keys", (long long)sh->cap_recv); return RQB200_ENOMEM; }
    if (status != 0) { set_error("sharded dedup: timed out waiting for a peer rank"); return RQB200_ESTATE; }
    return 0;
}

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

// per-CTA digit histogram of the current key order → hist[d * nblocks + block]
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ hist,
                  int nblocks) {
    __shared__ uint32_t s_hist[RADIX];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int64_t idx = base + i * SORT_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&s_hist[digit_of(keys[idx], shift)], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
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

__global__ void __launch_bounds__(256) radix_scan_digits_kernel(const uint32_t *__restrict__ digit_total,
                                                                uint32_t *__restrict__ digit_base) {
    __shared__ uint32_t s_warp[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t v = digit_total[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < wid; ++w) woff += s_warp[w];
    digit_base[threadIdx.x] = woff + inc - v;
}

// stable scatter: warp w owns keys [base + w*WARP_CHUNK, +WARP_CHUNK) in rounds of 32 lanes, so the
// order (warp, round, lane) equals the memory order.
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n,
                     int shift, const uint32_t *__restrict__ hist, const uint32_t *__restrict__ digit_base,
                     int nblocks) {
    __shared__ uint32_t s_cnt[SORT_WARPS][RADIX];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * SORT_TILE + wid * WARP_CHUNK;
    uint64_t k[SORT_ITEMS];
    uint32_t v[SORT_ITEMS];
    bool ok[SORT_ITEMS];
    // phase A: per-warp digit counts
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int64_t idx = wbase + i * 32 + lane;
        ok[i] = idx < n;
        k[i] = ok[i] ? keys_in[idx] : 0;
        v[i] = ok[i] ? vals_in[idx] : 0;
        uint32_t d = ok[i] ? digit_of(k[i], shift) : RADIX;      // RADIX = "no element"
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (ok[i] && lane == (__ffs(peers) - 1)) s_cnt[wid][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase B: thread d turns counts into start offsets per warp
    {
        uint32_t run = digit_base[tid] + hist[(int64_t)tid * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase C: ranked scatter
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        uint32_t d = ok[i] ? digit_of(k[i], shift) : RADIX;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t below = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (ok[i]) pos = s_cnt[wid][d] + below;
        __syncwarp();
        if (ok[i] && lane == (__ffs(peers) - 1)) s_cnt[wid][d] += __popc(peers);
        __syncwarp();
        if (ok[i]) {
            keys_out[pos] = k[i];
            vals_out[pos] = v[i];
        }
    }
}

// ---------------------------------------------------------------- segmented rank over sorted keys

constexpr int SEG_THREADS = 256;
constexpr int SEG_ITEMS = 8;
constexpr int SEG_TILE = SEG_THREADS * SEG_ITEMS;

// tile summary: index of the last run head inside the tile (or -1)
__global__ void __launch_bounds__(SEG_THREADS)
seg_tile_last_head_kernel(const uint64_t *__restrict__ ks, int64_t n, long long *__restrict__ tile_last) {
    __shared__ long long s_best;
    if (threadIdx.x == 0) s_best = -1;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SEG_TILE;
    long long best = -1;
#pragma unroll
    for (int i = 0; i < SEG_ITEMS; ++i) {
        int64_t idx = base + i * SEG_THREADS + threadIdx.x;
        if (idx < n) {
            bool head = idx == 0 || ks[idx] != ks[idx - 1];
            if (head) best = idx;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t > best ? t : best;
    }
    if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0) tile_last[blockIdx.x] = s_best;
}

// running max over tiles: carry[b] = last head strictly before tile b (single CTA, sequential chunks)
__global__ void __launch_bounds__(1024) seg_carry_kernel(const long long *__restrict__ tile_last,
                                                         long long *__restrict__ carry, int ntiles) {
    __shared__ long long s_warp[32];
    __shared__ long long s_run;
    if (threadIdx.x == 0) s_run = -1;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + threadIdx.x;
        long long v = i < ntiles ? tile_last[i] : -1;
        long long inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = t > inc ? t : inc;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            long long w = s_warp[lane];
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w = t > w ? t : w;
            }
            s_warp[lane] = w;   // inclusive max over warps 0..lane
        }
        __syncthreads();
        long long prev_warps = wid > 0 ? s_warp[wid - 1] : -1;
        long long run = s_run;
        long long incl = inc;
        incl = prev_warps > incl ? prev_warps : incl;
        incl = run > incl ? run : incl;
        // exclusive value = max of everything before i
        long long up = __shfl_up_sync(0xffffffffu, inc, 1);
        long long excl = lane > 0 ? up : -1;
        excl = prev_warps > excl ? prev_warps : excl;
        excl = run > excl ? run : excl;
        if (i < ntiles) carry[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = incl;
        __syncthreads();
    }
}

// rank[i] = i - (start of the run containing i), in sorted order.  Optionally also:
//   out rows:  out[item, 0..L) = codes[item], out[item, L] = rank (+ base[item])   (item = perm[i])
//   statistics: number of runs, longest run
__global__ void __launch_bounds__(SEG_THREADS)
seg_rank_kernel(const uint64_t *__restrict__ ks, const uint32_t *__restrict__ perm, int64_t n,
                const long long *__restrict__ carry, int64_t *__restrict__ rank_out,
                const int64_t *__restrict__ codes, int L, int64_t *__restrict__ out,
                unsigned long long *__restrict__ stats /* [0]=runs [1]=max run */) {
    __shared__ long long s_warp[SEG_THREADS / 32];
    __shared__ unsigned long long s_runs, s_maxrun;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_runs = 0; s_maxrun = 0; }
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
    long long before = carry[blockIdx.x];
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
        if (out) {
            int64_t item = perm[idx];
            for (int l = 0; l < L; ++l) out[item * (L + 1) + l] = codes[item * L + l];
            out[item * (L + 1) + L] = rk;
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

__global__ void iota_u32_kernel(uint32_t *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
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
    uint32_t *hist;
    long long *tile_last;   // also reused as u64 tile sums
    long long *carry;
    uint64_t *flags;
    unsigned long long *stats;   // [0]=runs [1]=max run [2..] column min/max
    int nblocks;
};

int carve(rqb200_model *m, int64_t n, SortScratch &sc) {
    const int nblocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
    const int ntiles = (int)((n + SEG_TILE - 1) / SEG_TILE);
    size_t need = 0;
    size_t o_k0 = need; need += align256(sizeof(uint64_t) * n);
    size_t o_k1 = need; need += align256(sizeof(uint64_t) * n);
    size_t o_v0 = need; need += align256(sizeof(uint32_t) * n);
    size_t o_v1 = need; need += align256(sizeof(uint32_t) * n);
    size_t o_h = need; need += align256(sizeof(uint32_t) * ((size_t)RADIX * nblocks + 2 * RADIX));
    size_t o_tl = need; need += align256(sizeof(long long) * (ntiles + 2));
    size_t o_c = need; need += align256(sizeof(long long) * (ntiles + 2));
    size_t o_f = need; need += align256(sizeof(uint64_t) * n);
    size_t o_s = need; need += 256;
    RQB_TRY(ws_reserve(m->sortws, need));
    char *p = (char *)m->sortws.ptr;
    sc.keys[0] = (uint64_t *)(p + o_k0); sc.keys[1] = (uint64_t *)(p + o_k1);
    sc.vals[0] = (uint32_t *)(p + o_v0); sc.vals[1] = (uint32_t *)(p + o_v1);
    sc.hist = (uint32_t *)(p + o_h);
    sc.tile_last = (long long *)(p + o_tl);
    sc.carry = (long long *)(p + o_c);
    sc.flags = (uint64_t *)(p + o_f);
    sc.stats = (unsigned long long *)(p + o_s);
    sc.nblocks = nblocks;
    return 0;
}

// sorts sc.keys[0]/vals[0]; returns the buffer index holding the result
int radix_sort(SortScratch &sc, int64_t n, int key_bits, cudaStream_t s, int *result_buf) {
    int cur = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        rqb::count_launch();
        radix_hist_kernel<<<sc.nblocks, SORT_THREADS, 0, s>>>(sc.keys[cur], n, shift, sc.hist, sc.nblocks);
        uint32_t *digit_total = sc.hist + (size_t)RADIX * sc.nblocks;
        uint32_t *digit_base = digit_total + RADIX;
        rqb::count_launch();
        radix_scan_rows_kernel<<<RADIX, 256, 0, s>>>(sc.hist, sc.nblocks, digit_total);
        rqb::count_launch();
        radix_scan_digits_kernel<<<1, 256, 0, s>>>(digit_total, digit_base);
        rqb::count_launch();
        radix_scatter_kernel<<<sc.nblocks, SORT_THREADS, 0, s>>>(sc.keys[cur], sc.vals[cur], sc.keys[cur ^ 1],
                                                               sc.vals[cur ^ 1], n, shift, sc.hist, digit_base, sc.nblocks);
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
        total += bits_for(mx[l]);
    }
    RQB_CHECK(total <= 64, "codes need %d key bits (> 64)", total);
    pa.L = L;
    *key_bits = total;
    return 0;
}

int sort_codes(rqb200_model *m, const int64_t *codes, int64_t n, int L, const int *K_host,
               SortScratch &sc, int *buf, cudaStream_t s) {
    RQB_CHECK(n < ((int64_t)1 << 32), "n too large for 32-bit item indices");
    RQB_TRY(carve(m, n, sc));
    PackArgs pa;
    int key_bits = 0;
    RQB_TRY(plan_pack(sc, codes, n, L, K_host, pa, &key_bits, s));
    rqb::count_launch();
    pack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(codes, n, pa, sc.keys[0], sc.vals[0]);
    RQB_LAUNCH_CHECK();
    return radix_sort(sc, n, key_bits, s, buf);
}

int run_seg_rank(SortScratch &sc, int buf, int64_t n, int64_t *rank_out, const int64_t *codes, int L,
                 int64_t *out, bool want_stats, cudaStream_t s) {
    const int ntiles = (int)((n + SEG_TILE - 1) / SEG_TILE);
    if (want_stats) RQB_CUDA(cudaMemsetAsync(sc.stats, 0, 2 * sizeof(unsigned long long), s));
    rqb::count_launch();
    seg_tile_last_head_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sc.keys[buf], n, sc.tile_last);
    rqb::count_launch();
    seg_carry_kernel<<<1, 1024, 0, s>>>(sc.tile_last, sc.carry, ntiles);
    rqb::count_launch();
    seg_rank_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sc.keys[buf], sc.vals[buf], n, sc.carry, rank_out, codes, L,
                                                   out, want_stats ? sc.stats : nullptr);
    RQB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

}  // namespace rqb

using namespace rqb;

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
    RQB_TRY(sort_codes(m, codes_dev, n, L, K_host, sc, &buf, s));
    const bool stats = n_distinct_host || max_group_host;
    RQB_TRY(run_seg_rank(sc, buf, n, nullptr, codes_dev, L, out_dev, stats, s));
    if (stats) {
        unsigned long long h[2];
        RQB_CUDA(cudaMemcpyAsync(h, sc.stats, sizeof(h), cudaMemcpyDeviceToHost, s));
        RQB_CUDA(cudaStreamSynchronize(s));
        if (n_distinct_host) *n_distinct_host = (int64_t)h[0];
        if (max_group_host) *max_group_host = (int64_t)h[1];
    }
    return 0;
}

extern "C" int rqb200_collision_groups(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L,
                                       const int *K_host, int64_t *items_dev, int64_t *offsets_dev,
                                       int64_t *n_groups_host, int64_t *n_items_host,
                                       int64_t *max_group_host, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RQB_CHECK(m != nullptr, "model is NULL");
    RQB_CHECK(n_groups_host && n_items_host, "count outputs are required");
    *n_groups_host = 0; *n_items_host = 0;
    if (max_group_host) *max_group_host = 0;
    if (n == 0) return 0;
    RQB_CUDA(cudaSetDevice(m->device));
    SortScratch sc;
    int buf = 0;
    RQB_TRY(sort_codes(m, codes_dev, n, L, K_host, sc, &buf, s));
    if (max_group_host) RQB_TRY(run_seg_rank(sc, buf, n, nullptr, nullptr, L, nullptr, true, s));
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    uint64_t *tile_sums = (uint64_t *)sc.tile_last;
    rqb::count_launch();
    group_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc.keys[buf], n, sc.flags);
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
    unsigned long long h[2] = {0, 0};
    RQB_CUDA(cudaMemcpyAsync(&total, tile_sums + ntiles, sizeof(total), cudaMemcpyDeviceToHost, s));
    if (max_group_host) RQB_CUDA(cudaMemcpyAsync(h, sc.stats, sizeof(h), cudaMemcpyDeviceToHost, s));
    RQB_CUDA(cudaStreamSynchronize(s));
    *n_items_host = (int64_t)(total & 0xffffffffull);
    *n_groups_host = (int64_t)(total >> 32);
    if (max_group_host) *max_group_host = (int64_t)h[1];
    return 0;
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
        total += bits_for(K_host[l] > 1 ? K_host[l] - 1 : 1);
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
    // values travel as 32-bit positions; the caller's int64 payload is permuted at the end
    rqb::count_launch();
    iota_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc.vals[0], n);
    RQB_LAUNCH_CHECK();
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
    RQB_TRY(ws_reserve(m->misc, sizeof(long long) * 2 * (size_t)(ntiles + 2) + 512));
    long long *tile_last = (long long *)m->misc.ptr;
    long long *carry = tile_last + ntiles + 2;
    rqb::count_launch();
    seg_tile_last_head_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sorted_keys_dev, n, tile_last);
    rqb::count_launch();
    seg_carry_kernel<<<1, 1024, 0, s>>>(tile_last, carry, ntiles);
    rqb::count_launch();
    seg_rank_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sorted_keys_dev, nullptr, n, carry, rank_dev, nullptr, 0,
                                                   nullptr, nullptr);
    RQB_LAUNCH_CHECK();
    return 0;
}

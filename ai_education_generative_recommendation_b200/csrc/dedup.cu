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
                unsigned long long *__restrict__ stats /* [0]=runs [1]=max run */,
                uint32_t *__restrict__ rank_by_item /* rank_by_item[perm[i]] = rank, may be NULL */) {
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
        if (rank_by_item) rank_by_item[perm[idx]] = (uint32_t)rk;
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
                 int64_t *out, bool want_stats, cudaStream_t s, uint32_t *rank_by_item = nullptr) {
    const int ntiles = (int)((n + SEG_TILE - 1) / SEG_TILE);
    if (want_stats) RQB_CUDA(cudaMemsetAsync(sc.stats, 0, 2 * sizeof(unsigned long long), s));
    rqb::count_launch();
    seg_tile_last_head_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sc.keys[buf], n, sc.tile_last);
    rqb::count_launch();
    seg_carry_kernel<<<1, 1024, 0, s>>>(sc.tile_last, sc.carry, ntiles);
    rqb::count_launch();
    seg_rank_kernel<<<ntiles, SEG_THREADS, 0, s>>>(sc.keys[buf], sc.vals[buf], n, sc.carry, rank_out, codes, L,
                                                   out, want_stats ? sc.stats : nullptr, rank_by_item);
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
    RQB_TRY(sort_codes(m, codes_dev, n, L, K_host, sc, &buf, s));
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
    if (max_group_host) RQB_TRY(run_seg_rank(sc, buf, n, nullptr, nullptr, L, nullptr, true, s));
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
                                                   nullptr, nullptr, nullptr);
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
    rqb::ShardPlan hp;
    RQB_CUDA(cudaMemcpyAsync(&hp, sh->plan, sizeof(hp), cudaMemcpyDeviceToHost, s));
    RQB_CUDA(cudaStreamSynchronize(s));
    const int64_t R = hp.status == 0 ? (int64_t)hp.recv_total : 0;
    if (R > 0) {
        RQB_CUDA(cudaMemcpyAsync(sc.keys[0], sh->block + sh->recv_off, sizeof(uint64_t) * (size_t)R, cudaMemcpyDeviceToDevice, s));
        rqb::count_launch();
        iota_u32_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(sc.vals[0], R);
        sc.nblocks = (int)((R + SORT_TILE - 1) / SORT_TILE);
        int buf = 0;
        RQB_TRY(radix_sort(sc, R, key_bits, s, &buf));
        RQB_TRY(run_seg_rank(sc, buf, R, nullptr, nullptr, 0, nullptr, false, s, rank_slot));
        int blocks = (int)((R + 255) / 256);
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
    if (hp.status != 0) status = hp.status;
    if (status == 1) { set_error("sharded dedup: a key owner would receive more than max_recv_items=%lld keys", (long long)sh->cap_recv); return RQB200_ENOMEM; }
    if (status != 0) { set_error("sharded dedup: timed out waiting for a peer rank"); return RQB200_ESTATE; }
    return 0;
}

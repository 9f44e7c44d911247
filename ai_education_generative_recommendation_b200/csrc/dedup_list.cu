// dedup_list.cu — the suffix column of infer.py:152-163 WITHOUT a sort: hash table + per-code item lists.
//
// STATUS: EXPERIMENTAL, OFF BY DEFAULT (RQB200_DEDUP_LIST=1 or debug flag 8192 selects it).  Written at the end of
// round 1 from the round-1 measurements; compiles for sm_100a, has NOT run on a B200 yet.  tools/check_dedup_list.py
// compares it with the sort path (dedup.cu), which stays the default and the fallback.
//
// Why: out[i, L] = #{j < i : codes[j] == codes[i]} is today a packed-key LSD radix sort (3-5 passes of four launches)
// plus a segmented rank, about 25 launches and 0.17-0.20 ms per 1 M items (4 % of the HBM roofline for the 56 bytes
// per item the stage has to move: it is launch- and pass-bound).  The rank does not need a total order, only the
// members of each code's group:
//   1. one memset (0xFF) of the table: 16 bytes per slot, 2 slots per item (32 MB per 1 M items — L2-resident);
//   2. insert kernel, one thread per item: open-addressing insert of the packed key (atomicCAS on the key word), then
//      the item is pushed on its slot's list (atomicExch on `head`, old head → next[i]) and the slot's member count is
//      bumped (atomicAdd); distinct codes and the largest group come out of the same atomics;
//   3. rank kernel, one thread per item: items alone in their slot write suffix 0 without touching the list; the others
//      walk their slot's list and count the members with a smaller item index.  List order is whatever the atomics
//      produced, the count is not: the result is deterministic and bit-identical to the sort path.
// Three launches, ≈ 120 MB of traffic per 1 M items, output rows written in item order (coalesced) instead of key order.
//
// The walk is quadratic in the group size, so groups above LIST_WALK_CAP are not ranked here: the caller reads the
// largest group size (it synchronises anyway to return the statistics) and falls back to the sort path when the cap
// was exceeded (collapsed codebooks: every item in one group).  Without requested statistics, without K_host, or with
// keys of 64 bits the sort path is used directly (this path would add a synchronisation / cannot mark empty slots).
#include <stdlib.h>

#include "common.cuh"

namespace rqb {

namespace {

constexpr uint32_t NIL = 0xFFFFFFFFu;
constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int LIST_WALK_CAP = 512;

struct ListPack {
    int shift[RQB200_MAX_LEVELS];
    int L;
};

__device__ __forceinline__ uint64_t mix64(uint64_t k) {       // splitmix64 finaliser
    k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull;
    k ^= k >> 27; k *= 0x94d049bb133111ebull;
    k ^= k >> 31;
    return k;
}

// table layout: [keys: M x u64][head: M x u32][count: M x u32], all bytes 0xFF when empty (count = members - 1)
__global__ void __launch_bounds__(256)
list_insert_kernel(const int64_t *__restrict__ codes, int64_t n, ListPack pa, unsigned long long *__restrict__ tkeys,
                   uint32_t *__restrict__ head, uint32_t *__restrict__ count, uint32_t mask, uint32_t *__restrict__ next,
                   uint32_t *__restrict__ slot_of, unsigned long long *__restrict__ stats) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool fresh = false;
    unsigned members = 0;
    if (i < n) {
        unsigned long long k = 0;
#pragma unroll
        for (int l = 0; l < RQB200_MAX_LEVELS; ++l)           // fixed trip count: pa.shift stays in the constant bank
            if (l < pa.L) k |= (unsigned long long)codes[i * pa.L + l] << pa.shift[l];
        uint32_t slot = (uint32_t)mix64(k) & mask;
        while (true) {
            const unsigned long long prev = atomicCAS(&tkeys[slot], EMPTY_KEY, k);
            if (prev == EMPTY_KEY) { fresh = true; break; }
            if (prev == k) break;
            slot = (slot + 1) & mask;
        }
        next[i] = atomicExch(&head[slot], (uint32_t)i);
        members = atomicAdd(&count[slot], 1u) + 2u;          // 0xFFFFFFFF + 2 wraps to 1 for the first member
        slot_of[i] = slot;
    }
    // statistics: distinct codes (slots claimed) and the largest group, one atomic per warp
    const unsigned fresh_mask = __ballot_sync(0xFFFFFFFFu, fresh);
    unsigned mx = members;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    if ((threadIdx.x & 31) == 0) {
        if (fresh_mask) atomicAdd(&stats[0], (unsigned long long)__popc(fresh_mask));
        if (mx > 1) atomicMax(&stats[1], (unsigned long long)mx);
    }
}

__global__ void __launch_bounds__(256)
list_rank_kernel(const int64_t *__restrict__ codes, int64_t n, int L, const uint32_t *__restrict__ head,
                 const uint32_t *__restrict__ count, const uint32_t *__restrict__ next, const uint32_t *__restrict__ slot_of,
                 int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = slot_of[i];
    const uint32_t members = count[slot] + 1u;
    int64_t rank = 0;
    if (members > 1u && members <= (uint32_t)LIST_WALK_CAP) {
        uint32_t j = head[slot];
        for (uint32_t step = 0; step < members && j != NIL; ++step) {
            rank += (j < (uint32_t)i) ? 1 : 0;
            j = next[j];
        }
    }
    for (int l = 0; l < L; ++l) out[i * (L + 1) + l] = codes[i * L + l];
    out[i * (L + 1) + L] = rank;          // groups above the cap are redone by the sort path (see the header comment)
}

int bits_for_list(long long maxval) {
    int b = 1;
    while (b < 63 && (maxval >> b) != 0) ++b;
    return b;
}

inline size_t align256l(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int tc_debug_flags();     // encode_tc.cu

bool dedup_list_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("RQB200_DEDUP_LIST"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1 || (tc_debug_flags() & 8192) != 0;
}

// Returns 0 and *done = 1 when out / the statistics are complete; *done = 0 when the caller has to run the sort path
// (unsupported key layout, or a group above LIST_WALK_CAP).  Synchronises the stream (it returns the statistics).
int suffix_dedup_list(rqb200_model *m, const int64_t *codes, int64_t n, int L, const int *K_host, int64_t *out,
                      int64_t *n_distinct_host, int64_t *max_group_host, cudaStream_t s, int *done) {
    *done = 0;
    if (!K_host || n >= ((int64_t)1 << 31) || L < 1 || L > RQB200_MAX_LEVELS) return 0;
    ListPack pa;
    int total = 0;
    for (int l = L - 1; l >= 0; --l) {
        pa.shift[l] = total;
        total += bits_for_list(K_host[l] > 1 ? K_host[l] - 1 : 1);
    }
    pa.L = L;
    if (total > 63) return 0;                                  // an all-ones key would look like an empty slot
    uint64_t slots = 1;
    while (slots < 2 * (uint64_t)n) slots <<= 1;
    if (slots < 1024) slots = 1024;
    const size_t table_bytes = (size_t)slots * 16;
    size_t need = 0;
    const size_t o_table = need; need += align256l(table_bytes);
    const size_t o_next = need; need += align256l(sizeof(uint32_t) * (size_t)n);
    const size_t o_slot = need; need += align256l(sizeof(uint32_t) * (size_t)n);
    const size_t o_stats = need; need += 256;
    RQB_TRY(ws_reserve(m->sortws, need));
    char *p = (char *)m->sortws.ptr;
    unsigned long long *tkeys = (unsigned long long *)(p + o_table);
    uint32_t *head = (uint32_t *)(p + o_table + (size_t)slots * 8);
    uint32_t *count = head + slots;
    uint32_t *next = (uint32_t *)(p + o_next);
    uint32_t *slot_of = (uint32_t *)(p + o_slot);
    unsigned long long *stats = (unsigned long long *)(p + o_stats);
    RQB_CUDA(cudaMemsetAsync(tkeys, 0xFF, table_bytes, s));
    RQB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned long long), s));
    const unsigned grid = (unsigned)((n + 255) / 256);
    count_launch();
    list_insert_kernel<<<grid, 256, 0, s>>>(codes, n, pa, tkeys, head, count, (uint32_t)(slots - 1), next, slot_of, stats);
    count_launch();
    list_rank_kernel<<<grid, 256, 0, s>>>(codes, n, L, head, count, next, slot_of, out);
    RQB_LAUNCH_CHECK();
    unsigned long long h[2];
    RQB_CUDA(cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, s));
    RQB_CUDA(cudaStreamSynchronize(s));
    const unsigned long long max_group = h[1] > 1 ? h[1] : 1;  // stats[1] is only raised by groups of two or more
    if (max_group > (unsigned long long)LIST_WALK_CAP) return 0;
    if (n_distinct_host) *n_distinct_host = (int64_t)h[0];
    if (max_group_host) *max_group_host = (int64_t)max_group;
    *done = 1;
    return 0;
}

}  // namespace rqb

"""Host twin of the experimental sort-free suffix dedup (csrc/dedup_list.cu): the same table layout, hash, list push and
rank-by-walk, executed with the inserts in a random completion order (atomics on the GPU finish in any order), against
the oracle's suffix column.  Checks the algorithm and its statistics; the CUDA kernels themselves are compared with the
sort path on the GPU by tools/check_dedup_list.py."""
import numpy as np
import pytest

M64 = (1 << 64) - 1
NIL = 0xFFFFFFFF


def _mix64(k):
    k ^= k >> 30; k = (k * 0xbf58476d1ce4e5b9) & M64
    k ^= k >> 27; k = (k * 0x94d049bb133111eb) & M64
    k ^= k >> 31
    return k


def _bits_for(maxval):
    b = 1
    while b < 63 and (maxval >> b) != 0:
        b += 1
    return b


def list_dedup_twin(codes, Ks, order):
    n, L = codes.shape
    shift, total = [0] * L, 0
    for l in range(L - 1, -1, -1):
        shift[l] = total
        total += _bits_for(Ks[l] - 1 if Ks[l] > 1 else 1)
    assert total <= 63
    slots = 1
    while slots < 2 * n:
        slots <<= 1
    slots = max(slots, 1024)
    mask = slots - 1
    tkeys, head, count = [M64] * slots, [NIL] * slots, [NIL] * slots        # memset 0xFF
    nxt, slot_of = [0] * n, [0] * n
    distinct, max_group = 0, 1
    for i in order:                                                          # list_insert_kernel, one thread per item
        k = 0
        for l in range(L):
            k |= int(codes[i, l]) << shift[l]
        s = (_mix64(k) & 0xFFFFFFFF) & mask
        while True:
            if tkeys[s] == M64:
                tkeys[s] = k
                distinct += 1
                break
            if tkeys[s] == k:
                break
            s = (s + 1) & mask
        nxt[i], head[s] = head[s], i                                         # atomicExch
        members = (count[s] + 2) & 0xFFFFFFFF                                # atomicAdd returns the old value
        count[s] = (count[s] + 1) & 0xFFFFFFFF
        max_group = max(max_group, members)
        slot_of[i] = s
    out = np.empty((n, L + 1), dtype=np.int64)
    out[:, :L] = codes
    for i in range(n):                                                       # list_rank_kernel
        s = slot_of[i]
        members = (count[s] + 1) & 0xFFFFFFFF
        rank = 0
        if members > 1:
            j, step = head[s], 0
            while step < members and j != NIL:
                rank += j < i
                j = nxt[j]
                step += 1
        out[i, L] = rank
    return out, distinct, max_group


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.mark.parametrize("n,L,K,seed", [(1, 3, 8, 0), (707, 3, 8, 1), (5000, 4, 12, 2), (3000, 2, 2, 3), (4000, 5, 4096, 4)])
def test_list_dedup_twin_matches_oracle(oracle, n, L, K, seed):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, K, size=(n, L), dtype=np.int64)
    got, distinct, max_group = list_dedup_twin(codes, [K] * L, rng.permutation(n))
    ref = oracle.suffix_dedup(codes)
    assert np.array_equal(got, ref)
    uniq, counts = np.unique(codes, axis=0, return_counts=True)
    assert distinct == len(uniq) and max_group == counts.max()


from hypothesis import given, settings, strategies as st       # noqa: E402


@settings(max_examples=80, deadline=None)
@given(st.integers(1, 80), st.integers(1, 5), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
def test_list_dedup_twin_property(n, L, K, seed):
    """Any codes, any completion order of the inserts: suffix[i] = #{j < i : codes[j] == codes[i]} (infer.py:152-163)."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, K, size=(n, L), dtype=np.int64)
    got, distinct, max_group = list_dedup_twin(codes, [K] * L, rng.permutation(n))
    want = np.array([sum(1 for j in range(i) if (codes[j] == codes[i]).all()) for i in range(n)], dtype=np.int64)
    assert np.array_equal(got[:, :L], codes) and np.array_equal(got[:, L], want)
    assert distinct == len({tuple(r) for r in codes.tolist()})

"""Pins the Lloyd iterations of the k-means codebook init against scikit-learn itself — the third-party code the
reference calls at RQ-VAE/models/layers.py:77 (`KMeans(n_clusters=K, max_iter=iters).fit(x)`).  TEST INFRASTRUCTURE ONLY.

With a GIVEN `init` array (and n_init=1, algorithm="lloyd", the defaults otherwise) scikit-learn's fit is deterministic,
so its centres are a golden vector for `oracle.kmeans_lloyd` and for the GPU Lloyd (`kmeans_gpu.kmeans_fit(init=…)`).
What stays unpinned is the SEEDING (k-means++ consumes numpy's global RNG inside scikit-learn): the product seeds with
its own greedy k-means++ on the device.  scikit-learn works in fp32 on centred data, the oracle accumulates in fp64, so
parity is a tolerance (fp32 rounding, 1e-6 relative on the centres) — plus, where a sample is equidistant from two
centres at fp32 precision, one differing label, which moves the centres of the two (tiny) clusters involved.

Run in the build container:  python oracle/make_golden_kmeans.py   → tests/golden/kmeans_sklearn.npz
"""
import os
import sys

import numpy as np
import sklearn
from sklearn.cluster import KMeans

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)
from oracle import oracle as O                                            # noqa: E402
from ai_education_generative_recommendation_b200 import synth              # noqa: E402

CASES = [  # (n, e, K, max_iter, n_total of the synthetic catalogue the rows are cut from)
    (8192, 32, 256, 10, 1_000_000),      # BASELINE config 2 codebook shape, first training batch
    (4096, 64, 256, 10, 1_000_000),      # config 3
    (707, 32, 8, 50, 40_000),            # config 1 (main.py defaults: 8 codes, 50 iterations … on the whole catalogue)
    (16384, 64, 1024, 5, 1_000_000),     # config 5 (29 clusters run empty after the first iteration → relocation)
    (3000, 16, 128, 6, 40_000),          # small case with empty clusters, for the host-logic test
]


def case_data(n, e, K, n_total):
    x = synth.synth_items(2024, 1, n, e, n_total)
    init = np.ascontiguousarray(x[(np.arange(K) * (n // K)) % n])
    if n == 3000:                       # the small case: four duplicated initial centres → four empty clusters at once
        init[1::32] = init[0::32]
    return x, init


if __name__ == "__main__":
    out = {"sklearn_version": sklearn.__version__, "n_cases": len(CASES)}
    for i, (n, e, K, iters, n_total) in enumerate(CASES):
        x, init = case_data(n, e, K, n_total)
        km = KMeans(n_clusters=K, init=init, n_init=1, max_iter=iters, algorithm="lloyd").fit(x)
        c = km.cluster_centers_.astype(np.float32)
        mine = O.kmeans_lloyd(x, init, iters)
        # compare as SETS of centres (rows sorted): the relocation of empty clusters may hand the far samples out in another order
        key = lambda a: a[np.lexsort(a.T[::-1])]
        rel = np.abs(key(c) - key(mine)).max(1) / np.abs(c).max()
        close = rel <= 1e-6
        print(f"case {i}: n={n} e={e} K={K} max_iter={iters}: scikit-learn ran {km.n_iter_} iterations; oracle centres within 1e-6 "
              f"of scikit-learn's: {int(close.sum())} / {K} (max rel diff {rel.max():.2e})")
        assert close.mean() >= 0.97
        out[f"meta{i}"] = np.array([n, e, K, iters, n_total], dtype=np.int64)
        out[f"centers{i}"] = c
        out[f"n_iter{i}"] = np.int64(km.n_iter_)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "kmeans_sklearn.npz"), **out)

"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference/RQ-VAE)
on CPU in the build container, and checks the C oracle against it bit for bit.

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists):  python oracle/make_golden.py
The GPU box never runs this; it only reads the committed fixtures.

What gets pinned:
  * get_indices / forward of the reference RQVAE on seeded synthetic inputs at the BASELINE config
    shapes (C1, C2, C3, C5 slices): codes, latent bits, x_q bits, losses;
  * the reference's whole `infer()` (pass 1, Sinkhorn re-encode rounds, suffix dedup, np.save) run
    verbatim through an in-memory h5py stand-in on a 707-item catalogue (BASELINE config 1);
  * the Sinkhorn branch of VectorQuantizer.forward on random groups;
  * the shipped artifact RQ-VAE/semantic_id_viz/course_semantic_id_alignment.csv (suffix rule).
Codebooks come from the reference's own k-means init (scikit-learn) and are stored in the fixture;
encoder/decoder weights and inputs are regenerated from integer hashing (package synth.py).
"""
import csv
import io
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/RQ-VAE"
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import oracle as O                                            # noqa: E402
from ai_education_generative_recommendation_b200 import synth              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 2024


_H5_ARRAYS = {}


def install_h5py_standin(arrays):
    """h5py is not installed; serve f['item_embs'][:] / f['meta'][()] (vision_data.py:18-21) from memory.
    The stand-in module is created once (the reference binds `h5py` at import time); later calls swap its contents."""
    _H5_ARRAYS.clear()
    _H5_ARRAYS.update(arrays)
    if "h5py" in sys.modules and getattr(sys.modules["h5py"], "_standin", False):
        return
    arrays = _H5_ARRAYS
    mod = types.ModuleType("h5py")
    mod._standin = True

    class _DS:
        def __init__(self, v): self.v = v
        def __getitem__(self, k): return self.v if k == () else self.v[k]

    class File:
        def __init__(self, path, mode="r"): self.d = arrays[path]
        def __enter__(self): return self
        def __exit__(self, *a): return False
        def __getitem__(self, k): return _DS(self.d[k])

    mod.File = File
    sys.modules["h5py"] = mod


def ref_model(cfg, sd_np):
    from models.rqvae import RQVAE
    m = RQVAE(in_dim=cfg["in_dim"], num_emb_list=cfg["num_emb_list"], e_dim=cfg["e_dim"], layers=cfg["layers"],
              dropout_prob=0.0, bn=False, loss_type="mse", quant_loss_weight=cfg.get("quant_loss_weight", 1.0),
              kmeans_init=True, kmeans_iters=cfg["kmeans_iters"], sk_epsilons=cfg["sk_epsilons"],
              sk_iters=cfg["sk_iters"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()})
    return m


def kmeans_init_codebooks(m, x_init):
    """The reference's own codebook init: first training-mode forward triggers init_emb level by level
    (vq.py:67-68, rq.py:45-47) — scikit-learn KMeans under numpy's global RNG (train.py:256 seeds it)."""
    np.random.seed(SEED)
    torch.manual_seed(SEED)
    m.train()
    with torch.no_grad():
        m(torch.from_numpy(x_init), use_sk=False)
    m.eval()
    return [q.embedding.weight.detach().numpy().copy() for q in m.rq.vq_layers]


def weights_of(sd_np, prefix, n):
    return ([sd_np[f"{prefix}.mlp_layers.{1 + 3 * i}.weight"] for i in range(n)],
            [sd_np[f"{prefix}.mlp_layers.{1 + 3 * i}.bias"] for i in range(n)])


def make_slice(name, cfg, n_rows, n_total, n_init):
    sd = synth.synth_state_dict(SEED, cfg["in_dim"], cfg["layers"], cfg["e_dim"], cfg["num_emb_list"])
    m = ref_model(cfg, sd)
    x_init = synth.synth_items(SEED, 1, n_init, cfg["in_dim"], n_total)
    cbs = kmeans_init_codebooks(m, x_init)
    x = synth.synth_items(SEED, 0, n_rows, cfg["in_dim"], n_total)
    with torch.no_grad():
        xt = torch.from_numpy(x)
        codes = m.get_indices(xt, use_sk=False).numpy()
        out, rq_loss, idx2 = m(xt, use_sk=False)
        total, recon = m.compute_loss(out, rq_loss, xs=xt)
        z = m.encoder(xt)
        x_q, _, _ = m.rq(z, use_sk=False)
    assert np.array_equal(codes, idx2.numpy())
    nl = len(cfg["layers"]) + 1
    ew, eb = weights_of(sd, "encoder", nl)
    dw, db = weights_of(sd, "decoder", nl)
    # ---- pin the oracle against the reference
    zo = O.mlp(x, ew, eb)
    assert np.array_equal(zo.view(np.int32), z.numpy().view(np.int32)), f"{name}: oracle encoder != reference"
    io_, xqo, ssq, _ = O.quantize(zo, cbs)
    assert np.array_equal(io_, codes), f"{name}: oracle codes != reference"
    assert np.array_equal(xqo.view(np.int32), x_q.numpy().view(np.int32)), f"{name}: oracle x_q != reference"
    outo = O.mlp(xqo, dw, db)
    assert np.array_equal(outo.view(np.int32), out.numpy().view(np.int32)), f"{name}: oracle decoder != reference"
    lo = O.rq_loss(ssq, n_rows, cfg["e_dim"], 0.25)
    assert abs(lo - float(rq_loss)) <= 1e-5 * abs(float(rq_loss)), (lo, float(rq_loss))
    keep = 64
    np.savez_compressed(
        os.path.join(GOLD, f"{name}.npz"),
        cfg=json.dumps(cfg), n_rows=n_rows, n_total=n_total, seed=SEED,
        **{f"codebook{l}": c for l, c in enumerate(cbs)},
        codes=codes.astype(np.int16), z_head=z.numpy()[:keep], xq_head=x_q.numpy()[:keep], out_head=out.numpy()[:8],
        rq_loss=np.float64(rq_loss), recon_loss=np.float64(recon), total_loss=np.float64(total))
    print(f"{name}: {n_rows} rows, distinct codes {len(np.unique(codes, axis=0))}, rq_loss {float(rq_loss):.6f} "
          f"recon {float(recon):.6f}  — oracle bit-equal")


TIE_GAP = 1e-10


def make_infer(name, cfg, n, n_total, init_rows, batch_sizes, dup=None, exact_final=False):
    """The reference's infer() verbatim (infer.py:44-184) on an n-item synthetic catalogue, plus the round-by-round
    trace of its re-encode loop; asserts that the oracle driver in group order reproduces BOTH exactly."""
    in_dim = cfg["in_dim"]
    sd = synth.synth_state_dict(SEED, in_dim, cfg["layers"], cfg["e_dim"], cfg["num_emb_list"])
    m = ref_model(cfg, sd)
    x = synth.synth_items(SEED, 0, n, in_dim, n_total)
    if dup is not None:                              # exact duplicates → guaranteed collision groups
        for dst, src, cnt in dup:
            x[dst:dst + cnt] = x[src:src + cnt]
    cbs = kmeans_init_codebooks(m, x[init_rows[0]:init_rows[1]])
    tmp = tempfile.mkdtemp()
    ckdir = os.path.join(tmp, "ckpt")
    os.makedirs(ckdir)
    torch.save({"args": {}, "epoch": 0, "best_loss": 0.0, "best_collision_rate": 0.0,
                "state_dict": m.state_dict(), "optimizer": {}}, os.path.join(ckdir, "best_collision_model.pth"),
               pickle_protocol=4)
    install_h5py_standin({"mem.h5": {"item_embs": x, "meta": json.dumps({"n": n}).encode("utf-8")}})
    import infer as ref_infer
    out_file = os.path.join(tmp, "out", "codes.npy")
    outs = {}
    for bs in batch_sizes:
        params = {"data_path": "mem.h5", "ckpt_dir": ckdir, "semantic_id_file": out_file, "device": "cpu",
                  "num_emb_list": cfg["num_emb_list"], "e_dim": cfg["e_dim"], "layers": cfg["layers"], "dropout": 0.1,
                  "batch_normalize": False, "loss_type": "mse", "quant_loss_weight": cfg.get("quant_loss_weight", 1.0),
                  "kmeans_init": True, "kmeans_iters": cfg["kmeans_iters"], "sk_epsilons": cfg["sk_epsilons"],
                  "sk_iters": cfg["sk_iters"], "batch_size": bs, "num_workers": 0}
        stdout = sys.stdout
        sys.stdout = io.StringIO()
        try:
            ref_infer.infer(params)
        finally:
            log, sys.stdout = sys.stdout.getvalue(), stdout
        outs[bs] = np.load(out_file)
        print(f"{name}: reference infer(batch_size={bs}):", outs[bs].shape, outs[bs].dtype,
              [l for l in log.splitlines() if "Collision Rate" in l or "Max number" in l])
    golden = outs[batch_sizes[0]]
    for bs in batch_sizes[1:]:
        print(f"{name}: infer() outputs identical for batch {batch_sizes[0]} vs {bs}:", np.array_equal(golden, outs[bs]))
    Lv = len(cfg["num_emb_list"])
    ew, eb = weights_of(sd, "encoder", len(cfg["layers"]) + 1)
    # Round-by-round trace with the reference's own model calls (the loop of infer.py:112-130): the codes
    # after every round, so each round can be checked as a pure function of the previous one.
    with torch.no_grad():
        cur = torch.cat([m.get_indices(torch.from_numpy(x[i:i + 4096]), use_sk=False)
                         for i in range(0, n, 4096)]).numpy()
    z = O.mlp(x, ew, eb)
    assert np.array_equal(cur, O.quantize(z, cbs, want_xq=False)[0])
    for vq in m.rq.vq_layers[:-1]:
        vq.sk_epsilon = 0.0
    eps = [0.0] * (Lv - 1) + [cfg["sk_epsilons"][-1]]
    trace = [cur.copy()]
    bad_groups = n_groups = bad_rows = tie_rows_equal = 0
    sizes = {}
    flagged = []          # per round: items whose Sinkhorn arg-max is an fp64 tie (top-2 relative gap <= TIE_GAP)
    for rnd in range(30):
        groups = O.collision_groups(cur)
        if not groups:
            break
        nxt = cur.copy()
        flag = []
        for g in groups:
            with torch.no_grad():
                nxt[g] = m.get_indices(torch.from_numpy(x[g]), use_sk=True).numpy()
            gaps = []
            mine = O.reencode_group(x[g], ew, eb, cbs, eps, cfg["sk_iters"], gaps=gaps)
            tie = gaps[-1] <= TIE_GAP
            flag.append(np.asarray(g)[tie])
            diff = (mine != nxt[g]).any(1)
            # Everything up to the fp64 Sinkhorn matrix is restated bit for bit (latent, distances); what is left is the
            # reference's own fp64 exp / reduction rounding, which only shows where a row's two best codes tie in fp64.
            assert np.array_equal(mine[:, :-1], nxt[g][:, :-1]), f"{name}: prefix codes differ in a group of {len(g)}"
            assert not (diff & ~tie).any(), f"{name}: a row outside the fp64-tie set differs (group of {len(g)}, gaps {gaps[-1][diff]})"
            bad_groups += int(diff.any()); bad_rows += int(diff.sum()); tie_rows_equal += int((tie & ~diff).sum())
            sizes[len(g)] = sizes.get(len(g), 0) + 1
        n_groups += len(groups)
        flagged.append(np.concatenate(flag) if flag else np.zeros(0, dtype=np.int64))
        cur = nxt
        trace.append(cur.copy())
    assert np.array_equal(O.suffix_dedup(cur), golden), "trace loop does not reproduce verbatim infer()"
    got, stats = O.generate_codes(x, ew, eb, cbs, cfg["sk_epsilons"], cfg["sk_iters"], group_order=True)
    nd = int((got != golden).any(1).sum())
    old, _ = O.generate_codes(x, ew, eb, cbs, cfg["sk_epsilons"], cfg["sk_iters"], group_order=False)
    n_flag = int(sum(len(f) for f in flagged))
    print(f"{name}: {n_groups} group calls, sizes {min(sizes)}..{max(sizes)}; rows differing from the reference inside a round: "
          f"{bad_rows} (in {bad_groups} groups), ALL of them fp64 ties of the reference's own Sinkhorn matrix (tie set: {n_flag} "
          f"row-rounds, {tie_rows_equal} of them equal anyway); final ids differing: {nd} / {n}; catalogue-order "
          f"re-quantisation (the round-1 driver) would differ in {int((old != golden).any(1).sum())} rows", stats)
    if exact_final:
        assert bad_rows == 0 and nd == 0, "the group-order restatement must reproduce the reference's infer() exactly"
    # the trace is stored as pass-1 codes + per-round overwrites (a round only rewrites the members of its groups)
    small = np.int8 if max(cfg["num_emb_list"]) <= 127 else np.int16
    chg_items, chg_codes = [], []
    for a, b in zip(trace[:-1], trace[1:]):
        rows = np.nonzero((a != b).any(1))[0]
        chg_items.append(rows.astype(np.int32))
        chg_codes.append(b[rows].astype(small))
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), cfg=json.dumps(cfg), n=n, n_total=n_total, seed=SEED,
                        dup=np.asarray(dup if dup is not None else np.zeros((0, 3)), dtype=np.int64),
                        **{f"codebook{l}": c for l, c in enumerate(cbs)}, semantic_ids=golden.astype(np.int16),
                        trace0=trace[0].astype(small), rounds=len(trace) - 1,
                        chg_items=np.concatenate(chg_items) if chg_items else np.zeros(0, np.int32),
                        chg_codes=np.concatenate(chg_codes) if chg_codes else np.zeros((0, Lv), small),
                        chg_offsets=np.cumsum([0] + [len(c) for c in chg_items]).astype(np.int64),
                        tie_gap=np.float64(TIE_GAP),
                        tie_items=np.concatenate(flagged).astype(np.int32) if flagged else np.zeros(0, np.int32),
                        tie_offsets=np.cumsum([0] + [len(f) for f in flagged]).astype(np.int64),
                        rows_differing_inside_rounds=bad_rows, final_rows_differing=nd)


def make_infer_c1():
    """BASELINE config 1: the reference's infer() verbatim on 707 items (main.py defaults)."""
    cfg = dict(in_dim=768, num_emb_list=[8, 8, 8], e_dim=32, layers=[256, 128], kmeans_iters=50,
               sk_epsilons=[0.01, 0.01, 0.01], sk_iters=50, quant_loss_weight=0.1)
    make_infer("c1_infer", cfg, 707, 707, (1, 65), (64, 707), exact_final=True)     # first training batch of 64 rows (main.py batch_size)


def make_infer_c2():
    """BASELINE config 2 shapes (3 x 256 codes, e 32) on a 60 000-item slice of the 1 M catalogue, with planted exact
    duplicates (pairs, triples and one 40-row block) so that groups of many sizes occur."""
    cfg = dict(in_dim=768, num_emb_list=[256] * 3, e_dim=32, layers=[256, 128], kmeans_iters=10,
               sk_epsilons=[0.0, 0.0, 0.003], sk_iters=50)
    dup = [(30000, 100, 60), (31000, 100, 20), (33000, 7, 1), (33001, 7, 1), (33002, 7, 1)]
    dup += [(34000 + i, 9, 1) for i in range(39)]
    make_infer("c2_infer", cfg, 60000, 1_000_000, (0, 8192), (64,), dup=dup)


def make_infer_c3():
    """BASELINE config 3 shapes (4 x 256 codes, e 64: the distance product takes the small-batch order for pairs)."""
    cfg = dict(in_dim=768, num_emb_list=[256] * 4, e_dim=64, layers=[256, 128], kmeans_iters=10,
               sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
    dup = [(10000, 100, 60), (11000, 100, 20)] + [(13000 + i, 9, 1) for i in range(20)]
    make_infer("c3_infer", cfg, 20000, 10_000_000, (0, 4096), (64,), dup=dup)


def make_small_batches():
    """Module-level calls on SMALL batches (2 … 20 rows): the reference's CPU GEMM switches its summation order below 16
    rows, so `get_indices` / `forward` of the same rows give other last bits than inside a large batch.  Pins the oracle's
    small-batch restatement and gives the GPU tests golden bits for RQVAE.get_indices / forward at those batch sizes."""
    out = {}
    names = []
    for tag, base in (("c2", "c2_slice"), ("c3", "c3_slice")):
        g = np.load(os.path.join(GOLD, f"{base}.npz"))
        cfg = json.loads(str(g["cfg"]))
        cbs = [g[f"codebook{l}"] for l in range(len(cfg["num_emb_list"]))]
        sd = synth.synth_state_dict(SEED, cfg["in_dim"], cfg["layers"], cfg["e_dim"], cfg["num_emb_list"])
        for l, c in enumerate(cbs):
            sd[f"rq.vq_layers.{l}.embedding.weight"] = c
        m = ref_model(cfg, sd)
        m.eval()
        nl = len(cfg["layers"]) + 1
        ew, eb = weights_of(sd, "encoder", nl)
        dw, db = weights_of(sd, "decoder", nl)
        x_all = synth.synth_items(SEED, 0, 4096, cfg["in_dim"], int(g["n_total"]))
        eps = [0.0] * (len(cbs) - 1) + [0.003]
        for M in (2, 3, 5, 6, 10, 11, 15, 16, 20):
            x = np.ascontiguousarray(x_all[100:100 + M])
            xt = torch.from_numpy(x)
            with torch.no_grad():
                z = m.encoder(xt).numpy()
                codes = m.get_indices(xt, use_sk=False).numpy()
                o, rq_loss, _ = m(xt, use_sk=False)
                for vq, e_ in zip(m.rq.vq_layers, eps):
                    vq.sk_epsilon = e_
                codes_sk = m.get_indices(xt, use_sk=True).numpy()
            zo = O.mlp_group(x, ew, eb)
            assert np.array_equal(zo.view(np.int32), z.view(np.int32)), (tag, M, "latent")
            pc, r = O.reencode_prefix(x, ew, eb, cbs)
            mine_sk = np.concatenate([pc, O.sinkhorn_last_level(r, cbs[-1], eps[-1], cfg["sk_iters"])[:, None]], 1)
            assert np.array_equal(mine_sk, codes_sk), (tag, M, "use_sk codes")
            assert np.array_equal(pc, codes[:, :-1]), (tag, M, "prefix codes")
            key = f"{tag}_m{M}"
            names.append(key)
            out[f"{key}_z"] = z
            out[f"{key}_codes"] = codes.astype(np.int16)
            out[f"{key}_codes_sk"] = codes_sk.astype(np.int16)
            out[f"{key}_out"] = o.numpy()[:, :16].copy()
            out[f"{key}_rq_loss"] = np.float64(rq_loss)
        print(f"small batches {tag}: oracle bit-equal to the reference for M in 2..20 (latent, prefix codes, use_sk codes)")
    out["names"] = np.array(names)
    out["first_row"] = np.int64(100)
    np.savez_compressed(os.path.join(GOLD, "small_batch.npz"), **out)


def make_sinkhorn_cases():
    from models.vq import VectorQuantizer
    rng = np.random.default_rng(SEED)
    cases = {}
    i = 0
    for (B, K, eps, iters) in [(2, 8, 0.01, 50), (3, 8, 0.01, 50), (5, 256, 0.01, 50), (17, 256, 0.01, 50),
                               (40, 256, 0.003, 100), (9, 1024, 0.01, 50), (64, 64, 0.05, 20)]:
        for rep in range(3):
            vq = VectorQuantizer(K, 32, sk_epsilon=eps, sk_iters=iters)
            cb = (rng.standard_normal((K, 32)) * 0.3).astype(np.float32)
            vq.embedding.weight.data.copy_(torch.from_numpy(cb))
            # a collision group: near-identical residuals
            base = (rng.standard_normal((1, 32)) * 0.3).astype(np.float32)
            r = (base + rng.standard_normal((B, 32)).astype(np.float32) * (1e-3 if rep else 0.2)).astype(np.float32)
            with torch.no_grad():
                _, _, ind = vq(torch.from_numpy(r), use_sk=True)
            d = O.quantize(r, [cb], want_xq=False, dist_level=0, threads=1)[3]
            mine = O.sinkhorn_assign(d, eps, iters)
            assert np.array_equal(mine, ind.numpy()), f"oracle sinkhorn mismatch on case {i}"
            cases[f"r{i}"] = r; cases[f"cb{i}"] = cb; cases[f"idx{i}"] = ind.numpy().astype(np.int16)
            cases[f"meta{i}"] = np.array([B, K, iters], dtype=np.int64); cases[f"eps{i}"] = np.float64(eps)
            i += 1
    cases["n_cases"] = np.int64(i)
    np.savez_compressed(os.path.join(GOLD, "sinkhorn_cases.npz"), **cases)
    print(f"sinkhorn: {i} cases, oracle == reference on all")


def make_csv_fixture():
    path = os.path.join(REF, "semantic_id_viz", "course_semantic_id_alignment.csv")
    rows = []
    with open(path, newline="", encoding="utf-8-sig") as f:
        for rec in csv.DictReader(f):
            rows.append([int(rec["code_1"]), int(rec["code_2"]), int(rec["code_3"]), int(rec["code_4"])])
    arr = np.array([[6, 1, 2, 0]] + rows, dtype=np.int16)      # padding row printed at RQVAE-T5/data_read.ipynb:13
    got = O.suffix_dedup(arr[:, :3].astype(np.int64))
    assert np.array_equal(got, arr.astype(np.int64)), "suffix rule does not reproduce the shipped artifact"
    np.save(os.path.join(GOLD, "course_semantic_ids.npy"), arr)
    print("csv artifact:", arr.shape, "groups", len(np.unique(arr[:, :3], axis=0)), "max suffix", arr[:, 3].max(),
          "— oracle suffix rule reproduces it")


def kblock_probe():
    """Re-derive the K-blocking of this host's CPU GEMM for the layer shapes in use (SURVEY.md §7 hard part 1)."""
    torch.manual_seed(0)
    for (K, N) in [(768, 256), (256, 128), (128, 32), (128, 64), (1024, 256), (32, 128), (64, 128), (128, 256), (256, 768),
                   (256, 1024)]:
        x = torch.randn(512, K); w = torch.randn(N, K) / K ** 0.5; b = torch.randn(N)
        ref = torch.nn.functional.linear(x, w, b).numpy()
        mine = O.linear(x.numpy(), w.numpy(), b.numpy(), relu=False)
        print(f"  linear {K:5d}->{N:4d} blocks {O.mkl_kblocks(K, N)}: bit-equal {np.array_equal(ref.view(np.int32), mine.view(np.int32))}")


def make_odd_slices():
    """Shapes outside the BASELINE configs (the reference accepts any): ragged codebook sizes with L = 5 (the maximum,
    infer.py:90) and e_dim 16 behind a narrow MLP; one level with e_dim 48; a five-Linear encoder."""
    make_slice("odd1_slice", dict(in_dim=104, num_emb_list=[100, 7, 300, 5, 64], e_dim=16, layers=[48, 24], kmeans_iters=5,
                                  sk_epsilons=[0.0] * 5, sk_iters=50), 2048, 100_000, 2048)
    make_slice("odd2_slice", dict(in_dim=768, num_emb_list=[256], e_dim=48, layers=[256, 128], kmeans_iters=5,
                                  sk_epsilons=[0.0], sk_iters=50), 2048, 100_000, 2048)
    make_slice("odd3_slice", dict(in_dim=512, num_emb_list=[64, 64], e_dim=16, layers=[512, 256, 128, 64], kmeans_iters=5,
                                  sk_epsilons=[0.0, 0.0], sk_iters=50), 2048, 100_000, 2048)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    print("torch", torch.__version__, "| sklearn", __import__("sklearn").__version__, "| numpy", np.__version__)
    if sys.argv[1:] == ["small"]:
        make_small_batches()
        sys.exit(0)
    if sys.argv[1:] == ["infer"]:        # only the infer() fixtures
        make_infer_c1()
        make_infer_c2()
        make_infer_c3()
        sys.exit(0)
    if sys.argv[1:] == ["odd"]:          # only the odd-shape slices (added later; the other fixtures stay as committed)
        make_odd_slices()
        sys.exit(0)
    kblock_probe()
    make_csv_fixture()
    make_sinkhorn_cases()
    make_slice("c2_slice", dict(in_dim=768, num_emb_list=[256] * 3, e_dim=32, layers=[256, 128], kmeans_iters=10,
                                sk_epsilons=[0.0, 0.0, 0.003], sk_iters=50), 8192, 1_000_000, 8192)
    make_slice("c3_slice", dict(in_dim=768, num_emb_list=[256] * 4, e_dim=64, layers=[256, 128], kmeans_iters=10,
                                sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50), 4096, 10_000_000, 4096)
    make_slice("c5_slice", dict(in_dim=1024, num_emb_list=[1024] * 4, e_dim=64, layers=[256, 128], kmeans_iters=5,
                                sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50), 2048, 100_000_000, 4096)
    make_slice("c1_slice", dict(in_dim=768, num_emb_list=[8, 8, 8], e_dim=32, layers=[256, 128], kmeans_iters=50,
                                sk_epsilons=[0.01, 0.01, 0.01], sk_iters=50), 707, 707, 64)
    make_infer_c1()
    make_infer_c2()
    make_infer_c3()
    make_small_batches()
    make_odd_slices()

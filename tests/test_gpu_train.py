"""Training step on the GPU against tests/golden/train_steps.npz — the reference's own loop
(train.py:108-121, AdamW + clip_grad_norm_ + linear warmup) run on CPU by oracle/make_golden_train.py.
Tolerance parity: 1e-4 relative (north_star); codes must be identical."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, train_case_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load():
    g = np.load(os.path.join(GOLD, "train_steps.npz"))
    return g, json.loads(str(g["cases"]))


def _model(cfg, sd):
    import ai_education_generative_recommendation_b200 as rq
    m = rq.RQVAE(in_dim=cfg["in_dim"], num_emb_list=cfg["num_emb_list"], e_dim=cfg["e_dim"], layers=cfg["layers"],
                 dropout_prob=0.0, bn=False, loss_type=cfg["loss_type"], quant_loss_weight=cfg["quant_loss_weight"],
                 beta=cfg["beta"], kmeans_init=False, kmeans_iters=10, sk_epsilons=list(cfg["sk_epsilons"]),
                 sk_iters=cfg["sk_iters"])
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
    return m.to(DEV).train()


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", ["sk_mse", "argmin_l1", "c1_shape"])
@pytest.mark.parametrize("fused", [True, False])
def test_training_steps_match_reference(name, fused):
    """fused=True: the B200 optimizer (FusedAdamW + warmup_lambda); fused=False: the reference's own loop objects
    (torch.optim.AdamW, clip_grad_norm_, transformers schedule) driving the product model unchanged."""
    import ai_education_generative_recommendation_b200 as rq
    g, cases = _load()
    cfg = cases[name]
    x_np, sd = train_case_state(cfg)
    m = _model(cfg, sd)
    x = torch.from_numpy(x_np).to(DEV)
    if fused:
        opt = rq.FusedAdamW(m.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"], max_norm=1.0)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, rq.warmup_lambda("linear", cfg["warmup_steps"], cfg["max_steps"]))
    else:
        from transformers import get_linear_schedule_with_warmup
        opt = torch.optim.AdamW(m.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
        sched = get_linear_schedule_with_warmup(opt, num_warmup_steps=cfg["warmup_steps"], num_training_steps=cfg["max_steps"])
    names = [n for n, _ in m.named_parameters()]
    assert names == [str(s) for s in g[f"{name}/names"]]
    for step in range(cfg["steps"]):
        opt.zero_grad()
        out, rq_loss, indices = m(x)
        loss, loss_recon = m.compute_loss(out, rq_loss, xs=x)
        loss.backward()
        assert np.array_equal(indices.cpu().numpy(), g[f"{name}/codes"][step].astype(np.int64)), f"codes differ at step {step}"
        for got, key in ((loss, "loss"), (loss_recon, "recon"), (rq_loss, "rq")):
            assert abs(got.item() - g[f"{name}/{key}"][step]) <= 1e-4 * abs(g[f"{name}/{key}"][step]), (key, step)
        if step == 0:
            for i, (n, p) in enumerate(m.named_parameters()):
                if f"{name}/grad0/{n}" in g:
                    assert _rel(p.grad.cpu().numpy(), g[f"{name}/grad0/{n}"]) <= 1e-4, n
                else:
                    ref = g[f"{name}/grad0_norms"][i]
                    assert abs(float(p.grad.double().norm()) - ref) <= 1e-4 * max(ref, 1e-12), n
        assert abs(sched.get_last_lr()[0] - g[f"{name}/lr"][step]) <= 1e-12
        if fused:
            opt.step()
            gn = float(opt.last_stats[0])
        else:
            gn = float(torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0))
            opt.step()
        sched.step()
        # until the first effective update (lr = 0 at step 0) everything is a function of the initial weights: 1e-4.
        # Afterwards weights agree to ~1e-8, but one ReLU unit switching for one sample moves a gradient norm by ~5e-4
        # (seen on c1_shape: fused and torch AdamW weights equal to 1.5e-8, one decoder unit flips) — hence 2e-3 there.
        tol = 1e-4 if step <= 1 else 2e-3
        assert abs(gn - g[f"{name}/gnorm"][step]) <= tol * g[f"{name}/gnorm"][step]
    for i, (n, p) in enumerate(m.named_parameters()):
        init = sd[n].astype(np.float64)
        got = p.detach().cpu().numpy().astype(np.float64)
        if f"{name}/final/{n}" in g:
            ref = g[f"{name}/final/{n}"].astype(np.float64)
            upd = np.linalg.norm(ref - init)
            assert np.linalg.norm(got - ref) <= 2e-3 * max(upd, 1e-12), (n, np.linalg.norm(got - ref), upd)
        else:
            ref = g[f"{name}/delta_norms"][i]
            assert abs(np.linalg.norm(got - init) - ref) <= 2e-3 * max(ref, 1e-12), n


def test_mlp_backward_matches_torch_autograd():
    """dW / db / dx of MLPFunction vs torch's fp32 autograd of the same MLP (plain PyTorch reference)."""
    from ai_education_generative_recommendation_b200.train_ops import MLPFunction
    torch.manual_seed(1)
    for n, dims in ((257, [768, 256, 128, 32]), (64, [32, 128, 256, 768]), (1000, [96, 40])):
        ps = []
        for a, b in zip(dims[:-1], dims[1:]):
            ps += [(torch.randn(b, a, device=DEV) / a ** 0.5).requires_grad_(), (0.1 * torch.randn(b, device=DEV)).requires_grad_()]
        x = torch.randn(n, dims[0], device=DEV, requires_grad=True)
        gy = torch.randn(n, dims[-1], device=DEV)
        y = MLPFunction.apply(x, 0.0, 0, None, *ps)
        got = torch.autograd.grad(y, [x] + ps, gy)
        h = x
        for i in range(len(ps) // 2):
            h = torch.nn.functional.linear(h, ps[2 * i], ps[2 * i + 1])
            if i < len(ps) // 2 - 1:
                h = torch.relu(h)
        ref = torch.autograd.grad(h, [x] + ps, gy)
        assert _rel(y.detach().cpu(), h.detach().cpu()) < 1e-5
        for a, b in zip(got, ref):
            assert _rel(a.cpu(), b.cpu()) < 1e-4


def test_dropout_mask_statistics_and_backward():
    from ai_education_generative_recommendation_b200.train_ops import MLPFunction, _dropout
    x = torch.ones(1 << 20, device=DEV)
    y = _dropout(x, 0.3, 12345)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.7) < 5e-3
    assert torch.allclose(y[y != 0], torch.full((1,), 1.0 / 0.7, device=DEV))
    assert torch.equal(y, _dropout(x, 0.3, 12345)) and not torch.equal(y, _dropout(x, 0.3, 12346))
    # gradient flows only through kept inputs, scaled like the forward
    W = torch.eye(64, device=DEV).requires_grad_()
    b = torch.zeros(64, device=DEV).requires_grad_()
    xin = torch.randn(512, 64, device=DEV, requires_grad=True)
    out = MLPFunction.apply(xin, 0.5, 77, None, W, b)
    out.sum().backward()
    mask = (out.detach() != 0)
    assert torch.allclose(xin.grad[mask], torch.full((1,), 2.0, device=DEV))
    assert float(xin.grad[~mask].abs().max()) == 0.0
    assert 0.45 < mask.float().mean().item() < 0.55


def test_fused_adamw_with_clipping_matches_torch():
    """Gradient norm far above 1 (clipping active), three steps, vs torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW."""
    import ai_education_generative_recommendation_b200 as rq
    torch.manual_seed(3)
    shapes = [(300, 70), (70,), (5, 33000), (1,)]
    a = [torch.randn(s, device=DEV).requires_grad_() for s in shapes]
    b = [t.detach().clone().requires_grad_() for t in a]
    fo = rq.FusedAdamW(a, lr=3e-3, weight_decay=0.05, max_norm=1.0)
    to = torch.optim.AdamW(b, lr=3e-3, weight_decay=0.05)
    for step in range(3):
        grads = [torch.randn(s, device=DEV) * (10.0 if step < 2 else 1e-3) for s in shapes]
        fo.zero_grad()
        for p, q, gr in zip(a, b, grads):
            p.grad.copy_(gr)
            q.grad = gr.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_(b, 1.0)
        to.step()
        fo.step()
        assert abs(float(fo.last_stats[0]) - float(ref_norm)) <= 1e-5 * float(ref_norm)
        for p, q in zip(a, b):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    sd = fo.state_dict()
    assert set(sd["state"][0].keys()) >= {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 3.0
    assert torch.allclose(sd["state"][0]["exp_avg"], to.state_dict()["state"][0]["exp_avg"], rtol=1e-5, atol=1e-7)


def test_trainer_fit_kmeans_init_and_checkpoint(tmp_path):
    """train.py end to end on the device: k-means codebook init on the first batch, two epochs, evaluation (collision
    rate from the dedup kernels), checkpoint in the reference's format that the encode driver loads."""
    import ai_education_generative_recommendation_b200 as rq
    from ai_education_generative_recommendation_b200 import synth
    n = 4096
    x = synth.synth_items(2024, 0, n, 768, 1_000_000)
    np.save(tmp_path / "embs.npy", x)
    params = dict(data_path=str(tmp_path / "embs.npy"), ckpt_dir=str(tmp_path / "ckpt"),
                  semantic_id_file=str(tmp_path / "codes.npy"), in_dim=768, num_emb_list=[32, 32, 32], e_dim=32,
                  layers=[256, 128], dropout=0.1, batch_normalize=False, loss_type="mse", quant_loss_weight=0.1, beta=0.25,
                  kmeans_init=True, kmeans_iters=10, lr=1e-3, epochs=2, warmup_epochs=1, batch_size=512, num_workers=0,
                  eval_step=1, sk_epsilons=[0.01, 0.01, 0.01], sk_iters=50, learner="AdamW", lr_scheduler_type="linear",
                  weight_decay=1e-4, save_limit=5, device=DEV)
    best_loss, best_cr = rq.train(params)
    assert np.isfinite(best_loss) and 0.0 <= best_cr < 1.0
    ck = torch.load(os.path.join(params["ckpt_dir"], "best_collision_model.pth"), map_location="cpu", weights_only=False)
    assert set(ck.keys()) == {"args", "epoch", "best_loss", "best_collision_rate", "state_dict", "optimizer"}
    assert float(ck["state_dict"]["rq.vq_layers.0.embedding.weight"].abs().sum()) > 0
    codes = rq.infer(params)                                   # the encode driver picks the checkpoint up
    assert codes.shape == (n, 4) and len(np.unique(codes, axis=0)) == n


def test_fused_and_torch_optimizers_train_alike_and_loss_decreases():
    """20 steps on C2-like shapes with k-means-initialised codebooks: FusedAdamW and torch's clip + AdamW follow the
    same loss trajectory, and the loss goes down."""
    import ai_education_generative_recommendation_b200 as rq
    from ai_education_generative_recommendation_b200 import synth
    x = torch.from_numpy(synth.synth_items(2024, 0, 8192, 768, 1_000_000)).to(DEV)
    torch.manual_seed(5)
    ma = rq.RQVAE(in_dim=768, num_emb_list=[64, 64, 64], e_dim=32, layers=[256, 128], dropout_prob=0.0, kmeans_init=True,
                  kmeans_iters=10, quant_loss_weight=0.1, sk_epsilons=[0.01, 0.01, 0.01], sk_iters=50).to(DEV).train()
    assert not ma.rq.vq_layers[0].initted
    out, ql, _ = ma(x[:1024])                           # first training batch: k-means init (vq.py:67-68)
    assert all(q.initted for q in ma.rq.vq_layers)
    mb = rq.RQVAE(in_dim=768, num_emb_list=[64, 64, 64], e_dim=32, layers=[256, 128], dropout_prob=0.0, kmeans_init=False,
                  quant_loss_weight=0.1, sk_epsilons=[0.01, 0.01, 0.01], sk_iters=50).to(DEV).train()
    mb.load_state_dict(ma.state_dict())
    fo = rq.FusedAdamW(ma.parameters(), lr=2e-4, weight_decay=1e-4)
    to = torch.optim.AdamW(mb.parameters(), lr=2e-4, weight_decay=1e-4)
    la, lb = [], []
    for step in range(20):
        xb = x[(step % 8) * 1024:(step % 8 + 1) * 1024]
        for m, o, acc in ((ma, fo, la), (mb, to, lb)):
            o.zero_grad()
            out, ql, _ = m(xb)
            loss, _ = m.compute_loss(out, ql, xs=xb)
            loss.backward()
            if o is to:
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            o.step()
            acc.append(loss.item())
    assert np.isfinite(la).all()
    assert np.allclose(la, lb, rtol=2e-3), (la, lb)
    # … and keeps going with the reference's schedule shape (30 warmup steps, then lr 1e-3).  What the unmodified reference
    # does on this data (run on CPU while writing this test): reconstruction loss 0.266 -> 0.242 after 110 steps while the
    # quantizer loss grows from 0.08 to 6 (encoder outruns the codebooks), i.e. the TOTAL loss rises; same picture here.
    sched = torch.optim.lr_scheduler.LambdaLR(fo, rq.warmup_lambda("constant", 30, 0))
    for g_ in fo.param_groups:
        g_["initial_lr"] = 1e-3
    sched.base_lrs = [1e-3]
    recon, quant = [], []
    for step in range(120):
        xb = x[(step % 8) * 1024:(step % 8 + 1) * 1024]
        fo.zero_grad()
        out, ql, _ = ma(xb)
        loss, lr_ = ma.compute_loss(out, ql, xs=xb)
        loss.backward()
        fo.step()
        sched.step()
        recon.append(lr_.item())
        quant.append(ql.item())
    assert np.isfinite(recon).all() and np.isfinite(quant).all()
    assert np.mean(recon[-10:]) < 0.96 * np.mean(recon[:10]), (recon[::10], quant[::10])


def test_training_forward_bits_equal_inference_forward():
    """Same Linear kernels and the same quantizer arithmetic: with dropout off and Sinkhorn off the differentiable
    forward returns the bits of the inference forward (which the golden vectors pin to the reference)."""
    import ai_education_generative_recommendation_b200 as rq
    from conftest import build_model, load_golden
    from ai_education_generative_recommendation_b200 import synth
    g, cfg, cbs = load_golden("c2_slice")
    m = build_model(cfg, cbs)
    x = torch.from_numpy(synth.synth_items(2024, 0, 3000, cfg["in_dim"], 1_000_000)).to(DEV)
    out_i, loss_i, idx_i = m(x, use_sk=False)
    m.train()
    out_t, loss_t, idx_t = m(x, use_sk=False)
    assert out_t.requires_grad and not out_i.requires_grad
    assert torch.equal(idx_t, idx_i) and torch.equal(out_t.detach(), out_i)
    assert abs(float(loss_t) - float(loss_i)) <= 1e-6 * abs(float(loss_i))


@pytest.mark.parametrize("B,K", [(2048, 256), (5000, 256), (1024, 64), (1111, 1024), (300, 64), (17, 8)])
def test_batch_sinkhorn_on_the_whole_gpu_matches_oracle(oracle, B, K):
    """vq.py:74-83 on a training-size batch: the cooperative multi-CTA kernel (rows split over the SMs; used from 1024
    rows up, smaller batches keep the one-CTA kernel) picks the same code as the oracle's restatement of
    layers.py:85-108 for every row of these batches."""
    import ai_education_generative_recommendation_b200 as rq
    rng = np.random.default_rng(B + K)
    e = 32
    centres = rng.standard_normal((K, e)).astype(np.float32) * 0.3
    z = (centres[rng.integers(0, K, size=B)] + 0.25 * rng.standard_normal((B, e))).astype(np.float32)
    _, _, _, d = oracle.quantize(z, [centres], want_xq=False, dist_level=0, threads=1)
    ref = oracle.sinkhorn_assign(d, 0.01, 50)
    dt = torch.from_numpy(d).to(DEV)
    scratch = torch.empty((B, K), dtype=torch.float64, device=DEV)
    idx = torch.empty((B,), dtype=torch.int64, device=DEV)
    from ai_education_generative_recommendation_b200._cabi import check, lib, ptr, stream_ptr
    check(lib().rqb200_sinkhorn_assign(ptr(dt), B, K, 0.01, 50, ptr(scratch), ptr(idx), stream_ptr(dt.device)))
    got = idx.cpu().numpy()
    assert np.array_equal(got, ref), f"{int((got != ref).sum())} of {B} rows differ"
    # balanced assignment: no code is starved or flooded beyond what the arg-max of a doubly-normalised Q allows
    counts = np.bincount(got, minlength=K)
    assert counts.max() <= max(4 * B // K, 8)


def test_graph_replayed_steps_equal_eager_steps(tmp_path):
    """The Trainer captures the step as a CUDA graph after three eager steps; with dropout off the replayed steps give the
    losses and weights of the eager loop (same kernels, per-step scalars read from device memory)."""
    import ai_education_generative_recommendation_b200 as rq
    from ai_education_generative_recommendation_b200 import synth
    x = torch.from_numpy(synth.synth_items(2024, 0, 12 * 1024, 768, 1_000_000)).to(DEV)
    results = []
    for use_graph in (True, False):
        torch.manual_seed(11)
        m = rq.RQVAE(in_dim=768, num_emb_list=[64, 64, 64], e_dim=32, layers=[256, 128], dropout_prob=0.0, kmeans_init=True,
                     kmeans_iters=5, quant_loss_weight=0.1, sk_epsilons=[0.01, 0.01, 0.0], sk_iters=50)
        params = dict(lr=5e-4, learner="AdamW", lr_scheduler_type="linear", weight_decay=1e-4, epochs=2, warmup_epochs=1,
                      save_limit=1, eval_step=5, device=DEV, ckpt_dir=str(tmp_path / f"ck{int(use_graph)}"), cuda_graph=use_graph)
        loader = rq.DeviceBatches(x, 1024, shuffle=False)
        tr = rq.Trainer(params, m, len(loader))
        l0 = tr._train_epoch(loader, 0)
        l1 = tr._train_epoch(loader, 1)
        results.append((l0, l1, [p.detach().clone() for p in m.parameters()], tr.graph_replays, tr.optimizer._steps))
    (a0, a1, pa, ra, sa), (b0, b1, pb, rb, sb) = results
    assert ra >= 20 and rb == 0 and sa == sb == 24
    assert np.allclose(a0, b0, rtol=1e-4) and np.allclose(a1, b1, rtol=1e-3), (a0, b0, a1, b1)
    for p, q in zip(pa, pb):
        assert float((p - q).norm()) <= 2e-2 * float((q - q.mean()).norm() + 1e-12)


def test_graph_replay_changes_the_dropout_mask_every_step():
    import ai_education_generative_recommendation_b200 as rq
    from ai_education_generative_recommendation_b200.train_ops import _dropout
    x = torch.ones(1 << 16, device=DEV)
    seed_dev = torch.zeros((1,), dtype=torch.int64, device=DEV)
    y0 = _dropout(x, 0.5, 99, seed_dev)
    g = torch.cuda.CUDAGraph()
    static = torch.empty_like(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        seed_dev.add_(1)
        static.copy_(_dropout(x, 0.5, 99, seed_dev))
    g.replay()
    y1 = static.clone()
    g.replay()
    y2 = static.clone()
    assert not torch.equal(y1, y2) and not torch.equal(y0, y1)
    assert abs((y1 != 0).float().mean().item() - 0.5) < 0.02

"""Differentiable pieces of the RQ-VAE training step (reference RQ-VAE/train.py:97-124), as torch.autograd
Functions whose forward AND backward run in librqvae_b200.so (csrc/train.cu) — torch only carries the graph.

  MLPFunction    layers.py:18-43   [Dropout, Linear, ReLU]* Linear, backward = dW / db / dx GEMMs
  RQFunction     rq.py:39-56 + vq.py:63-99 with use_sk (Sinkhorn arg-max, layers.py:85-108): straight-through
                 gradient to the latent, commitment gradient of level 0, codebook-loss gradient per level
  ReconFunction  rqvae.py:73-84    mse / l1 reconstruction loss
  FusedAdamW     train.py:75-78,116-117   clip_grad_norm_(…, 1.0) + AdamW.step() in three launches

Gradient algebra of the quantizer (why only level 0 feeds the latent): x_res_l = r_l + (q_l − r_l).detach() has
identity Jacobian w.r.t. r_l, and r_{l+1} = r_l − x_res_l, so d r_{l+1} / d r_l = I − I = 0: every path from a deeper
level back to z cancels exactly, leaving d x_q / d z = I (through level 0) and the commitment term β·mse(q_0, z).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr

_MASK64 = (1 << 64) - 1


def _mix(seed: int, a: int) -> int:
    x = (seed ^ (a * 0xD6E8FEB86659FD93)) & _MASK64
    x = (x + 0x9E3779B97F4A7C15) & _MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK64
    return x ^ (x >> 31)


def _dropout(x: torch.Tensor, p: float, seed: int, seed_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """seed_dev: optional int64[1] device tensor mixed into the seed on the device (CUDA-graph replays, see Trainer)."""
    y = torch.empty_like(x)
    check(_cabi.lib().rqb200_dropout_dev(ptr(x), x.numel(), float(p), ctypes.c_uint64(seed), ptr(seed_dev), ptr(y),
                                         stream_ptr(x.device)))
    return y


_CONST_CACHE = {}


def _const_tensor(values, dtype, device) -> torch.Tensor:
    """Small constant device tensors (betas …) are created once: a host→device copy is not allowed while a CUDA graph
    is being captured."""
    key = (tuple(float(v) for v in values), dtype, str(device))
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.tensor(list(values), dtype=dtype, device=device)
        _CONST_CACHE[key] = t
    return t


class MLPFunction(torch.autograd.Function):
    """y = MLP(x) for parameters (W0, b0, W1, b1, …): ReLU after every layer but the last, Dropout(p) in front of
    every Linear when p > 0 (masks are regenerated from `seed` in backward, never stored)."""

    @staticmethod
    def forward(ctx, x, p_drop, seed, seed_dev, *params):
        lib = _cabi.lib()
        n_layers = len(params) // 2
        h = x.contiguous()
        n = h.shape[0]
        s = stream_ptr(h.device)
        saved = []
        for i in range(n_layers):
            W, b = params[2 * i].contiguous(), params[2 * i + 1].contiguous()
            hd = _dropout(h, p_drop, _mix(seed, i), seed_dev) if p_drop > 0 else h
            y = torch.empty((n, W.shape[0]), dtype=torch.float32, device=h.device)
            check(lib.rqb200_linear_forward(ptr(hd), ptr(W), ptr(b), n, W.shape[1], W.shape[0],
                                            1 if i < n_layers - 1 else 0, ptr(y), s))
            saved += [hd, y]
            h = y
        ctx.p_drop, ctx.seed, ctx.n_layers, ctx.seed_dev = p_drop, seed, n_layers, seed_dev
        ctx.save_for_backward(*saved, *params)
        return h

    @staticmethod
    def backward(ctx, gy):
        lib = _cabi.lib()
        nl = ctx.n_layers
        tensors = ctx.saved_tensors
        acts, params = tensors[:2 * nl], tensors[2 * nl:]
        dy = gy.contiguous().clone()                 # masked in place by the ReLU layers
        n = dy.shape[0]
        s = stream_ptr(dy.device)
        grads: List[Optional[torch.Tensor]] = [None] * (2 * nl)
        dx = None
        for i in range(nl - 1, -1, -1):
            W = params[2 * i].contiguous()
            hd, y = acts[2 * i], acts[2 * i + 1]
            out_dim, in_dim = W.shape
            need_dx = i > 0 or ctx.needs_input_grad[0]
            dx = torch.empty((n, in_dim), dtype=torch.float32, device=dy.device) if need_dx else None
            dW = torch.empty_like(W)
            db = torch.empty((out_dim,), dtype=torch.float32, device=dy.device)
            nscr = int(lib.rqb200_linear_backward_scratch_floats(n, in_dim, out_dim))
            nscr = min(nscr, 64 * in_dim * out_dim)
            scratch = torch.empty((nscr,), dtype=torch.float32, device=dy.device)
            check(lib.rqb200_linear_backward(ptr(hd), ptr(W), ptr(y), ptr(dy), n, in_dim, out_dim,
                                             1 if i < nl - 1 else 0, ptr(dx), ptr(dW), ptr(db), ptr(scratch), nscr, s))
            grads[2 * i], grads[2 * i + 1] = dW, db
            if need_dx and ctx.p_drop > 0:
                dx = _dropout(dx, ctx.p_drop, _mix(ctx.seed, i), ctx.seed_dev)
            dy = dx
        return (dx if ctx.needs_input_grad[0] else None, None, None, None, *grads)


class RQFunction(torch.autograd.Function):
    """(x_q, mean loss, indices) = ResidualVectorQuantizer(z) with per-level Sinkhorn when eps > 0."""

    @staticmethod
    def forward(ctx, z, betas, eps_list, sk_iters, *codebooks):
        lib = _cabi.lib()
        z = z.contiguous()
        n, e = z.shape
        dev = z.device
        s = stream_ptr(dev)
        Lv = len(codebooks)
        xq = torch.empty_like(z)
        sumsq = torch.zeros((Lv,), dtype=torch.float64, device=dev)
        idx_all = torch.empty((Lv, n), dtype=torch.int64, device=dev)
        residuals = []
        r = z
        for l, cb in enumerate(codebooks):
            cb = cb.contiguous()
            K = cb.shape[0]
            cnorm = torch.empty((K,), dtype=torch.float32, device=dev)
            idx = idx_all[l]
            if eps_list[l] > 0:
                d = torch.empty((n, K), dtype=torch.float32, device=dev)
                check(lib.rqb200_kmeans_distances(ptr(r), n, e, ptr(cb), K, ptr(cnorm), ptr(d), s))
                scratch = torch.empty((n, K), dtype=torch.float64, device=dev)
                check(lib.rqb200_sinkhorn_assign(ptr(d), n, K, float(eps_list[l]), int(sk_iters), ptr(scratch), ptr(idx), s))
            else:
                check(lib.rqb200_kmeans_assign(ptr(r), n, e, ptr(cb), K, ptr(cnorm), ptr(idx), s))
            r_next = torch.empty_like(r)
            check(lib.rqb200_rq_level_apply(ptr(r), ptr(idx), ptr(cb), n, e, 1 if l == 0 else 0, ptr(xq), ptr(r_next),
                                            sumsq[l:].data_ptr(), s))
            residuals.append(r)
            r = r_next
        mse = sumsq / float(max(n * e, 1))
        bt = _const_tensor(betas, torch.float64, dev)
        loss = (mse + bt * mse).mean().to(torch.float32)
        ctx.betas, ctx.Lv = list(betas), Lv
        ctx.save_for_backward(idx_all, *residuals, *codebooks)
        indices = idx_all.t().contiguous()
        ctx.mark_non_differentiable(indices)
        return xq, loss, indices

    @staticmethod
    def backward(ctx, g_xq, g_loss, _g_idx):
        lib = _cabi.lib()
        Lv = ctx.Lv
        t = ctx.saved_tensors
        idx_all, residuals, codebooks = t[0], t[1:1 + Lv], t[1 + Lv:]
        z = residuals[0]
        n, e = z.shape
        dev = z.device
        s = stream_ptr(dev)
        g = (g_loss if g_loss is not None else torch.zeros((), device=dev)).reshape(1).to(torch.float32).contiguous()
        base = 2.0 / float(max(n * e, 1)) / float(Lv)
        dz = None
        if ctx.needs_input_grad[0]:
            dz = torch.empty_like(z)
            gx = g_xq.contiguous() if g_xq is not None else None
            check(lib.rqb200_rq_latent_grad(ptr(z), ptr(idx_all[0]), ptr(codebooks[0].contiguous()), ptr(gx), n, e,
                                            base * ctx.betas[0], ptr(g), ptr(dz), s))
        dEs = []
        for l in range(Lv):
            cb = codebooks[l].contiguous()
            if not ctx.needs_input_grad[4 + l]:
                dEs.append(None)
                continue
            dE = torch.empty_like(cb)
            check(lib.rqb200_vq_codebook_grad(ptr(residuals[l]), ptr(idx_all[l]), ptr(cb), n, e, cb.shape[0], base,
                                              ptr(g), ptr(dE), s))
            dEs.append(dE)
        return (dz, None, None, None, *dEs)


class ReconFunction(torch.autograd.Function):
    """mean((out − xs)²) or mean(|out − xs|) (rqvae.py:75-78)."""

    @staticmethod
    def forward(ctx, out, xs, l1):
        out_c, xs_c = out.contiguous(), xs.contiguous()
        sums = torch.zeros((2,), dtype=torch.float64, device=out.device)
        check(_cabi.lib().rqb200_recon_loss(ptr(out_c), ptr(xs_c), out_c.numel(), ptr(sums), stream_ptr(out.device)))
        ctx.l1 = bool(l1)
        ctx.save_for_backward(out_c, xs_c)
        return (sums[1 if l1 else 0] / float(max(out_c.numel(), 1))).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        out_c, xs_c = ctx.saved_tensors
        d = torch.empty_like(out_c)
        gg = g.reshape(1).to(torch.float32).contiguous()
        check(_cabi.lib().rqb200_recon_grad(ptr(out_c), ptr(xs_c), out_c.numel(), 1 if ctx.l1 else 0, ptr(gg), ptr(d),
                                            stream_ptr(out_c.device)))
        return d, None, None


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW(params, lr, weight_decay) preceded by clip_grad_norm_(params, max_norm) — what
    train.py:116-117 does every step — as ONE C-ABI call (three launches, no host synchronisation).

    Gradients live in one flat buffer (every `p.grad` is a view into it), which is also what a data-parallel run
    all-reduces (`flat_grad`); `grad_scale` folds the 1/world averaging into the step.  `state_dict()` has the layout of
    torch.optim.AdamW's, so checkpoints keep the reference's format (train.py:157-165)."""

    CHUNK = 1 << 14

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.max_norm = float(max_norm)
        self.grad_scale = 1.0
        self._flat = None
        self._table = None
        self.last_stats = None          # device tensor [total grad norm, clip coefficient] of the last step

    def _params(self):
        return [p for g in self.param_groups for p in g["params"] if p.requires_grad]

    def _build(self):
        ps = self._params()
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW needs CUDA parameters (there is no CPU fallback)")
        total = sum(p.numel() for p in ps)
        self._flat = torch.zeros((total,), dtype=torch.float32, device=dev)
        self._m = torch.zeros_like(self._flat)
        self._v = torch.zeros_like(self._flat)
        rows, off = [], 0
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdamW: float32 contiguous parameters only")
            k = p.numel()
            view = self._flat[off:off + k].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            st = self.state[p]
            st["step"] = st.get("step", torch.tensor(0.0))
            for name, buf in (("exp_avg", self._m), ("exp_avg_sq", self._v)):
                bview = buf[off:off + k].view_as(p)
                if name in st:
                    bview.copy_(st[name])
                st[name] = bview
            for c0 in range(0, k, self.CHUNK):
                c = min(self.CHUNK, k - c0)
                rows.append([p.data_ptr() + 4 * c0, self._flat.data_ptr() + 4 * (off + c0),
                             self._m.data_ptr() + 4 * (off + c0), self._v.data_ptr() + 4 * (off + c0), c])
            off += k
        self._table = torch.tensor(rows, dtype=torch.int64, device=dev)
        self._partial = torch.empty((len(rows),), dtype=torch.float64, device=dev)
        self._stats = torch.zeros((2,), dtype=torch.float32, device=dev)
        self._sig = [(p.data_ptr(), p.numel()) for p in ps]
        self._steps = int(max([float(self.state[p]["step"]) for p in ps] + [0.0]))

    # ---- CUDA-graph replay support: per-step scalars in device memory --------------------------------
    def refresh_hyper(self):
        """Host side of one graph-replayed step: advances the step count and refreshes the six scalars the captured
        `step_from_device` launch reads (decay, lr / bc1, sqrt(bc2), beta1, beta2, eps).  Call right before replay."""
        if self._flat is None:
            self._build()
        g = self.param_groups[0]
        self._steps += 1
        b1, b2 = float(g["betas"][0]), float(g["betas"][1])
        lr, wd = float(g["lr"]), float(g["weight_decay"])
        bc1, bc2 = 1.0 - b1 ** self._steps, 1.0 - b2 ** self._steps
        if getattr(self, "_hyper", None) is None:
            self._hyper = torch.zeros((6,), dtype=torch.float32, device=self._flat.device)
        vals = (ctypes.c_float * 6)(1.0 - lr * wd, lr / bc1, bc2 ** 0.5, b1, b2, float(g["eps"]))
        check(_cabi.lib().rqb200_set_floats(ptr(self._hyper), 6, vals, stream_ptr(self._flat.device)))

    def step_from_device(self):
        """The launch a CUDA graph captures: clip + AdamW with the scalars of `refresh_hyper` (nothing else — no host
        bookkeeping, no pointer checks)."""
        if getattr(self, "_hyper", None) is None:
            raise RuntimeError("call refresh_hyper() once before capturing step_from_device()")
        check(_cabi.lib().rqb200_adamw_clip_step_dev(ptr(self._table), self._table.shape[0], ptr(self._partial), ptr(self._stats),
                                                     float(self.grad_scale), self.max_norm, ptr(self._hyper),
                                                     stream_ptr(self._flat.device)))
        self.last_stats = self._stats

    def after_replay(self):
        """Host bookkeeping after a replayed step (parameter version counters)."""
        _bump_versions(self._params())

    def state_dict(self):
        for p in self._params():
            if p in self.state:
                self.state[p]["step"] = torch.tensor(float(getattr(self, "_steps", 0)))
        return super().state_dict()

    @property
    def flat_grad(self) -> torch.Tensor:
        if self._flat is None:
            self._build()
        return self._flat

    def zero_grad(self, set_to_none: bool = False):
        """Keeps the flat gradient views alive (set_to_none is ignored on purpose)."""
        if self._flat is None:
            self._build()
        self._flat.zero_()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat = None                     # rebuilt (and exp_avg / exp_avg_sq re-adopted) at the next step

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closures are not part of the reference's loop")
        if self._flat is None:
            self._build()
        ps = self._params()
        if [(p.data_ptr(), p.numel()) for p in ps] != self._sig:
            self._flat = None
            self._build()
            ps = self._params()
        off = 0
        for p in ps:                          # a gradient that was re-created outside the flat buffer is copied in
            k = p.numel()
            if p.grad is None:
                self._flat[off:off + k].zero_()
            elif p.grad.data_ptr() != self._flat.data_ptr() + 4 * off:
                view = self._flat[off:off + k].view_as(p)
                view.copy_(p.grad)
                p.grad = view
            off += k
        group = self.param_groups[0]
        for g in self.param_groups[1:]:
            if (g["lr"], g["betas"], g["eps"], g["weight_decay"]) != (group["lr"], group["betas"], group["eps"],
                                                                        group["weight_decay"]):
                raise NotImplementedError("FusedAdamW: one hyper-parameter set for all parameters (as in train.py:75-78)")
        self._steps += 1
        dev = self._flat.device
        check(_cabi.lib().rqb200_adamw_clip_step(ptr(self._table), self._table.shape[0], ptr(self._partial), ptr(self._stats),
                                                 float(self.grad_scale), self.max_norm, float(group["lr"]),
                                                 float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                                 float(group["weight_decay"]), self._steps, stream_ptr(dev)))
        for p in ps:
            self.state[p]["step"] = torch.tensor(float(self._steps))
        self.last_stats = self._stats
        _bump_versions(ps)
        return None


def _bump_versions(params):
    """The kernel wrote the parameters behind torch's back: bump their version counters (host-side only) so that
    anything keyed on `_version` (RQVAE._sync's upload cache, autograd's saved-tensor checks) sees the change."""
    for p in params:
        torch.autograd.graph.increment_version(p)

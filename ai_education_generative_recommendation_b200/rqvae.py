"""Host-side mirror of the reference's RQ-VAE module tree, backed by the sm_100a C-ABI library.

Same class names, constructor arguments, sub-module paths and state_dict keys as the reference
(RQ-VAE/models/rqvae.py:9-84, rq.py:7-56, vq.py:7-99, layers.py:7-108), so a reference checkpoint
loads with ``load_state_dict`` and the reference's driver code (``model.get_indices``,
``model.rq.vq_layers[i].sk_epsilon = 0.0`` …) runs unchanged.  The arithmetic itself never runs in
PyTorch: every forward / get_indices call goes through ``librqvae_b200.so``; tensors must live on a
CUDA device (RuntimeError otherwise — there is no CPU fallback).

With gradients enabled, ``forward`` / ``compute_loss`` build an autograd graph out of ``train_ops`` Functions
(forward and backward both in the CUDA library), so the reference's training loop (train.py:97-124) runs
unchanged on this class; ``trainer.py`` is the B200-side Trainer (fused clip + AdamW, data-parallel all-reduce).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _cabi, torch_ops
from ._cabi import check, ptr, stream_ptr


def _require_cuda_tensor(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: the B200 RQ-VAE path has no CPU fallback "
                           f"(got device {t.device})")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what} must be float32 (got {t.dtype})")


# --------------------------------------------------------------------------------------- layers

def activation_layer(activation_name="relu", emb_dim=None):
    """Same contract as reference layers.py:45-67 (only ReLU is ever used on this path)."""
    if activation_name is None:
        return None
    if isinstance(activation_name, str):
        table = {"sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "relu": nn.ReLU, "leakyrelu": nn.LeakyReLU}
        key = activation_name.lower()
        if key == "none":
            return None
        if key in table:
            return table[key]()
        return None
    if isinstance(activation_name, type) and issubclass(activation_name, nn.Module):
        return activation_name()
    raise NotImplementedError("activation function {} is not implemented".format(activation_name))


class MLPLayers(nn.Module):
    """Parameter container with the reference's ``mlp_layers`` Sequential layout (layers.py:18-32):
    [Dropout, Linear, (BatchNorm1d), (ReLU)] per layer, no BN / activation after the last Linear.
    ``forward`` runs the exact-order CUDA kernels (eval semantics: Dropout is the identity)."""

    def __init__(self, layers, dropout=0.0, activation="relu", bn=False):
        super().__init__()
        self.layers = layers
        self.dropout = dropout
        self.activation = activation
        self.use_bn = bn
        mods: List[nn.Module] = []
        n = len(layers) - 1
        for i in range(n):
            mods.append(nn.Dropout(p=dropout))
            mods.append(nn.Linear(layers[i], layers[i + 1]))
            last = i == n - 1
            if bn and not last:
                mods.append(nn.BatchNorm1d(num_features=layers[i + 1]))
            act = activation_layer(activation, layers[i + 1])
            if act is not None and not last:
                mods.append(act)
        self.mlp_layers = nn.Sequential(*mods)
        for m in self.mlp_layers:
            if isinstance(m, nn.Linear):                  # layers.py:35-40
                nn.init.xavier_normal_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.fill_(0.0)
        if isinstance(activation, str) and activation.lower() != "relu":
            raise NotImplementedError("only the ReLU MLP of the reference's RQVAE is implemented in CUDA")
        self._owner = None      # (RQVAE, which) set by RQVAE so forward can reach the C handle

    def linears(self):
        """[(Linear, BatchNorm1d | None)] in order."""
        out = []
        mods = list(self.mlp_layers)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) else None
                out.append((m, bn))
        return out

    def forward(self, input_feature):
        if self._owner is None:
            raise RuntimeError("MLPLayers must be owned by an RQVAE to run (standalone use is not part of the path)")
        model, which = self._owner
        return model._mlp(which, input_feature)


# --------------------------------------------------------------------------------------- k-means / sinkhorn

def kmeans(samples, num_clusters, num_iters=10, seed: Optional[int] = None, init: Optional[torch.Tensor] = None,
           group=None, tol: float = 1e-4):
    """Drop-in for ``kmeans(samples, num_clusters, num_iters)`` (layers.py:69-82): k-means++ seeding then
    Lloyd iterations, all on the GPU (the reference copies to the host and calls scikit-learn).
    ``group``: optional torch.distributed process group — samples are then one shard per rank and the
    per-cluster sums / counts are all-reduced every iteration (NCCL over NVLink).
    Parity with scikit-learn is statistical only (its seeding consumes numpy's global RNG)."""
    from .kmeans_gpu import kmeans_fit
    return kmeans_fit(samples, num_clusters, num_iters, seed=seed, init=init, group=group, tol=tol)


@torch.no_grad()
def sinkhorn_algorithm(distances, epsilon, sinkhorn_iterations):
    """Drop-in for layers.py:85-108; ``distances`` [B,K] (any float dtype, CUDA) → Q [B,K] float64."""
    if not distances.is_cuda:
        raise RuntimeError("sinkhorn_algorithm: CUDA tensor required (no CPU fallback)")
    Q = distances.detach().to(torch.float64).contiguous().clone()
    B, K = Q.shape
    check(_cabi.lib().rqb200_sinkhorn(ptr(Q), B, K, float(epsilon), int(sinkhorn_iterations),
                                      stream_ptr(Q.device)))
    return Q


# --------------------------------------------------------------------------------------- quantizers

class VectorQuantizer(nn.Module):
    """Same attributes as reference vq.py:9-30; ``sk_epsilon`` stays a plain writable attribute
    (the encode driver sets it to 0.0 on all but the last level, infer.py:109-110)."""

    def __init__(self, n_e, e_dim, beta=0.25, kmeans_init=False, kmeans_iters=10, sk_epsilon=0.003, sk_iters=100):
        super().__init__()
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.sk_epsilon = sk_epsilon
        self.sk_iters = sk_iters
        self.embedding = nn.Embedding(n_e, e_dim)
        if not kmeans_init:
            self.initted = True
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)
        else:
            self.initted = False
            self.embedding.weight.data.zero_()
        self._owner = None      # (RQVAE, level)

    def get_codebook(self):
        return self.embedding.weight

    def get_codebook_entry(self, indices, shape=None):
        z_q = self.embedding.weight.detach()[indices]
        if shape is not None:
            z_q = z_q.view(shape)
        return z_q

    def init_emb(self, data):
        # `_kmeans_group` (set by the data-parallel Trainer): every rank contributes its rows of the first batch and all
        # ranks end up with the same centres
        centers = kmeans(data, self.n_e, self.kmeans_iters, group=getattr(self, "_kmeans_group", None))
        self.embedding.weight.data.copy_(centers)
        self.initted = True

    @staticmethod
    def center_distance_for_constraint(distances):
        """vq.py:51-61, elementwise fp32 (only used by the generic per-level path)."""
        max_distance = distances.max()
        min_distance = distances.min()
        middle = (max_distance + min_distance) / 2
        amplitude = max_distance - middle + 1e-5
        assert amplitude > 0
        return (distances - middle) / amplitude

    @torch.no_grad()
    def forward(self, x, use_sk=True):
        """One level (vq.py:63-99): returns (x + (x_q - x), loss, indices)."""
        if self._owner is None:
            raise RuntimeError("VectorQuantizer must be owned by an RQVAE to run")
        model, level = self._owner
        return model._vq_level(level, x, use_sk)


class ResidualVectorQuantizer(nn.Module):
    def __init__(self, n_e_list, e_dim, sk_epsilons, beta=0.25, kmeans_init=False, kmeans_iters=100, sk_iters=100):
        super().__init__()
        self.n_e_list = n_e_list
        self.e_dim = e_dim
        self.num_quantizers = len(n_e_list)
        self.beta = beta
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.sk_epsilons = sk_epsilons
        self.sk_iters = sk_iters
        self.vq_layers = nn.ModuleList([
            VectorQuantizer(n_e, e_dim, beta=beta, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                            sk_epsilon=eps, sk_iters=sk_iters)
            for n_e, eps in zip(n_e_list, sk_epsilons)])
        self._owner = None

    def get_codebook(self):
        return torch.stack([q.get_codebook() for q in self.vq_layers])

    @torch.no_grad()
    def forward(self, x, use_sk=True):
        """rq.py:39-56: returns (x_q, mean_losses, all_indices[..., L])."""
        if self._owner is None:
            raise RuntimeError("ResidualVectorQuantizer must be owned by an RQVAE to run")
        return self._owner._rq(x, use_sk)


# --------------------------------------------------------------------------------------- RQVAE

class RQVAE(nn.Module):
    """Drop-in for reference rqvae.py:9-84 (same constructor signature and defaults)."""

    def __init__(self, in_dim=768, num_emb_list=None, e_dim=64, layers=None, dropout_prob=0.0, bn=False,
                 loss_type="mse", quant_loss_weight=1.0, beta=0.25, kmeans_init=False, kmeans_iters=100,
                 sk_epsilons=None, sk_iters=100):
        super().__init__()
        self.in_dim = in_dim
        self.num_emb_list = num_emb_list
        self.e_dim = e_dim
        self.layers = layers
        self.dropout_prob = dropout_prob
        self.bn = bn
        self.loss_type = loss_type
        self.quant_loss_weight = quant_loss_weight
        self.beta = beta
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.sk_epsilons = sk_epsilons
        self.sk_iters = sk_iters

        self.encode_layer_dims = [in_dim] + list(layers) + [e_dim]
        self.encoder = MLPLayers(layers=self.encode_layer_dims, dropout=dropout_prob, bn=bn)
        self.rq = ResidualVectorQuantizer(num_emb_list, e_dim, beta=beta, kmeans_init=kmeans_init,
                                          kmeans_iters=kmeans_iters, sk_epsilons=sk_epsilons, sk_iters=sk_iters)
        self.decode_layer_dims = self.encode_layer_dims[::-1]
        self.decoder = MLPLayers(layers=self.decode_layer_dims, dropout=dropout_prob, bn=bn)

        # non-module back references (object.__setattr__ keeps them out of the module tree)
        object.__setattr__(self.encoder, "_owner", (self, 0))
        object.__setattr__(self.decoder, "_owner", (self, 1))
        object.__setattr__(self.rq, "_owner", self)
        for lvl, q in enumerate(self.rq.vq_layers):
            object.__setattr__(q, "_owner", (self, lvl))

        self._handle = ctypes.c_void_p(None)
        self._handle_device = None
        self._synced = {}
        self.encode_mode = _cabi.ENCODE_EXACT     # or _cabi.ENCODE_FAST (tensor-core path, same codes)
        self.kblocks = {}                          # optional {("enc"|"dec", layer): [k-block sizes]}
        self._stats_fast = False

    # ---- C handle management -------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                _cabi.lib().rqb200_model_destroy(self._handle)
                self._handle = ctypes.c_void_p(None)
        except Exception:
            pass

    def _device(self) -> torch.device:
        return self.rq.vq_layers[0].embedding.weight.device

    def _ensure_handle(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("RQVAE parameters must be on a CUDA device (model.to('cuda:0')): the B200 path "
                               "has no CPU fallback")
        _cabi.require_cuda()
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._handle.value and self._handle_device == index:
            return
        if self._handle.value:
            _cabi.lib().rqb200_model_destroy(self._handle)
            self._handle = ctypes.c_void_p(None)
        dims = _cabi.int_array(self.encode_layer_dims)
        Ks = _cabi.int_array(self.num_emb_list)
        h = ctypes.c_void_p(None)
        check(_cabi.lib().rqb200_model_create(ctypes.byref(h), index, len(self.encode_layer_dims) - 1, dims,
                                              len(self.num_emb_list), Ks))
        self._handle = h
        self._handle_device = index
        self._synced = {}

    @staticmethod
    def _fold_bn(lin: nn.Linear, bn: Optional[nn.BatchNorm1d]):
        W = lin.weight.detach()
        b = lin.bias.detach() if lin.bias is not None else torch.zeros(W.shape[0], device=W.device)
        if bn is None:
            return W.contiguous(), b.contiguous()
        # eval-mode BatchNorm is an affine map; folded (not bit-exact vs Linear→BN, tolerance path)
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        return (W * scale[:, None]).contiguous(), ((b - bn.running_mean) * scale + bn.bias.detach()).contiguous()

    def _sync(self):
        """Upload parameters whose storage or version changed since the last call."""
        self._ensure_handle()
        L = _cabi.lib()
        with torch.cuda.device(self._handle_device):
            for which, mlp, tag in ((0, self.encoder, "enc"), (1, self.decoder, "dec")):
                for i, (lin, bn) in enumerate(mlp.linears()):
                    key = (tag, i)
                    sig = (lin.weight.data_ptr(), lin.weight._version,
                           None if lin.bias is None else (lin.bias.data_ptr(), lin.bias._version),
                           None if bn is None else (bn.weight._version, bn.bias._version,
                                                    bn.running_mean._version, bn.running_var._version),
                           tuple(self.kblocks.get(key, ())))
                    if self._synced.get(key) == sig:
                        continue
                    W, b = self._fold_bn(lin, bn)
                    kb = self.kblocks.get(key)
                    kb_arr = _cabi.int_array(kb) if kb else None
                    torch.cuda.current_stream().synchronize()
                    check(L.rqb200_model_set_linear(self._handle, which, i, ptr(W.float()), ptr(b.float()), kb_arr,
                                                    len(kb) if kb else 0))
                    self._synced[key] = sig
            for lvl, q in enumerate(self.rq.vq_layers):
                w = q.embedding.weight
                sig = (w.data_ptr(), w._version)
                if self._synced.get(("cb", lvl)) == sig:
                    continue
                torch.cuda.current_stream().synchronize()
                check(L.rqb200_model_set_codebook(self._handle, lvl, ptr(w.detach().float().contiguous())))
                self._synced[("cb", lvl)] = sig

    def set_gate(self, gamma: float, floor_abs: float = 1e-3):
        """Margin-gate parameters of the tensor-core route (see rqb200_model_set_gate)."""
        self._ensure_handle()
        check(_cabi.lib().rqb200_model_set_gate(self._handle, float(gamma), float(floor_abs)))

    def fast_route_supported(self) -> bool:
        """Shapes the tensor-core route is built for (everything else runs the exact SIMT route)."""
        dims = list(self.encode_layer_dims)
        return (not self.bn and dims[0] % 8 == 0 and all(d in (32, 64, 128, 256) for d in dims[1:])
                and self.e_dim in (32, 64))

    def set_screen(self, enabled, gamma1: float = 0.0):
        """Screening tier of the fast route (rqb200_model_set_screen): 0 / False off, 2 / "tf32" the TMA-fed TF32 first
        layer, 1 / True one fp16 pass through all layers."""
        self._ensure_handle()
        kind = 2 if enabled == "tf32" else int(enabled)
        check(_cabi.lib().rqb200_model_set_screen(self._handle, kind, float(gamma1)))

    @torch.no_grad()
    def encode_tc(self, x: torch.Tensor, passes: int = 3) -> torch.Tensor:
        """Encoder MLP on the tensor cores (not bit-exact) — diagnostic / building block.  passes = 3: split-fp16,
        fp32-class accuracy; 2: TF32 first layer fed by TMA + three-pass tail (screening tier); 1: one fp16 pass."""
        _require_cuda_tensor(x, "input")
        self._sync()
        x2 = x.reshape(-1, self.in_dim).contiguous()
        y = torch.empty((x2.shape[0], self.e_dim), dtype=torch.float32, device=x.device)
        check(_cabi.lib().rqb200_debug_mlp_tc(self._handle, 0, ptr(x2), x2.shape[0], ptr(y), int(passes), stream_ptr(x.device)))
        return y

    # ---- building blocks -----------------------------------------------------------------
    @torch.no_grad()
    def _mlp(self, which: int, x: torch.Tensor) -> torch.Tensor:
        _require_cuda_tensor(x, "input")
        self._sync()
        dims = self.encode_layer_dims if which == 0 else self.decode_layer_dims
        x2 = x.reshape(-1, dims[0]).contiguous()
        y = torch.empty((x2.shape[0], dims[-1]), dtype=torch.float32, device=x.device)
        check(_cabi.lib().rqb200_mlp_exact(self._handle, which, ptr(x2), 0, x2.shape[0], ptr(y), stream_ptr(x.device)))
        return y.view(*x.shape[:-1], dims[-1])

    def _distances(self, level: int, r: torch.Tensor) -> torch.Tensor:
        K = self.num_emb_list[level]
        d = torch.empty((r.shape[0], K), dtype=torch.float32, device=r.device)
        check(_cabi.lib().rqb200_distances(self._handle, level, ptr(r), r.shape[0], ptr(d), stream_ptr(r.device)))
        return d

    def _sinkhorn_assign(self, d: torch.Tensor, eps: float, iters: int) -> torch.Tensor:
        B, K = d.shape
        scratch = torch.empty((B, K), dtype=torch.float64, device=d.device)
        idx = torch.empty((B,), dtype=torch.int64, device=d.device)
        check(_cabi.lib().rqb200_sinkhorn_assign(ptr(d), B, K, float(eps), int(iters), ptr(scratch), ptr(idx),
                                                 stream_ptr(d.device)))
        return idx

    @torch.no_grad()
    def _vq_level(self, level: int, x: torch.Tensor, use_sk: bool):
        """Generic single-level path (vq.py:63-99), used when Sinkhorn is active on a level."""
        _require_cuda_tensor(x, "latent")
        q = self.rq.vq_layers[level]
        latent = x.reshape(-1, self.e_dim).contiguous()
        if not q.initted and self.training:
            q.init_emb(latent)
        self._sync()
        d = self._distances(level, latent)
        if not use_sk or q.sk_epsilon <= 0:
            indices = torch.argmin(d, dim=-1)
        else:
            indices = self._sinkhorn_assign(d, q.sk_epsilon, q.sk_iters)
        # gather, straight-through output and the loss numerator in one kernel (vq.py:87-95): x_q = x + (E[idx] - x)
        cb = q.embedding.weight.detach().contiguous()
        n = latent.shape[0]
        x_q = torch.empty_like(latent)
        r_next = torch.empty_like(latent)
        sumsq = torch.zeros((1,), dtype=torch.float64, device=latent.device)
        if n:
            check(_cabi.lib().rqb200_rq_level_apply(ptr(latent), ptr(indices), ptr(cb), n, self.e_dim, 1, ptr(x_q), ptr(r_next),
                                                    ptr(sumsq), stream_ptr(latent.device)))
        mse = (sumsq[0] / float(max(n * self.e_dim, 1))).to(torch.float32)
        loss = mse + q.beta * mse
        return x_q.view(x.shape), loss, indices.view(x.shape[:-1])

    @torch.no_grad()
    def _rq(self, x: torch.Tensor, use_sk: bool):
        _require_cuda_tensor(x, "latent")
        layers = self.rq.vq_layers
        needs_init = self.training and any(not q.initted for q in layers)
        sk_active = use_sk and any(q.sk_epsilon > 0 for q in layers)
        if needs_init or sk_active:
            # level by level, exactly like rq.py:43-54
            all_losses, all_indices = [], []
            x_q = 0
            residual = x
            for lvl in range(len(layers)):
                x_res, loss, ind = self._vq_level(lvl, residual, use_sk)
                residual = residual - x_res
                x_q = x_q + x_res
                all_losses.append(loss)
                all_indices.append(ind)
            return x_q, torch.stack(all_losses).mean(), torch.stack(all_indices, dim=-1)
        # fused kernel: all levels, residual in registers
        self._sync()
        z = x.reshape(-1, self.e_dim).contiguous()
        n, Lv = z.shape[0], len(layers)
        codes = torch.empty((n, Lv), dtype=torch.int64, device=z.device)
        xq = torch.empty_like(z)
        sumsq = torch.zeros((Lv,), dtype=torch.float64, device=z.device)
        check(_cabi.lib().rqb200_quantize(self._handle, ptr(z), n, ptr(codes), 0, ptr(xq), ptr(sumsq), 0,
                                          stream_ptr(z.device)))
        mse = sumsq / float(max(n * self.e_dim, 1))
        betas = torch.tensor([q.beta for q in layers], dtype=torch.float64, device=z.device)
        mean_loss = (mse + betas * mse).mean().to(torch.float32)
        return xq.view(x.shape), mean_loss, codes.view(*x.shape[:-1], Lv)

    # ---- reference API -------------------------------------------------------------------
    def forward(self, x, use_sk=True):
        """rqvae.py:60-65.  In training mode with gradients enabled this builds the autograd graph of the training
        step (train.py:113); in eval mode or under no_grad it is the inference forward (no graph is kept — the one
        deviation from the reference, where an eval-mode forward outside no_grad still records one)."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(x, use_sk)
        return self._forward_inference(x, use_sk)

    @torch.no_grad()
    def _forward_inference(self, x, use_sk=True):
        x_e = self.encoder(x)
        x_q, rq_loss, indices = self.rq(x_e, use_sk=use_sk)
        out = self.decoder(x_q)
        return out, rq_loss, indices

    def _mlp_params(self, mlp):
        if self.bn:
            raise NotImplementedError("BatchNorm training is not part of the reference's default path (bn=False)")
        out = []
        for lin, _ in mlp.linears():
            out += [lin.weight, lin.bias]
        return out

    def _forward_train(self, x, use_sk=True):
        """Differentiable forward: Dropout (training mode) → encoder → residual quantizer with Sinkhorn where
        ``use_sk`` and ε > 0 (k-means codebook init on the first training batch, vq.py:67-68) → decoder."""
        from .train_ops import MLPFunction, RQFunction
        _require_cuda_tensor(x, "input")
        x2 = x.reshape(-1, self.in_dim)
        p = float(self.dropout_prob) if self.training else 0.0
        self._train_calls = getattr(self, "_train_calls", 0) + 1
        seed = (int(getattr(self, "dropout_seed", 2024)) << 32) ^ self._train_calls
        seed_dev = getattr(self, "_dropout_seed_dev", None)      # set by the Trainer when the step is a replayed CUDA graph
        z = MLPFunction.apply(x2, p, seed * 2, seed_dev, *self._mlp_params(self.encoder))
        layers = self.rq.vq_layers
        if self.training and any(not q.initted for q in layers):
            with torch.no_grad():                         # rq.py:43-48 level by level on this first batch
                r = z.detach()
                for q in layers:
                    if not q.initted:
                        q.init_emb(r)
                    e_l = float(q.sk_epsilon) if (use_sk and q.sk_epsilon is not None and q.sk_epsilon > 0) else 0.0
                    _, _, idx = RQFunction.apply(r, [q.beta], [e_l], q.sk_iters, q.embedding.weight.detach())
                    qv = q.embedding.weight.detach()[idx[:, 0]]
                    r = r - (r + (qv - r))
        eps = [float(q.sk_epsilon) if (use_sk and q.sk_epsilon is not None and q.sk_epsilon > 0) else 0.0 for q in layers]
        x_q, rq_loss, indices = RQFunction.apply(z, [q.beta for q in layers], eps, int(layers[0].sk_iters),
                                                 *[q.embedding.weight for q in layers])
        out = MLPFunction.apply(x_q, p, seed * 2 + 1, seed_dev, *self._mlp_params(self.decoder))
        return out.view(*x.shape[:-1], self.in_dim), rq_loss, indices.view(*x.shape[:-1], len(layers))

    @property
    def last_stats(self) -> dict:
        """Rows the last fast-route `get_indices` re-ran on its three-pass and exact tiers (rqb200_model_last_tier_rows: the
        counts stay on the device until somebody asks, so get_indices itself never waits for the GPU)."""
        if not getattr(self, "_stats_fast", False) or not self._handle.value:
            return {"rescued_rows": 0, "three_pass_rows": 0}
        tiers = (ctypes.c_int64 * 2)()
        check(_cabi.lib().rqb200_model_last_tier_rows(self._handle, tiers))
        return {"rescued_rows": int(tiers[1]), "three_pass_rows": int(tiers[0])}

    @torch.no_grad()
    def get_indices(self, xs, use_sk=False):
        """rqvae.py:67-71.  With use_sk=False (the catalogue pass) this is one fused C-ABI call."""
        _require_cuda_tensor(xs, "xs")
        sk_active = use_sk and any(q.sk_epsilon > 0 for q in self.rq.vq_layers)
        if sk_active or (self.training and any(not q.initted for q in self.rq.vq_layers)):
            x_e = self.encoder(xs)
            _, _, indices = self.rq(x_e, use_sk=use_sk)
            return indices
        self._sync()
        x2 = xs.reshape(-1, self.in_dim).contiguous()
        n, Lv = x2.shape[0], len(self.num_emb_list)
        if torch_ops.available():
            # the registered op (csrc/torch_ops.cpp): same C-ABI call, output from the caching allocator, current stream
            codes = torch.ops.rqvae_b200.encode_indices(torch_ops.handle(self), x2, int(self.encode_mode))
        else:
            codes = torch.empty((n, Lv), dtype=torch.int64, device=xs.device)
            check(_cabi.lib().rqb200_get_indices(self._handle, int(self.encode_mode), ptr(x2), n, ptr(codes), 0, None,
                                                 stream_ptr(xs.device)))
        self._stats_fast = self.encode_mode == _cabi.ENCODE_FAST          # nothing waited: the tier counts are read on demand
        return codes.view(*xs.shape[:-1], Lv)

    def compute_loss(self, out, quant_loss, xs=None):
        """rqvae.py:73-84 (the reconstruction term and its gradient run in the CUDA library)."""
        if self.loss_type not in ("mse", "l1"):
            raise ValueError("incompatible loss type")
        from .train_ops import ReconFunction
        _require_cuda_tensor(out, "out")
        loss_recon = ReconFunction.apply(out, xs, self.loss_type == "l1")
        loss_total = loss_recon + self.quant_loss_weight * quant_loss
        return loss_total, loss_recon

    @torch.no_grad()
    def forward_losses(self, x):
        """Fused forward(use_sk=False) + compute_loss through one C-ABI call (rqb200_forward):
        returns (out, rq_loss, indices, loss_total, loss_recon)."""
        _require_cuda_tensor(x, "x")
        if self.loss_type not in ("mse", "l1"):
            raise ValueError("incompatible loss type")
        self._sync()
        x2 = x.reshape(-1, self.in_dim).contiguous()
        n, Lv = x2.shape[0], len(self.num_emb_list)
        out = torch.empty_like(x2)
        codes = torch.empty((n, Lv), dtype=torch.int64, device=x.device)
        sumsq = torch.zeros((Lv,), dtype=torch.float64, device=x.device)
        recon = torch.zeros((2,), dtype=torch.float64, device=x.device)
        check(_cabi.lib().rqb200_forward(self._handle, ptr(x2), n, ptr(out), ptr(codes), ptr(sumsq), ptr(recon),
                                         stream_ptr(x.device)))
        mse = sumsq / float(max(n * self.e_dim, 1))
        betas = torch.tensor([q.beta for q in self.rq.vq_layers], dtype=torch.float64, device=x.device)
        rq_loss = (mse + betas * mse).mean()
        loss_recon = (recon[0] if self.loss_type == "mse" else recon[1]) / float(max(x2.numel(), 1))
        total = loss_recon + self.quant_loss_weight * rq_loss
        return (out.view(x.shape), rq_loss.to(torch.float32), codes.view(*x.shape[:-1], Lv),
                total.to(torch.float32), loss_recon.to(torch.float32))

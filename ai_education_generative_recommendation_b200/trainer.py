"""B200-side mirror of the reference's training driver (RQ-VAE/train.py): same `Trainer(params, model, data_num)`
surface, same params keys (main.py:6-36), same checkpoint dictionary (train.py:153-171) and the same schedule
(transformers.get_linear_schedule_with_warmup / get_constant_schedule_with_warmup, restated in `warmup_lambda`).

What changes underneath:
  * forward / backward of every layer run in librqvae_b200.so through `train_ops` (no torch arithmetic);
  * `clip_grad_norm_(…, 1.0)` + `AdamW.step()` (train.py:116-117) is one fused C-ABI call (`FusedAdamW`);
  * losses stay on the device during an epoch — the reference's per-step `.item()` (train.py:120-121) and NaN check
    (train.py:92-94,115) become one read at the end of the epoch (`ValueError("Training loss is nan")` is still raised);
  * data-parallel: one process per GPU, each rank trains on its shard of every batch, the flat gradient buffer is
    all-reduced (NCCL over NVLink) before the fused step, which folds in the 1/world averaging.  The Sinkhorn
    balancing of a batch (vq.py:77-83) then acts per rank — stated, not hidden: the reference has no multi-GPU mode;
  * `DeviceBatches`: a catalogue resident in HBM is shuffled and sliced on the device (randperm + row gather), so an
    epoch over millions of items is not bounded by a host DataLoader;
  * the whole step (zero_grad → forward → losses → backward → [all-reduce] → clip + AdamW) is ≈ 70 of our launches plus
    torch's graph glue, ≈ 3 ms of GPU time at batch 4096 but 4–6 ms of Python: after three eager steps it is captured
    once as a CUDA graph and replayed per batch (`params["cuda_graph"]`, default on, single-GPU runs only; batches of
    another size and any capture failure fall back to the eager step).  Per-step scalars (learning rate, bias corrections, dropout seed)
    live in device memory so the replay is exact.
"""
from __future__ import annotations

import collections
import logging
import os
from time import time
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from .rqvae import RQVAE
from .train_ops import FusedAdamW


def warmup_lambda(kind: str, warmup_steps: int, max_steps: int):
    """The LambdaLR factor of transformers' linear / constant schedule with warmup (train.py:80-91)."""
    def linear(step: int) -> float:
        if step < warmup_steps:
            return float(step) / float(max(1, warmup_steps))
        return max(0.0, float(max_steps - step) / float(max(1, max_steps - warmup_steps)))

    def constant(step: int) -> float:
        if step < warmup_steps:
            return float(step) / float(max(1.0, warmup_steps))
        return 1.0
    return linear if kind.lower() == "linear" else constant


class DeviceBatches:
    """Shuffled mini-batches of a catalogue that already lives on the GPU (the DataLoader(shuffle=True) of
    train.py:250-252 without the host round trip).  `len()` = batches per epoch, like a DataLoader."""

    def __init__(self, embeddings: torch.Tensor, batch_size: int, shuffle: bool = True, seed: int = 2024,
                 drop_last: bool = False):
        if not embeddings.is_cuda:
            raise RuntimeError("DeviceBatches needs a CUDA tensor (use a DataLoader for host data)")
        self.x = embeddings.contiguous()
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.gen = torch.Generator(device=self.x.device).manual_seed(seed)

    def __len__(self):
        n = self.x.shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.x.shape[0]
        order = torch.randperm(n, device=self.x.device, generator=self.gen) if self.shuffle else None
        for b in range(len(self)):
            lo, hi = b * self.batch_size, min(n, (b + 1) * self.batch_size)
            yield self.x[order[lo:hi]] if order is not None else self.x[lo:hi]


class Trainer(object):
    def __init__(self, params, model: RQVAE, data_num: int, group=None):
        self.params = params
        self.model = model
        self.logger = logging.getLogger()
        self.lr = params["lr"]
        self.learner = params["learner"]
        self.lr_scheduler_type = params["lr_scheduler_type"]
        self.weight_decay = params["weight_decay"]
        self.epochs = params["epochs"]
        self.warmup_steps = params["warmup_epochs"] * data_num
        self.max_steps = params["epochs"] * data_num
        self.save_limit = params["save_limit"]
        self.eval_step = min(params["eval_step"], self.epochs)
        self.device = torch.device(params["device"])
        self.ckpt_dir = params["ckpt_dir"]
        os.makedirs(self.ckpt_dir, exist_ok=True)
        self.best_loss = np.inf
        self.best_collision_rate = np.inf
        self.best_loss_ckpt = "best_loss_model.pth"
        self.best_collision_ckpt = "best_collision_model.pth"
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.model = self.model.to(self.device)
        self.slice_batches = True        # world > 1: every rank sees the same batches and keeps its rows; set False
        #                                  when each rank iterates over its own shard of the catalogue
        if self.world > 1:
            for q in self.model.rq.vq_layers:
                object.__setattr__(q, "_kmeans_group", group)
        self.optimizer = self._build_optimizer()
        self.scheduler = self._get_scheduler()
        self.last_epoch_steps = 0
        # single-GPU only: with the NCCL all-reduce inside the captured graph the step got slower (6.8 vs 4.0 ms at N = 2)
        # and process-group teardown hung (measured once, round 1) — data-parallel runs keep the eager step
        self.use_cuda_graph = bool(params.get("cuda_graph", True)) and self.world == 1
        self._graph = None               # (CUDAGraph, static input, static [loss, recon]) once captured
        self._graph_failed = False
        self._eager_steps = 0
        self.graph_replays = 0

    def _build_optimizer(self):
        if self.learner.lower() != "adamw":
            # train.py:51-79 offers adam / sgd / adagrad / rmsprop too; main.py:29 (the shipped configuration) uses AdamW
            raise NotImplementedError(f"learner '{self.learner}': only AdamW (main.py:29) has a CUDA step here")
        opt = FusedAdamW(self.model.parameters(), lr=self.lr, weight_decay=self.weight_decay, max_norm=1.0)
        opt.grad_scale = 1.0 / self.world
        return opt

    def _get_scheduler(self):
        lam = warmup_lambda(self.lr_scheduler_type, self.warmup_steps, self.max_steps)
        return torch.optim.lr_scheduler.LambdaLR(self.optimizer, lam)

    def _check_nan(self, loss):
        if torch.isnan(loss):
            raise ValueError("Training loss is nan")

    def _eager_step(self, data):
        self.optimizer.zero_grad()
        out, rq_loss, indices = self.model(data)
        loss, loss_recon = self.model.compute_loss(out, rq_loss, xs=data)
        loss.backward()
        if self.world > 1:
            dist.all_reduce(self.optimizer.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.optimizer.step()
        self._eager_steps += 1
        return torch.stack([loss.detach(), loss_recon.detach()]).to(torch.float64)

    def _capture(self, data):
        """Capture one training step for batches shaped like `data` (called after the eager warm-up steps)."""
        model, opt = self.model, self.optimizer
        if getattr(model, "_dropout_seed_dev", None) is None:
            object.__setattr__(model, "_dropout_seed_dev", torch.zeros((1,), dtype=torch.int64, device=self.device))
        static_x = data.clone()
        static_out = torch.zeros((2,), dtype=torch.float64, device=self.device)
        opt.refresh_hyper()                      # allocates the device scalars; the count is set back below
        opt._steps -= 1
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            model._dropout_seed_dev.add_(1)
            opt.zero_grad()
            out, rq_loss, indices = model(static_x)
            loss, loss_recon = model.compute_loss(out, rq_loss, xs=static_x)
            loss.backward()
            if self.world > 1:
                dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            opt.step_from_device()
            static_out.copy_(torch.stack([loss.detach(), loss_recon.detach()]).to(torch.float64))
        return graph, static_x, static_out

    def _train_step(self, data):
        """One optimisation step on `data`; returns a device tensor [loss, recon]."""
        ready = all(q.initted for q in self.model.rq.vq_layers)
        if (self.use_cuda_graph and not self._graph_failed and self._graph is None and ready and self._eager_steps >= 3
                and data.shape[0] == getattr(self, "_last_rows", -1)):
            try:
                self._graph = self._capture(data)
            except Exception as exc:       # capture is an optimisation: fall back to the eager step
                self._graph_failed = True
                self.logger.warning(f"CUDA graph capture of the training step failed ({exc}); continuing eagerly")
                torch.cuda.synchronize()
        self._last_rows = data.shape[0]
        if self._graph is not None and data.shape == self._graph[1].shape:
            graph, static_x, static_out = self._graph
            static_x.copy_(data)
            self.optimizer.refresh_hyper()
            graph.replay()
            self.optimizer.after_replay()
            self.graph_replays += 1
            return static_out.clone()
        return self._eager_step(data)

    def _train_epoch(self, train_data, epoch_idx):
        self.model.train()
        sums = torch.zeros((2,), dtype=torch.float64, device=self.device)
        steps = 0
        for data in train_data:
            data = data.to(self.device, non_blocking=True)
            if self.world > 1 and self.slice_batches:   # this rank's rows of the batch
                lo, hi = (self.rank * data.shape[0]) // self.world, ((self.rank + 1) * data.shape[0]) // self.world
                data = data[lo:hi]
            sums += self._train_step(data.contiguous())
            self.scheduler.step()
            steps += 1
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
            sums /= self.world
        total_loss, total_recon_loss = (float(v) for v in sums.cpu())
        self._check_nan(torch.tensor(total_loss))
        self.last_epoch_steps = steps
        return total_loss, total_recon_loss

    @torch.no_grad()
    def _valid_epoch(self, valid_data):
        """Collision rate over the data (train.py:126-151) — distinct codes counted by the dedup kernels instead of a
        Python set of strings."""
        from .generate_code import suffix_dedup
        self.model.eval()
        chunks = []
        for data in valid_data:
            data = data.to(self.device, non_blocking=True)
            if self.world > 1 and self.slice_batches:   # every rank iterates over the same batches: keep this rank's rows,
                lo, hi = (self.rank * data.shape[0]) // self.world, ((self.rank + 1) * data.shape[0]) // self.world
                data = data[lo:hi].contiguous()         # so that every item is counted once in the global statistics
            if data.shape[0]:
                chunks.append(self.model.get_indices(data).view(-1, len(self.model.num_emb_list)))
        codes = torch.cat(chunks) if chunks else torch.zeros((0, len(self.model.num_emb_list)), dtype=torch.int64, device=self.device)
        if self.world > 1:
            from .sharding import CudaShardOps, global_stats, global_suffix
            ops = getattr(self, "shard_ops", None) or CudaShardOps(self.model)
            out = global_suffix(codes, self.model.num_emb_list, ops, self.group)
            return global_stats(out, self.group)["collision_rate"]
        _, stats = suffix_dedup(self.model, codes)
        return stats["collision_rate"]

    def _save_checkpoint(self, epoch, collision_rate=1, ckpt_file=None):
        ckpt_path = os.path.join(self.ckpt_dir, ckpt_file) if ckpt_file \
            else os.path.join(self.ckpt_dir, "epoch_%d_collision_%.4f_model.pth" % (epoch, collision_rate))
        if self.rank == 0:
            state = {
                "args": self.params,
                "epoch": epoch,
                "best_loss": self.best_loss,
                "best_collision_rate": self.best_collision_rate,
                "state_dict": self.model.state_dict(),
                "optimizer": self.optimizer.state_dict(),
            }
            torch.save(state, ckpt_path, pickle_protocol=4)
            self.logger.info(f"Saving current: {ckpt_path}")
        return ckpt_path

    def fit(self, data):
        """train.py:186-250: epochs of training; every `eval_step` epochs the collision rate is measured, the two
        "best" checkpoints are refreshed and an epoch checkpoint is written and filed with the keeper."""
        keeper = _CheckpointKeeper(self.save_limit, self.rank)
        self.checkpoints = keeper
        for epoch in range(self.epochs):
            started = time()
            loss, recon = self._train_epoch(data, epoch)
            self.logger.info("epoch %d training [time: %.2fs, train loss: %.4f, reconstruction loss: %.4f]"
                             % (epoch, time() - started, loss, recon))
            if (epoch + 1) % self.eval_step:
                continue
            started = time()
            rate = self._valid_epoch(data)
            if loss < self.best_loss:
                self.best_loss = loss
                self._save_checkpoint(epoch=epoch, ckpt_file=self.best_loss_ckpt)
            if rate < self.best_collision_rate:
                self.best_collision_rate = rate
                self._save_checkpoint(epoch, collision_rate=rate, ckpt_file=self.best_collision_ckpt)
            self.logger.info("epoch %d evaluating [time: %.2fs, collision_rate: %f]" % (epoch, time() - started, rate))
            keeper.file(rate, self._save_checkpoint(epoch, collision_rate=rate))
        return self.best_loss, self.best_collision_rate


class _CheckpointKeeper:
    """Which epoch checkpoints stay on disk (the rule of train.py:232-248, stated as sets): the `limit` most recent ones
    and the `limit` with the lowest collision rate; a file is removed when it belongs to neither set any more.  As in
    the reference a checkpoint enters the "best" set only while that set is not full or when it beats the worst member
    (ties: the worst member is the one with the highest rate, then the smallest path)."""

    def __init__(self, limit: int, rank: int = 0):
        self.limit, self.rank = int(limit), rank
        self.recent = collections.deque()
        self.best = []

    def file(self, rate: float, path: str):
        entry = (rate, path)
        if len(self.recent) < self.limit:
            self.recent.append(entry)
            self.best.append(entry)
            return
        leaving = [self.recent.popleft()]
        self.recent.append(entry)
        worst = min(self.best, key=lambda e: (-e[0], e[1]))
        if rate < worst[0]:
            self.best.remove(worst)
            self.best.append(entry)
            leaving.insert(0, worst)
        for old in leaving:
            if old not in self.recent and old not in self.best:
                _delete_file(old[1], self.rank)


def _delete_file(path, rank=0):
    if rank == 0 and os.path.exists(path):
        os.remove(path)


def train(params, group=None, on_device: bool = True):
    """train.py:253-287: seeds, dataset, model, loader, Trainer.fit.  `on_device=True` keeps the catalogue in HBM and
    batches it there (`DeviceBatches`); False uses the reference's host DataLoader."""
    import random
    from .dataset import EmbDataset
    seed = 2024
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    logging.basicConfig(level=logging.DEBUG)
    data = EmbDataset(params["data_path"])
    model = RQVAE(in_dim=data.dim, num_emb_list=params["num_emb_list"], e_dim=params["e_dim"], layers=params["layers"],
                  dropout_prob=params["dropout"], bn=params["batch_normalize"], loss_type=params["loss_type"],
                  quant_loss_weight=params["quant_loss_weight"], beta=params["beta"], kmeans_init=params["kmeans_init"],
                  kmeans_iters=params["kmeans_iters"], sk_epsilons=params["sk_epsilons"], sk_iters=params["sk_iters"])
    print(model)
    if on_device:
        loader = DeviceBatches(torch.from_numpy(np.ascontiguousarray(data.embeddings)).to(params["device"]),
                               params["batch_size"], shuffle=True, seed=seed)
    else:
        loader = torch.utils.data.DataLoader(data, num_workers=params["num_workers"], batch_size=params["batch_size"],
                                             shuffle=True, pin_memory=True)
    trainer = Trainer(params, model, len(loader), group=group)
    best_loss, best_collision_rate = trainer.fit(loader)
    print("Best Loss", best_loss)
    print("Best Collision Rate", best_collision_rate)
    return best_loss, best_collision_rate

"""Compute engines: forward + backward of one replica on one batch.

An engine owns the flat fp32 parameter arena and the flat gradient arena of its
replica (allocated through the aggregation backend so that, on the GPU path,
both live in NVLink-symmetric memory) and exposes four calls:

* ``load_batch(images, labels)``  -- host (pinned) -> device copy of the step's inputs
  (``pack_batch`` / ``load_packed``: the same with one contiguous buffer and one copy per step)
* ``forward_backward(step)``      -- loss, train accuracy and all gradients
* ``loss_acc()``                  -- device -> host read of the step's scalars
* ``evaluate(images, labels)``    -- inference-only loss/accuracy (evaluator)

:class:`TorchEngine` is the plain-PyTorch implementation: the execution path on
CPU (gloo plumbing config) and the cuDNN/cuBLAS *baseline* on GPU.  The product
path on B200 is ``CudaLeNetEngine`` / ``CudaMlpEngine`` in ``engine_cuda.py``
(hand-written sm_100a kernels).  Unlike the reference, loss and accuracy come
from the training forward itself -- the reference runs a second forward per
iteration only for logging (src/distributed_train.py:334).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch

from .models import ModelSpec, dropout_keep_mask, dropout_seed_mix, get_model, loss_and_accuracy
from .models.lenet import FC1_OUT


class ComputeEngine:
    spec: ModelSpec
    params: torch.Tensor
    grads: torch.Tensor
    batch_size: int

    def load_batch(self, images, labels) -> None:
        raise NotImplementedError

    def forward_backward(self, step: int) -> None:
        raise NotImplementedError

    def loss_acc(self) -> Tuple[float, float]:
        raise NotImplementedError

    def evaluate(self, images, labels) -> Tuple[float, float]:
        raise NotImplementedError

    def h2d_bytes_per_step(self) -> int:
        raise NotImplementedError

    # ---- packed batches: one contiguous buffer per step (fp32 images [B,784] then int64 labels [B]) -----------------
    def pack_batch(self, images, labels, pin: bool = True) -> torch.Tensor:
        """Input pipelines fill one (page-locked) buffer per batch; ``load_packed`` then moves it with a single copy
        (the sm_100a engines keep images and labels of a slot in one device allocation for exactly this)."""
        B = self.batch_size
        buf = torch.empty(B * 784 * 4 + B * 8, dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            buf = buf.pin_memory()
        buf[:B * 784 * 4].view(torch.float32).copy_(torch.as_tensor(images, dtype=torch.float32).reshape(B * 784))
        buf[B * 784 * 4:].view(torch.int64).copy_(torch.as_tensor(labels, dtype=torch.int64).reshape(B))
        return buf

    def load_packed(self, packed: torch.Tensor) -> None:
        B = self.batch_size
        self.load_batch(packed[:B * 784 * 4].view(torch.float32).view(B, 28, 28, 1), packed[B * 784 * 4:].view(torch.int64))


class TorchEngine(ComputeEngine):
    def __init__(self, model: str, batch_size: int, device: torch.device, allocate: Callable[[int], torch.Tensor],
                 seed: int = 66478, rank: int = 0, keep_prob: float = 0.5, mlp_hidden: int = 1024,
                 autocast_bf16: bool = False):
        self.spec, self._fwd = get_model(model, mlp_hidden)
        self.model = model
        self.batch_size = batch_size
        self.device = device
        self.seed, self.rank, self.keep_prob = seed, rank, keep_prob
        self.autocast_bf16 = autocast_bf16 and device.type == "cuda"
        self.params = allocate(self.spec.arena_numel)
        self.params.copy_(self.spec.init_flat(seed))
        self.grads = allocate(self.spec.arena_numel)
        self._images: Optional[torch.Tensor] = None
        self._labels: Optional[torch.Tensor] = None
        self._loss = torch.zeros((), device=device)
        self._acc = torch.zeros((), device=device)
        self._h2d = 0

    def load_batch(self, images, labels) -> None:
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images))
            labels = torch.from_numpy(np.ascontiguousarray(labels))
        self._h2d = images.numel() * images.element_size() + labels.numel() * labels.element_size()
        nb = self.device.type == "cuda"
        self._images = images.to(self.device, non_blocking=nb)
        self._labels = labels.to(self.device, non_blocking=nb)

    def _logits(self, flat: torch.Tensor, images: torch.Tensor, train: bool, step: int) -> torch.Tensor:
        views = self.spec.views(flat)
        mask = None
        if train and self.model == "lenet":
            mix = dropout_seed_mix(self.seed, step, self.rank)
            mask = dropout_keep_mask(mix, images.shape[0], FC1_OUT, self.keep_prob, device=str(flat.device))
        if self.autocast_bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self._fwd(views, images, train=train, keep_mask=mask, keep_prob=self.keep_prob).float()
        return self._fwd(views, images, train=train, keep_mask=mask, keep_prob=self.keep_prob)

    def forward_backward(self, step: int) -> None:
        p = self.params.detach().requires_grad_(True)
        logits = self._logits(p, self._images, True, step)
        loss, acc = loss_and_accuracy(logits, self._labels)
        (g,) = torch.autograd.grad(loss, p)
        self.grads.copy_(g)
        self._loss, self._acc = loss.detach(), acc.detach()

    def loss_acc(self) -> Tuple[float, float]:
        return float(self._loss.item()), float(self._acc.item())

    @torch.no_grad()
    def evaluate(self, images, labels) -> Tuple[float, float]:
        if isinstance(images, np.ndarray):
            images, labels = torch.from_numpy(np.ascontiguousarray(images)), torch.from_numpy(np.ascontiguousarray(labels))
        images, labels = images.to(self.device), labels.to(self.device)
        tot_loss, tot_hit, n = 0.0, 0.0, images.shape[0]
        for s in range(0, n, 2048):
            lg = self._logits(self.params, images[s:s + 2048], False, 0)
            loss, acc = loss_and_accuracy(lg, labels[s:s + 2048])
            m = lg.shape[0]
            tot_loss += float(loss) * m
            tot_hit += float(acc) * m
        return tot_loss / n, tot_hit / n

    def h2d_bytes_per_step(self) -> int:
        return self._h2d


def make_engine(flags, ctx, backend, force_torch: bool = False) -> ComputeEngine:
    """GPU -> hand-written sm_100a engine (fails loudly if the extension is absent);
    CPU -> torch engine."""
    if ctx.on_gpu and not force_torch and flags.backend in ("auto", "fused"):
        from .engine_cuda import make_cuda_engine
        return make_cuda_engine(flags, ctx, backend)
    return TorchEngine(flags.model, flags.batch_size, ctx.device, backend.allocate, seed=flags.seed,
                       rank=ctx.rank, keep_prob=flags.dropout_keep_prob, mlp_hidden=flags.mlp_hidden,
                       autocast_bf16=(flags.compute_dtype == "bf16"))

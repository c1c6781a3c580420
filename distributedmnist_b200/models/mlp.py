"""2- and 3-layer MLPs on flattened 28x28 inputs.

Not in the reference's model zoo (its only model is src/mnist.py), but named by
BASELINE.json's configs: the 2-layer MLP is the CPU/gloo plumbing model and the
3-layer MLP at batch 8192/replica enlarges the gradient message to exercise the
large-message path of the fused allreduce+SGD kernel.  Same init distributions
and checkpoint naming convention as the convnet.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .spec import ModelSpec, ParamSpec

IN_FEATURES = 28 * 28


def mlp_spec(num_layers: int = 2, hidden: int = 1024, num_classes: int = 10) -> ModelSpec:
    assert num_layers in (2, 3)
    dims = [IN_FEATURES] + [hidden] * (num_layers - 1) + [num_classes]
    params = []
    for i in range(num_layers):
        params.append(ParamSpec("fc%d_weights" % (i + 1), (dims[i], dims[i + 1]), "truncated_normal", 0.1))
        params.append(ParamSpec("fc%d_biases" % (i + 1), (dims[i + 1],), "constant", 0.1))
    return ModelSpec(name="mlp%d" % num_layers, params=params)


def mlp_forward(p: Dict[str, torch.Tensor], images: torch.Tensor, train: bool = True,
                emulate_bf16: bool = False, **_unused) -> torch.Tensor:
    def q(t):
        return t.to(torch.bfloat16).to(torch.float32) if emulate_bf16 else t
    x = q(images.reshape(images.shape[0], -1))
    n = len(p) // 2
    for i in range(1, n + 1):
        x = x @ q(p["fc%d_weights" % i]) + p["fc%d_biases" % i]
        if i < n:
            x = q(F.relu(x))
    return x

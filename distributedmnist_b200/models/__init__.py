"""Model families: the reference convnet and the MLPs named by BASELINE.json."""
from __future__ import annotations

from typing import Callable, Tuple

from .lenet import (dropout_keep_mask, dropout_seed_mix, lenet_forward, lenet_spec,
                    loss_and_accuracy)
from .mlp import mlp_forward, mlp_spec
from .spec import ModelSpec, ParamSpec


def get_model(name: str, mlp_hidden: int = 1024) -> Tuple[ModelSpec, Callable]:
    """``name`` -> (spec, torch reference forward)."""
    if name == "lenet":
        return lenet_spec(), lenet_forward
    if name == "mlp2":
        return mlp_spec(2, mlp_hidden), mlp_forward
    if name == "mlp3":
        return mlp_spec(3, mlp_hidden), mlp_forward
    raise ValueError("unknown model %r (lenet | mlp2 | mlp3)" % name)


__all__ = ["ModelSpec", "ParamSpec", "get_model", "lenet_spec", "lenet_forward", "mlp_spec",
           "mlp_forward", "loss_and_accuracy", "dropout_keep_mask", "dropout_seed_mix"]

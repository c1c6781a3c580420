"""Compat module for the reference's src/nn_eval.py (``do_eval``, ``evaluate``)."""
import _bootstrap  # noqa: F401

from distributedmnist_b200.evaluator import do_eval, evaluate  # noqa: F401
from distributedmnist_b200.flags import FLAGS  # noqa: F401

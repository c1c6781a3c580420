"""Compat module: ``import distributed_train; distributed_train.train(...)``
(reference src/distributed_train.py:109).  The implementation lives in
``distributedmnist_b200.train``; FLAGS is the shared registry."""
import _bootstrap  # noqa: F401

from distributedmnist_b200.flags import FLAGS  # noqa: F401
from distributedmnist_b200.train import LOG_FORMAT, train  # noqa: F401

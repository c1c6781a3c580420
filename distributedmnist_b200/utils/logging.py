"""``tf.logging``-style logger: ``INFO:dmnist:<message>`` on stderr.

The reference's log lines are a de-facto API (tools/benchmark.py regex-scrapes
them), so messages are emitted verbatim after the level prefix.
"""
from __future__ import annotations

import logging
import sys

_configured = False


def get_logger() -> logging.Logger:
    global _configured
    lg = logging.getLogger("dmnist")
    if not _configured:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter("%(levelname)s:%(name)s:%(message)s"))
        lg.addHandler(h)
        lg.setLevel(logging.INFO)
        lg.propagate = False
        _configured = True
    return lg


def set_verbosity(level: int) -> None:
    get_logger().setLevel(level)

"""Compat module for the reference's src/timeout_manager.py (``launch_manager``,
``TimeoutServer``, ``TimeoutClient``)."""
import _bootstrap  # noqa: F401

from distributedmnist_b200.parallel.timeout_manager import (TimeoutClient, TimeoutServer,  # noqa: F401
                                                            launch_manager)

"""Compat module: ``TimeoutReplicasOptimizer`` under the reference's import path
(src/sync_replicas_optimizer_modified/sync_replicas_optimizer_modified.py:146)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: F401,E402

from distributedmnist_b200.parallel.aggregators import (SyncReplicasOptimizer,  # noqa: F401,E402
                                                        TimeoutReplicasOptimizer)

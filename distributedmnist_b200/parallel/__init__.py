"""Replica parallelism: context, commit protocol, backends, aggregation policies."""
from .aggregators import SyncReplicasOptimizer, TimeoutReplicasOptimizer, parse_straggler_spec  # noqa: F401
from .backends import Backend, GlooBackend, LocalBackend, NcclBackend, StepInfo, make_backend  # noqa: F401
from .context import ReplicaContext, init_context, shutdown_context  # noqa: F401
from .protocol import CommitBoard, Decision, StoreCommitBoard  # noqa: F401

"""Process / replica context: one process per GPU, ``torch.distributed`` for plumbing.

Replaces the reference's cluster bring-up -- ``tf.train.ClusterSpec`` +
in-process ``tf.train.Server`` per task, ``ps`` tasks that ``server.join()``
forever (src/mnist_distributed_train.py:25-35) -- with a process-group
rendezvous.  There is no parameter server: every replica keeps the weights in
its own HBM and coherence comes from the deterministic fused update
(SURVEY §2.5 X4).  ``--job_name=ps`` is accepted and exits successfully.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


@dataclass
class ReplicaContext:
    rank: int
    world_size: int
    local_rank: int
    device: torch.device
    backend: str                  # "gloo" | "nccl" | "none"
    store: Optional[object] = None  # c10d Store (TCPStore / PrefixStore) when world_size > 1

    @property
    def is_chief(self) -> bool:  # reference: is_chief = (FLAGS.task_id == 0), distributed_train.py:130
        return self.rank == 0

    @property
    def on_gpu(self) -> bool:
        return self.device.type == "cuda"


def _env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def resolve_world(flags=None) -> tuple:
    """(rank, world_size, local_rank) from torchrun env, else from reference flags.

    With the reference's flags, ``len(worker_hosts.split(','))`` is the replica
    count and ``task_id`` the rank (mnist_distributed_train.py:20-21, :31).
    """
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        return _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", _env_int("RANK", 0))
    if flags is not None and getattr(flags, "worker_hosts", ""):
        hosts = [h for h in flags.worker_hosts.split(",") if h]
        return int(flags.task_id), max(len(hosts), 1), int(flags.task_id)
    return 0, 1, 0


def init_context(flags=None, want_gpu: Optional[bool] = None, timeout_s: int = 600) -> ReplicaContext:
    rank, world, local_rank = resolve_world(flags)
    use_gpu = torch.cuda.is_available() if want_gpu is None else (want_gpu and torch.cuda.is_available())
    if use_gpu:
        ndev = torch.cuda.device_count()
        device = torch.device("cuda", local_rank % max(ndev, 1))
        torch.cuda.set_device(device)
    else:
        device = torch.device("cpu")
    backend = "none"
    store = None
    if world > 1:
        backend = "nccl" if use_gpu else "gloo"
        if not dist.is_initialized():
            if "MASTER_ADDR" not in os.environ:
                # Reference-flag launch (no torchrun): rendezvous on the first worker
                # host's name and --rpc_port (the reference's control-plane port,
                # timeout_manager.py:124,203).  Single box => 127.0.0.1.
                os.environ["MASTER_ADDR"] = "127.0.0.1"
                os.environ.setdefault("MASTER_PORT", str(getattr(flags, "rpc_port", 1235) if flags else 1235))
            kw = {}
            if use_gpu:
                kw["device_id"] = device
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
        try:
            store = dist.distributed_c10d._get_default_store()
        except Exception:  # pragma: no cover - private API moved
            store = None
    return ReplicaContext(rank=rank, world_size=world, local_rank=local_rank, device=device,
                          backend=backend, store=store)


def shutdown_context(ctx: ReplicaContext) -> None:
    if ctx.world_size > 1 and dist.is_initialized():
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()

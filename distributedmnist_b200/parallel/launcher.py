"""Single-box launcher: one process per replica (per GPU), torchrun-style env.

Replaces the reference's EC2 fan-out -- role assignment, ``host:1234`` strings from
private IPs, SSH ``exec_command`` of templated train commands with stdout
redirected to ``out_<ROLE_ID>`` (tools/tf_ec2.py:445-615) -- for the one-box case:
roles become local processes, the "NFS directory" becomes a local directory, and
the per-role ``out_master`` / ``out_worker_<i>`` / ``out_evaluator`` files keep
their names because tools/benchmark.py reads them.
"""
from __future__ import annotations

import os
import socket
import subprocess
import sys
from typing import Dict, List, Optional, Sequence


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def role_name(rank: int) -> str:
    """Reference naming: task 0 is the "master", others "worker_<i>" with i = task_id-1
    (tools/tf_ec2.py:498-518)."""
    return "master" if rank == 0 else "worker_%d" % (rank - 1)


def spawn_replicas(argv: Sequence[str], nprocs: int, out_dir: Optional[str] = None,
                   env: Optional[Dict[str, str]] = None, port: Optional[int] = None,
                   python: str = sys.executable) -> List[subprocess.Popen]:
    """Start ``nprocs`` copies of ``python argv...`` with RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* set.

    With ``out_dir`` each replica's stdout+stderr goes to ``out_dir/out_<role>``.
    Returns the Popen objects (caller waits / kills by exact PID)."""
    port = port or free_port()
    procs = []
    for rank in range(nprocs):
        e = dict(os.environ)
        e.update(env or {})
        e.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(nprocs),
                  "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)
            f = open(os.path.join(out_dir, "out_%s" % role_name(rank)), "wb")
            p = subprocess.Popen([python] + list(argv), env=e, stdout=f, stderr=subprocess.STDOUT)
            p._dmnist_log = f  # keep the handle alive with the process object
        else:
            p = subprocess.Popen([python] + list(argv), env=e)
        procs.append(p)
    return procs


def wait_all(procs: List[subprocess.Popen], timeout: Optional[float] = None) -> List[int]:
    codes = []
    try:
        for p in procs:
            codes.append(p.wait(timeout=timeout))
    except subprocess.TimeoutExpired:
        for p in procs:
            if p.poll() is None:
                p.kill()   # exact PIDs we started, never a pattern
        raise
    finally:
        for p in procs:
            f = getattr(p, "_dmnist_log", None)
            if f is not None and p.poll() is not None:
                f.close()
    return codes


def run_replicas(argv: Sequence[str], nprocs: int, timeout: Optional[float] = None, **kw) -> List[int]:
    return wait_all(spawn_replicas(argv, nprocs, **kw), timeout=timeout)

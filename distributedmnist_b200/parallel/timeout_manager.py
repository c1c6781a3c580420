"""Start-up barrier + straggler timing telemetry (the reference's "timeout manager").

reference: src/timeout_manager.py -- a Twisted Perspective-Broker full mesh on TCP
``--rpc_port`` doing three jobs: (1) a ready barrier polled at 1 Hz
(:150-158, 207-209); (2) per-iteration ``worker_dequeued_token`` /
``worker_finished_computing_gradients`` broadcasts from which every server
derives ``compute time = finished - dequeued`` and logs ``ELAPSED TIMES`` /
``ITERATION TIMES`` every 50 iterations for iterations 10..500 (:48-70);
(3) a ``parameters_updated`` broadcast whose kill hook is commented out (:38-46).

Here the mesh is the process group that already exists: the barrier is a
rendezvous barrier, and timing events are recorded locally (host clock, or the
``%globaltimer`` stamps the fused kernel writes) and exchanged with one
``all_gather_object`` per reporting period instead of 2*N^2 RPCs per iteration.
The log-line formats are unchanged because tools/benchmark.py scrapes them.
The reference's aliasing bug (``[{}] * n`` shares one dict, :31-32) is not
reproduced: tables are per worker.
"""
from __future__ import annotations

import logging
import time
from typing import Dict, List, Optional, Tuple

from .backends import Backend

log = logging.getLogger("dmnist")

ITERATION_START_TRACKING = 10   # reference timeout_manager.py:35
ITERATION_END_TRACKING = 500    # reference timeout_manager.py:36
REPORT_EVERY = 50               # reference timeout_manager.py:68


class TimeoutServer:
    """Holds the merged timing tables and produces the two log lines."""

    def __init__(self, worker_id: int, n_total_workers: int, n_to_collect: int):
        self.worker_id = worker_id
        self.n_total_workers = n_total_workers
        self.n_to_collect = n_to_collect
        self.ready_to_start = False
        self.worker_dequeue_times: List[Dict[int, float]] = [dict() for _ in range(n_total_workers)]
        self.worker_finished_computing_gradients_times: List[Dict[int, float]] = [dict() for _ in range(n_total_workers)]
        self.compute_times: List[Tuple[int, float, int]] = []   # (worker, elapsed, iteration)
        self.iteration_start_times: Dict[int, float] = {}

    def remote_parameters_updated(self, step: int) -> None:
        log.info("Parameters have been updated on step %d.." % step)

    def remote_worker_dequeued_token(self, worker_id: int, iteration: int, t: Optional[float] = None) -> None:
        t = time.time() if t is None else t
        self.worker_dequeue_times[worker_id][iteration] = t
        if iteration not in self.iteration_start_times or t < self.iteration_start_times[iteration]:
            self.iteration_start_times[iteration] = t

    def remote_worker_finished_computing_gradients(self, worker_id: int, iteration: int,
                                                   t: Optional[float] = None) -> None:
        t = time.time() if t is None else t
        self.worker_finished_computing_gradients_times[worker_id][iteration] = t
        start = self.worker_dequeue_times[worker_id].get(iteration)
        if start is not None:
            self.compute_times.append((worker_id, t - start, iteration))

    def elapsed_times(self) -> List[Tuple[float, int, int]]:
        """Sorted ``(seconds, worker, iteration)`` for tracked iterations (reference :64)."""
        return sorted([(x[1], x[0], x[2]) for x in self.compute_times if x[2] > ITERATION_START_TRACKING],
                      key=lambda x: x[0])

    def iteration_times(self) -> List[float]:
        sel = [t for i, t in sorted(self.iteration_start_times.items()) if i > ITERATION_START_TRACKING]
        return [sel[i + 1] - sel[i] for i in range(len(sel) - 1)]

    def report(self) -> None:
        log.info("ELAPSED TIMES %s" % str(self.elapsed_times()))
        log.info("ITERATION TIMES %s" % str(self.iteration_times()))

    def remote_notify_ready_to_start(self) -> None:
        log.info("Server ready to start!")
        self.ready_to_start = True

    def remote_is_ready_to_start(self) -> Tuple[int, bool]:
        return (self.worker_id, self.ready_to_start)


class TimeoutClient:
    """Buffers this worker's events; ``flush`` exchanges them with all peers."""

    def __init__(self, backend: Backend, server: TimeoutServer):
        self.backend = backend
        self.server = server
        self.worker_id = backend.ctx.rank
        self._pending: List[Tuple[str, int, float]] = []
        self._last_flush_iter = -1

    def ready_to_start(self) -> bool:
        return self.server.ready_to_start

    def broadcast_parameters_updated(self, step: int) -> None:
        self.server.remote_parameters_updated(step)

    def broadcast_worker_dequeued_token(self, iteration: int, t: Optional[float] = None) -> None:
        self._pending.append(("d", iteration, time.time() if t is None else t))

    def broadcast_worker_finished_computing_gradients(self, iteration: int, t: Optional[float] = None) -> None:
        self._pending.append(("f", iteration, time.time() if t is None else t))
        # Same reporting cadence as the reference (:68): every 50th iteration and at
        # the end of the tracked window.  The exchange is collective, which is safe in
        # cdf mode because the full barrier keeps iteration counters in lockstep.
        if iteration > ITERATION_START_TRACKING and (
                iteration % REPORT_EVERY == 0 or iteration == ITERATION_END_TRACKING):
            self.flush()
            if self.worker_id == 0:
                self.server.report()

    def flush(self) -> None:
        mine, self._pending = self._pending, []
        for wid, events in enumerate(self.backend.all_gather_object(mine)):
            for kind, it, t in sorted(events, key=lambda e: (e[1], e[0])):
                if kind == "d":
                    self.server.remote_worker_dequeued_token(wid, it, t)
                else:
                    self.server.remote_worker_finished_computing_gradients(wid, it, t)


def launch_manager(backend: Backend, flags) -> Tuple[TimeoutClient, TimeoutServer]:
    """Reference ``launch_manager`` (timeout_manager.py:198-211): returns once every
    worker is up.  Even when timing is unused this is the start-up barrier."""
    n = backend.ctx.world_size
    k = flags.num_replicas_to_aggregate if flags.num_replicas_to_aggregate > 0 else n
    server = TimeoutServer(backend.ctx.rank, n, k)
    log.info("Worker %d: starting status server..." % backend.ctx.rank)
    client = TimeoutClient(backend, server)
    backend.barrier()
    server.remote_notify_ready_to_start()
    log.info("Num servers ready: %d vs %d" % (n, n))
    return client, server

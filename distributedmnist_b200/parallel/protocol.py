"""K-of-N commit protocol: the executable specification of the device-side arrival bitmap.

What it replaces (reference, SURVEY §2.3 mode A / §2.5 X1-X8): per-variable
``ConditionalAccumulator``s on the parameter server that accept a gradient iff
``local_step >= accumulator_step`` and release the mean once ``>= K`` gradients
have arrived (sync_replicas_optimizer_modified.py:57-90), plus token queues that
carry the new global step back to the workers (:394-398).

The protocol, identical on every implementation (this host model, the
c10d-store version used by the CPU/gloo backend, and ``csrc/fused_sync.cu``):

* there is one *arrival bitmap* and one *commit word* per global step;
* a replica arriving with ``local_step == global_step`` ORs its bit into the
  bitmap; whoever first observes ``popcount >= K`` publishes the frozen bitmap as
  the step's commit mask (compare-and-swap, exactly one winner);
* every replica then reads the same mask: contributors are the set bits, the
  divisor is ``popcount(mask)`` (mean over the number actually accumulated,
  reference :72-81), replicas not in the mask are *late* -- their gradient is
  discarded but they still receive the update, so replicas never diverge;
* a replica arriving with ``local_step < global_step`` is *stale* (reference
  :59-62): dropped immediately, it fast-forwards to the current step.

Because the mask is a single word for the whole flat arena, the contributor set
is uniform across variables (the reference's per-variable accumulators allow
it to differ; SURVEY §2.3 note).
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass
from typing import Dict, Optional


@dataclass
class Decision:
    step: int            # the global step this decision belongs to
    accepted: bool       # this replica's gradient is part of the mean
    stale: bool          # arrived with local_step < global_step (dropped before arrival)
    mask: int            # committed contributor bitmap of `step`
    count: int           # popcount(mask) == divisor
    global_step: int     # global step after the commit (what the replica moves on to)


def popcount(x: int) -> int:
    return bin(x).count("1")


class CommitBoard:
    """In-process, thread-safe model (one thread per replica in tests)."""

    def __init__(self, num_replicas: int, replicas_to_aggregate: int):
        assert 1 <= replicas_to_aggregate <= num_replicas <= 32
        self.n = num_replicas
        self.k = replicas_to_aggregate
        self._cv = threading.Condition()
        self._bitmap: Dict[int, int] = {}
        self._commit: Dict[int, int] = {}
        self._global_step = 0

    @property
    def global_step(self) -> int:
        with self._cv:
            return self._global_step

    def commit_mask(self, step: int) -> Optional[int]:
        with self._cv:
            return self._commit.get(step)

    def arrive(self, rank: int, local_step: int, timeout: Optional[float] = None,
               deadline_passed: Optional[bool] = None) -> Decision:
        """Replica ``rank`` finished the gradient it computed from step ``local_step`` weights.

        ``deadline_passed`` switches to interval semantics (mode C): the commit
        happens when the caller says the interval elapsed and >= 1 gradient is
        present, instead of when ``popcount >= K``.
        """
        bit = 1 << rank
        with self._cv:
            if local_step < self._global_step:
                m = self._commit.get(local_step, 0)
                return Decision(local_step, False, True, m, popcount(m), self._global_step)
            if local_step > self._global_step:
                raise RuntimeError("replica %d is ahead of the global step (%d > %d)"
                                   % (rank, local_step, self._global_step))
            bm = self._bitmap.get(local_step, 0) | bit
            self._bitmap[local_step] = bm
            ready = popcount(bm) >= self.k if deadline_passed is None else (deadline_passed and bm != 0)
            if ready and local_step not in self._commit:
                self._commit[local_step] = bm
                self._global_step = local_step + 1
                self._cv.notify_all()
            end = None if timeout is None else time.monotonic() + timeout
            while local_step not in self._commit:
                remaining = None if end is None else end - time.monotonic()
                if remaining is not None and remaining <= 0:
                    raise TimeoutError("step %d never committed (bitmap=%#x, K=%d)"
                                       % (local_step, self._bitmap.get(local_step, 0), self.k))
                self._cv.wait(remaining)
            m = self._commit[local_step]
            return Decision(local_step, bool(m & bit), False, m, popcount(m), self._global_step)


class StoreCommitBoard:
    """Same protocol over a c10d ``Store`` (TCPStore): multi-process CPU implementation.

    ``store.add`` is the atomic OR (each replica adds its own bit exactly once
    per step, so the sum equals the OR) and ``store.compare_set`` is the CAS that
    publishes the commit word -- the same two primitives the device kernel uses
    (``atom.or.sys`` / ``atom.cas.sys`` on the chief's control block).
    """

    def __init__(self, store, num_replicas: int, replicas_to_aggregate: int, prefix: str = "commit_board"):
        assert 1 <= replicas_to_aggregate <= num_replicas <= 32
        self.store = store
        self.n = num_replicas
        self.k = replicas_to_aggregate
        self.prefix = prefix

    def _k_arr(self, step: int) -> str:
        return "%s/arr/%d" % (self.prefix, step)

    def _k_commit(self, step: int) -> str:
        return "%s/commit/%d" % (self.prefix, step)

    def commit_mask(self, step: int) -> Optional[int]:
        key = self._k_commit(step)
        if not self.store.check([key]):
            return None
        return int(self.store.get(key).decode())

    def arrive(self, rank: int, local_step: int, timeout: Optional[float] = None,
               deadline_passed: Optional[bool] = None) -> Decision:
        bit = 1 << rank
        ckey = self._k_commit(local_step)
        if self.store.check([ckey]):
            m = int(self.store.get(ckey).decode())
            return Decision(local_step, False, True, m, popcount(m), local_step + 1)
        bm = int(self.store.add(self._k_arr(local_step), bit))
        ready = popcount(bm) >= self.k if deadline_passed is None else (deadline_passed and bm != 0)
        if ready:
            self.store.compare_set(ckey, "", str(bm))
        if timeout is not None:
            import datetime
            self.store.wait([ckey], datetime.timedelta(seconds=timeout))
        else:
            self.store.wait([ckey])
        m = int(self.store.get(ckey).decode())
        return Decision(local_step, bool(m & bit), False, m, popcount(m), local_step + 1)

    def gc(self, step: int) -> None:
        """Drop keys of steps older than ``step`` (best effort; bounded store growth)."""
        for key in (self._k_arr(step), self._k_commit(step)):
            try:
                self.store.delete_key(key)
            except Exception:
                pass

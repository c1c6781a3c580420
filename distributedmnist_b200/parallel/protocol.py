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


class BucketedExchange:
    """Executable specification of the bucketed, overlapped aggregation of ``csrc/fused_sync.cu`` (K == N).

    Threads play the replicas, numpy arrays the symmetric arenas, integer flag words (value = epoch + 1, monotonic) the
    control block.  ``early(rank)`` / ``late(rank)`` do exactly what ``fused_sync_early_kernel`` /
    ``fused_sync_late_kernel`` do, in the same order:

    * early: arrive_e all-to-all -> reduce my shard of the early bucket over every replica's gradient arena -> SGD ->
      store the new weights into EVERY replica's parameter arena -> done_e flags.  Does not wait for the peers' pushes.
    * late: arrive all-to-all -> one-shot: sum all contributions of the late ranges in rank order, update the local
      parameters only -> done flags (my reads of the peers' gradient arenas are complete) -> wait done_e of every peer
      (their early pushes have landed here) -> wait done of every peer (nobody still reads MY gradient arena, which the
      next step overwrites) -> epoch + 1.

    The model checks the two hazards the waits exist for: a reader must never see a gradient arena of another step
    (``grad_step``), and nobody may write a parameter range its owner still reads in the same step (``busy_early``).
    """

    def __init__(self, n: int, numel: int, early_range, lr: float = 0.1):
        import numpy as np
        self.np, self.n, self.numel, self.lr = np, n, numel, lr
        self.e0, self.e1 = early_range
        rng = np.random.default_rng(0)
        w0 = rng.standard_normal(numel).astype(np.float32)
        self.params = [w0.copy() for _ in range(n)]
        self.grads = [np.zeros(numel, np.float32) for _ in range(n)]
        self.grad_step = [-1] * n              # which step's gradients arena r currently holds
        self.busy_early = [False] * n          # replica r's compute still reads its early-range parameters
        self.early_from = [[-1] * n for _ in range(n)]   # [r][q]: step of the last early-bucket shard replica q pushed into r
        self.epoch = [0] * n
        self.arrive_e = [[0] * n for _ in range(n)]   # [owner][peer]
        self.done_e = [[0] * n for _ in range(n)]
        self.arrive = [[0] * n for _ in range(n)]
        self.done = [[0] * n for _ in range(n)]
        self._cv = threading.Condition()

    def _post(self, table, me: int, value: int) -> None:
        with self._cv:
            for q in range(self.n):
                table[q][me] = value
            self._cv.notify_all()

    def _wait(self, row, value: int, timeout: float = 10.0) -> None:
        end = time.monotonic() + timeout
        with self._cv:
            while not all(v >= value for v in row):
                left = end - time.monotonic()
                if left <= 0:
                    raise TimeoutError("flag wait timed out: %s < %d" % (row, value))
                self._cv.wait(left)

    def write_grads(self, r: int, g) -> None:
        """The backward pass of replica r overwrites its gradient arena for its current step."""
        self.grads[r][:] = g
        self.grad_step[r] = self.epoch[r]

    def early(self, r: int) -> None:
        np, n, ep = self.np, self.n, self.epoch[r]
        self._post(self.arrive_e, r, ep + 1)
        self._wait(self.arrive_e[r], ep + 1)
        size = self.e1 - self.e0
        shard = (size + n - 1) // n
        b, e = self.e0 + r * shard, min(self.e0 + (r + 1) * shard, self.e1)
        acc = np.zeros(max(e - b, 0), np.float32)
        for q in range(n):                                    # rank order: the same sum on every implementation
            assert self.grad_step[q] == ep, "replica %d read replica %d's gradients of step %d in step %d" % (r, q, self.grad_step[q], ep)
            acc += self.grads[q][b:e]
        nw = self.params[r][b:e] - np.float32(self.lr / n) * acc
        for q in range(n):
            assert not self.busy_early[q], "push into replica %d's early range while its compute still reads it" % q
            self.params[q][b:e] = nw
            self.early_from[q][r] = ep
        self._post(self.done_e, r, ep + 1)

    def late(self, r: int) -> None:
        np, n, ep = self.np, self.n, self.epoch[r]
        self._post(self.arrive, r, ep + 1)
        self._wait(self.arrive[r], ep + 1)
        for (b, e) in ((0, self.e0), (self.e1, self.numel)):
            acc = np.zeros(e - b, np.float32)
            for q in range(n):
                assert self.grad_step[q] == ep, "replica %d read replica %d's late gradients of step %d in step %d" % (r, q, self.grad_step[q], ep)
                acc += self.grads[q][b:e]
            self.params[r][b:e] -= np.float32(self.lr / n) * acc
        self._post(self.done, r, ep + 1)
        self._wait(self.done_e[r], ep + 1)
        # (the bf16 shadow refresh and the next forward pass read the early range from here on: every shard must be this step's)
        assert all(v == ep for v in self.early_from[r]), "replica %d reads early-bucket shards of steps %s in step %d" % (r, self.early_from[r], ep)
        self._wait(self.done[r], ep + 1)
        self.epoch[r] = ep + 1


# ----------------------------------------------------------------------------------------------------------------------
# Interval mode (mode C) -- executable model of csrc/fused_interval.cu
# ----------------------------------------------------------------------------------------------------------------------
class IntervalBoard:
    """Word-level model of the device-side interval protocol (one thread per replica in tests; every method body is one
    "memory operation" of the kernels and takes the lock, so threads interleave at exactly the granularity at which the
    kernels' system-scope loads / stores / CAS interleave on the GPUs under sequentially consistent fences).

    State per replica (the words of ``SyncCtrl``): ``epoch``, ``iv_state`` = (step tag, count), ``iv_busy``, the local copy of
    the commit ring, ``done`` flags; the chief's authoritative ``commit`` ring.  The accumulator is modelled as a float per
    replica plus an ``in_flight`` marker the committer must never observe while it reads (the property the Dekker pair
    ``busy := step+1; fence; read commit``  ||  ``write commit; fence; read busy`` provides).
    """

    def __init__(self, num_replicas: int, lr: float = 1.0):
        self.n = num_replicas
        self.lr = lr
        self._lock = threading.Lock()
        self.epoch = [0] * self.n
        self.state = [(1, 0)] * self.n             # (step + 1 tag, count)
        self.busy = [0] * self.n
        self.commit_local = [dict() for _ in range(self.n)]   # step -> (mask, committer)
        self.done = [[0] * self.n for _ in range(self.n)]     # done[q][p]: p's pushes for step s landed on q (value s+1)
        self.commit: Dict[int, tuple] = {}         # chief: step -> (mask, committer)
        self.acc = [0.0] * self.n
        self.in_flight = [False] * self.n
        self.weights = [0.0] * self.n
        self.ticks = []                            # (step, committer, mask, total)
        self.dropped = [0] * self.n
        self.accumulated = [0] * self.n
        self.lost = [0] * self.n                   # accumulated for a step whose mask was frozen a moment earlier (count was 0)

    # -- iv_adopt ------------------------------------------------------------------------------------------------------------
    def adopt(self, r: int) -> int:
        with self._lock:
            e, adopted = self.epoch[r], 0
            while e in self.commit_local[r]:
                mask, committer = self.commit_local[r][e]
                if self.done[r][committer] < e + 1:
                    break                          # its pushes are still landing
                tag, c = self.state[r]
                if tag == e + 1 and c > 0 and not (mask >> r) & 1:
                    self.lost[r] += c              # arrived after the mask was read: discarded with the step (reference: stale)
                e += 1
                adopted += 1
            if adopted:
                self.epoch[r] = e
                self.state[r] = (e + 1, 0)
            return adopted

    # -- iv_gate: busy := step+1; fence; commit word present? ------------------------------------------------------------------
    def gate_set_busy(self, r: int) -> int:
        with self._lock:
            self.busy[r] = self.epoch[r] + 1
            return self.epoch[r]

    def gate_check(self, r: int, e: int) -> bool:
        with self._lock:
            go = e not in self.commit_local[r]
            if not go:
                self.busy[r] = 0
                self.dropped[r] += 1
            return go

    # -- iv_accumulate (two operations so a reader could observe it half done) ---------------------------------------------------
    def accumulate_begin(self, r: int) -> None:
        with self._lock:
            self.in_flight[r] = True

    def accumulate_end(self, r: int, g: float) -> None:
        with self._lock:
            first = self.state[r][1] == 0
            self.acc[r] = g if first else self.acc[r] + g
            self.in_flight[r] = False

    # -- iv_close ---------------------------------------------------------------------------------------------------------------------
    def close(self, r: int, e: int) -> None:
        with self._lock:
            tag, c = self.state[r]
            assert tag == e + 1
            self.state[r] = (tag, c + 1)
            self.busy[r] = 0
            self.accumulated[r] += 1

    def try_commit(self, r: int, e: int) -> Optional[int]:
        """Deadline passed: read every replica's state word, CAS the chief's commit word.  Returns the mask if this replica won."""
        with self._lock:
            if e in self.commit:
                return None
        mask = 0
        for q in range(self.n):
            with self._lock:                        # one remote load per replica
                tag, c = self.state[q]
            if tag == e + 1 and c > 0:
                mask |= 1 << q
        with self._lock:                            # the CAS
            if e in self.commit:
                return None
            self.commit[e] = (mask, r)
        for q in range(self.n):
            with self._lock:                        # broadcast of the commit word, one store per replica
                self.commit_local[q][e] = (mask, r)
        return mask

    def collect(self, r: int, e: int, mask: int) -> int:
        """Committer: wait until no contributor is mid-accumulate for this step, then read the final counts."""
        total = 0
        for q in range(self.n):
            if not (mask >> q) & 1:
                continue
            while True:
                with self._lock:
                    if self.busy[q] != e + 1:
                        break
                time.sleep(0)
            with self._lock:
                tag, c = self.state[q]
                assert tag == e + 1, "a contributor moved on before its accumulator was read"
                total += c
        return total

    # -- iv_apply ----------------------------------------------------------------------------------------------------------------------
    def apply(self, r: int, e: int, mask: int, total: int) -> None:
        s = 0.0
        for q in range(self.n):
            if (mask >> q) & 1:
                with self._lock:
                    assert not self.in_flight[q], "committer read an accumulator while its owner was adding to it"
                    s += self.acc[q]
        with self._lock:
            new_w = self.weights[r] - self.lr * s / total
        for q in range(self.n):
            with self._lock:
                self.weights[q] = new_w
        with self._lock:
            self.ticks.append((e, r, mask, total))
        for q in range(self.n):
            with self._lock:
                self.done[q][r] = e + 1

    # -- one local iteration of replica r (what the step graph does) -----------------------------------------------------------------------
    def iteration(self, r: int, g: float, deadline_passed: bool, compute=lambda: None) -> None:
        self.adopt(r)
        compute()
        e = self.gate_set_busy(r)
        if not self.gate_check(r, e):
            return
        self.accumulate_begin(r)
        self.accumulate_end(r, g)
        self.close(r, e)
        if deadline_passed:
            mask = self.try_commit(r, e)
            if mask is not None:
                total = self.collect(r, e, mask)
                self.apply(r, e, mask, total)

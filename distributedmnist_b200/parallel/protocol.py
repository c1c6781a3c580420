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


# ----------------------------------------------------------------------------------------------------------------------
# K == N, bucketed (default path) -- executable model of csrc/fused_bucket.cu, explorable schedule by schedule
# ----------------------------------------------------------------------------------------------------------------------
class ProtocolHazard(AssertionError):
    """A replica consumed data of another step, or overwrote data its consumer had not read yet."""


class BucketV2Model:
    """Word-level model of ``bucket_early_kernel`` + ``bucket_late_ll_kernel`` (csrc/fused_bucket.cu).

    Every replica is a generator; each ``yield`` is one memory operation visible to the other replicas (a flag store, the
    loads of one element from all peers, the stores of one element to one peer, one LL line landing at ONE destination -- a
    multicast store reaches its destinations at different times) or a blocking wait (``yield predicate``).  A scheduler
    picks which runnable replica performs its next operation, so a test can enumerate EVERY interleaving of a small instance
    (:func:`explore_schedules`) or sample large ones.  The memory cells carry (step, kind) tags next to their values and every
    read checks them: the model fails with :class:`ProtocolHazard` exactly when the kernels would consume torn / stale data.

    What the model is for -- the three places where the kernels rely on an ordering argument instead of a flag:

    * the LL inbox has no "consumed" handshake: slot ``[parity][sender]`` is rewritten two steps later.  Safe because a sender
      can only reach step s + 2 after it received every peer's step s + 1 lines, which a peer sends only after its step-s
      kernel (and therefore its step-s polls) completed.  ``parities=1`` removes the double buffer and must fail.
    * the in-place reduction of the bf16 early bucket: rank r overwrites shard r of EVERY replica's buffer with the sum.  A
      replica may read its whole buffer only after ``done_e`` of all ranks, and may start the next backward pass (which
      overwrites the buffer with new gradients) only then too.  ``wait_done_e=False`` must fail.
    * gradients are read from peers only after ``arrive_e`` of that peer (``wait_arrive_e=False`` must fail).
    """

    def __init__(self, n: int, early_len: int, late_len: int, steps: int, parities: int = 2, wait_arrive_e: bool = True,
                 wait_done_e: bool = True, lr: float = 0.5):
        self.n, self.early_len, self.late_len, self.steps, self.parities, self.lr = n, early_len, late_len, steps, parities, lr
        self.wait_arrive_e, self.wait_done_e = wait_arrive_e, wait_done_e
        self.w_early = [[0.0] * early_len for _ in range(n)]
        self.w_late = [[0.0] * late_len for _ in range(n)]
        self.g16 = [[(-1, "none", 0.0)] * early_len for _ in range(n)]            # (step, "raw" | "sum", value)
        self.inbox = [[[[(0, 0.0)] * late_len for _ in range(n)] for _ in range(parities)] for _ in range(n)]   # [dst][parity][src][i] = (tag, v)
        self.consumed = [[[[0] * late_len for _ in range(n)] for _ in range(parities)] for _ in range(n)]      # highest tag dst has read
        self.arrive_e = [[0] * n for _ in range(n)]      # [owner][peer]
        self.done_e = [[0] * n for _ in range(n)]
        self.epoch = [0] * n

    @staticmethod
    def grad(r: int, step: int, i: int, late: bool) -> float:
        return float((r + 1) * (step + 1) + 3 * i + (100 if late else 0))         # small integers: sums are exact

    def expected(self, late: bool, i: int) -> float:
        return -self.lr / self.n * sum(self.grad(r, s, i, late) for r in range(self.n) for s in range(self.steps))

    # -- one replica ------------------------------------------------------------------------------------------------------
    def replica(self, r: int):
        n = self.n
        shard = (self.early_len + n - 1) // n
        b, e = min(r * shard, self.early_len), min((r + 1) * shard, self.early_len)
        for _ in range(self.steps):
            ep = self.epoch[r]
            # fc1_wgrad epilogue: my bf16 gradient of this step (local stores, but peers will read / overwrite these cells)
            for i in range(self.early_len):
                st, kind, _v = self.g16[r][i]
                if st == ep - 1 and kind == "raw" and ep > 0:
                    raise ProtocolHazard("replica %d starts step %d before shard of element %d was reduced" % (r, ep, i))
                self.g16[r][i] = (ep, "raw", self.grad(r, ep, i, False))
            if self.early_len:
                yield None
                # ---- bucket_early_kernel: exchange (operations on my own memory are not scheduling points of their own) ----------
                for q in range(n):
                    self.arrive_e[q][r] = ep + 1
                    if q != r:
                        yield None
                if self.wait_arrive_e:
                    yield (lambda: all(v >= ep + 1 for v in self.arrive_e[r]))
                for i in range(b, e):
                    total = 0.0
                    for q in range(n):                   # multimem.ld_reduce / N peer loads
                        st, kind, v = self.g16[q][i]
                        if (st, kind) != (ep, "raw"):
                            raise ProtocolHazard("replica %d reduced element %d of replica %d holding %s of step %d in step %d"
                                                 % (r, i, q, kind, st, ep))
                        total += v
                    self.g16[r][i] = (ep, "sum", total)  # (the kernel writes its own copy through the local path)
                    yield None
                    for q in range(n):                   # multimem.st: lands per destination
                        if q != r:
                            self.g16[q][i] = (ep, "sum", total)
                            yield None
                for q in range(n):
                    self.done_e[q][r] = ep + 1
                    if q != r:
                        yield None
                # ---- apply --------------------------------------------------------------------------------------------------
                if self.wait_done_e:
                    yield (lambda: all(v >= ep + 1 for v in self.done_e[r]))
                for i in range(self.early_len):
                    st, kind, v = self.g16[r][i]
                    if (st, kind) != (ep, "sum"):
                        raise ProtocolHazard("replica %d applied element %d holding %s of step %d in step %d" % (r, i, kind, st, ep))
                    self.w_early[r][i] -= self.lr / n * v
                yield None
            # ---- bucket_late_ll_kernel: push lines, poll, rank-ordered sum, SGD, epoch + 1 -----------------------------------------
            tag, par = ep + 1, ep % self.parities
            for q in range(n):
                if q == r:
                    continue
                for i in range(self.late_len):
                    old_tag = self.inbox[q][par][r][i][0]
                    if self.consumed[q][par][r][i] < old_tag:
                        raise ProtocolHazard("replica %d overwrote line %d (tag %d) in replica %d's inbox before it was read"
                                             % (r, i, old_tag, q))
                    self.inbox[q][par][r][i] = (tag, self.grad(r, ep, i, True))
                yield None
            for i in range(self.late_len):
                yield (lambda i=i: all(self.inbox[r][par][c][i][0] >= tag for c in range(n) if c != r))
                acc = 0.0
                for c in range(n):
                    if c == r:
                        acc += self.grad(r, ep, i, True)
                        continue
                    t, v = self.inbox[r][par][c][i]
                    if t != tag:
                        raise ProtocolHazard("replica %d polled line %d of replica %d: tag %d, wanted %d (overwritten)" % (r, i, c, t, tag))
                    self.consumed[r][par][c][i] = t
                    acc += v
                self.w_late[r][i] -= self.lr / n * acc
            self.epoch[r] = ep + 1

    # -- scheduler ----------------------------------------------------------------------------------------------------------
    def run(self, choose) -> None:
        """``choose(runnable_ranks) -> rank``.  Raises ``ProtocolHazard`` on a violated invariant or a deadlock."""
        gens = {r: self.replica(r) for r in range(self.n)}
        waiting = {}
        for r in list(gens):
            self._advance(r, gens, waiting)
        while gens:
            runnable = [r for r in sorted(gens) if r not in waiting or waiting[r]()]
            if not runnable:
                raise ProtocolHazard("deadlock: replicas %s wait forever (epochs %s)" % (sorted(gens), self.epoch))
            r = choose(runnable)
            waiting.pop(r, None)
            self._advance(r, gens, waiting)
        for r in range(self.n):
            for i in range(self.early_len):
                if self.w_early[r][i] != self.expected(False, i):
                    raise ProtocolHazard("replica %d early weight %d = %r, expected %r" % (r, i, self.w_early[r][i], self.expected(False, i)))
            for i in range(self.late_len):
                if self.w_late[r][i] != self.expected(True, i):
                    raise ProtocolHazard("replica %d late weight %d = %r, expected %r" % (r, i, self.w_late[r][i], self.expected(True, i)))

    @staticmethod
    def _advance(r, gens, waiting) -> None:
        try:
            w = next(gens[r])
        except StopIteration:
            del gens[r]
            return
        if w is not None:
            waiting[r] = w


def explore_schedules(make_model, max_runs: int = 200000) -> int:
    """Run ``make_model().run`` under EVERY schedule (depth-first over the scheduler's choice points, replaying prefixes).
    Returns the number of complete schedules; raises the first ``ProtocolHazard``."""
    stack, runs = [[]], 0
    while stack:
        prefix = stack.pop()
        trace = []

        def choose(runnable, prefix=prefix, trace=trace):
            i = len(trace)
            c = prefix[i] if i < len(prefix) else 0
            trace.append((c, len(runnable)))
            return runnable[c]

        make_model().run(choose)
        runs += 1
        if runs > max_runs:
            raise RuntimeError("more than %d schedules" % max_runs)
        for i in range(len(prefix), len(trace)):
            for alt in range(1, trace[i][1]):
                stack.append([t[0] for t in trace[:i]] + [alt])
    return runs


# ----------------------------------------------------------------------------------------------------------------------
# K < N (backup workers) -- word-level model of decide<KOFN> + fused_sync_sgd_kernel (csrc/fused_sync.cu)
# ----------------------------------------------------------------------------------------------------------------------
class KofNModel:
    """Every replica is a generator over the memory operations of its training loop (same scheduler as
    :class:`BucketV2Model`): read the parameter arena element by element (forward/backward), write the gradient arena, then
    the kernel -- local commit word, ``atomicOr`` into the chief's bitmap slot, CAS on the chief's commit word by whoever
    sees ``popcount >= K``, ``last_in_mask`` / ``global_step`` / broadcast of the commit word / wipe of the bitmap slot half
    a ring ahead, reduction of the owned shard over the masked contributors, push of the new weights to ALL replicas,
    ``done`` flags, the waits, fast-forward of late and stale replicas.

    Invariants checked while it runs (``ProtocolHazard`` otherwise):

    * a gradient that is ACCEPTED for step s was computed on exactly the weights of step s (a replica whose arena is being
      pushed into while it computes must always turn out late);
    * a team member only reads gradient arenas holding step-s gradients, and only pushes version s + 1 over version s;
    * the committed mask contains only replicas that really arrived for that step (no phantom bit from a recycled bitmap slot);
    * at the end every replica holds the same weights: the serial application of the committed means.

    The bitmap slots carry no tag (one ``atomicOr`` per arrival instead of a CAS loop); the committer of step s wipes the slot
    of step s + ring/2.  A replica that stalls for more than ring/2 global steps BETWEEN its commit-word read and its
    ``atomicOr`` therefore leaves a phantom bit -- with the shipped ring of 64 that is a stall of > 32 steps (~4 ms) between
    two instructions of one warp; the model shows it with ``ring=2`` (``tests/test_protocol.py``).  ``tagged=True`` models the
    designed fix (arrival word = (tag, bits), CAS loop, no wipe): no phantom bit under the same schedules.
    """

    def __init__(self, n: int, k: int, steps: int, ring: int = 8, numel: Optional[int] = None, lr: float = 0.5,
                 tagged: bool = False):
        assert 1 <= k <= n and ring >= 2 and ring % 2 == 0
        self.n, self.k, self.steps, self.ring, self.lr = n, k, steps, ring, lr
        self.tagged = tagged                                                # NEXT_STEPS item 7: (tag, bits) arrival words, CAS loop, no wipe
        self.bitmap_tag = [0] * ring
        self.numel = numel if numel is not None else n
        self.full = (1 << n) - 1
        self.params = [[(0, 0.0)] * self.numel for _ in range(n)]          # (version = number of updates applied, value)
        self.grads = [(-1, False, None)] * n                                # (step, computed on consistent step weights, values)
        self.epoch = [0] * n
        self.arrived = [-1] * n                                             # last step a replica ORed its bit for
        self.commit_local = [[(0, 0)] * ring for _ in range(n)]            # (tag, mask)
        self.done = [[0] * n for _ in range(n)]                             # [owner][peer]
        self.bitmap = [0] * ring                                            # chief
        self.commit = [(0, 0)] * ring                                       # chief
        self.last_in_mask = [0] * n                                         # chief
        self.global_step = 0                                                # chief
        self.log: Dict[int, int] = {}                                       # step -> committed mask
        self.at = [""] * n                                                  # program point labels for targeted schedulers
        self.accepted = [0] * n
        self.dropped = [0] * n

    @staticmethod
    def grad(r: int, step: int, i: int) -> float:
        return float((r + 1) + 7 * step + 3 * i)

    def replica(self, r: int):
        n, ring, bit = self.n, self.ring, 1 << r
        while self.epoch[r] < self.steps:
            ep = self.epoch[r]
            # ---- forward / backward on my parameter arena (peers may push into it concurrently) -------------------------------
            consistent = True
            for i in range(self.numel):
                consistent = consistent and self.params[r][i][0] == ep
                yield None
            self.grads[r] = (ep, consistent, [self.grad(r, ep, i) for i in range(self.numel)])
            yield None
            # ---- decide<KOFN> ----------------------------------------------------------------------------------------------
            slot, want = ep % ring, ep + 1
            cw = self.commit_local[r][slot]
            self.at[r] = "after_commit_word_read"
            yield None
            self.at[r] = ""
            if cw[0] < want:
                now = 0
                if not self.tagged:
                    self.bitmap[slot] |= bit                                # atomicOr_system
                    self.arrived[r] = ep
                    now = self.bitmap[slot] & self.full
                    yield None
                else:
                    while True:                                             # CAS loop on the (tag, bits) arrival word
                        seen = (self.bitmap_tag[slot], self.bitmap[slot])
                        yield None
                        if seen[0] > want:
                            break                                           # the slot already belongs to a later step: I am stale
                        new = (want, bit) if seen[0] < want else (want, seen[1] | bit)
                        if (self.bitmap_tag[slot], self.bitmap[slot]) == seen:
                            self.bitmap_tag[slot], self.bitmap[slot] = new
                            self.arrived[r] = ep
                            now = new[1] & self.full
                            yield None
                            break
                        yield None
                if popcount(now) >= self.k:
                    old = self.commit[slot]
                    yield None
                    if old[0] < want:
                        won = self.commit[slot] == old                      # atomicCAS_system
                        if won:
                            for q in range(n):
                                if (now >> q) & 1 and self.arrived[q] != ep:
                                    raise ProtocolHazard("step %d committed with a phantom bit of replica %d (it last arrived for step %d)"
                                                         % (ep, q, self.arrived[q]))
                            self.commit[slot] = (want, now)
                            self.log[ep] = now
                        yield None
                        if won:
                            for q in range(n):
                                if (now >> q) & 1:
                                    self.last_in_mask[q] = ep + 1
                            yield None
                            self.global_step = max(self.global_step, ep + 1)
                            yield None
                            for q in range(n):
                                self.commit_local[q][slot] = (want, now)
                                if q != r:
                                    yield None
                            if not self.tagged:
                                self.bitmap[(slot + ring // 2) % ring] = 0
                                yield None
                yield (lambda: self.commit_local[r][slot][0] >= want)
                cw = self.commit_local[r][slot]
            if cw[0] == want:
                mask, late = cw[1], not (cw[1] & bit)
            else:
                mask, late = 0, True                                        # committed a ring lap ago
            target = ep + 1
            if late:
                g = self.global_step
                yield None
                target = max(g, ep + 1)
            count = popcount(mask)
            if not late:
                step_g, ok, _vals = self.grads[r]
                if not ok:
                    raise ProtocolHazard("replica %d: gradient accepted for step %d was computed on weights of mixed steps" % (r, ep))
                # ---- reduce my shard over the contributors, SGD, push to every replica -----------------------------------------
                members = [q for q in range(n) if (mask >> q) & 1]
                my_idx = members.index(r)
                shard = (self.numel + count - 1) // count
                for i in range(my_idx * shard, min((my_idx + 1) * shard, self.numel)):
                    acc = 0.0
                    for q in members:
                        gs, _ok, vals = self.grads[q]
                        if gs != ep:
                            raise ProtocolHazard("replica %d read replica %d's gradient arena of step %d in step %d" % (r, q, gs, ep))
                        acc += vals[i]
                    yield None
                    ver, w = self.params[r][i]
                    if ver != ep:
                        raise ProtocolHazard("replica %d updates element %d at version %d in step %d" % (r, i, ver, ep))
                    new = (ep + 1, w - self.lr / count * acc)
                    for q in range(n):
                        if q != r and self.params[q][i][0] != ep:
                            raise ProtocolHazard("replica %d pushes step %d over version %d of element %d at replica %d"
                                                 % (r, ep, self.params[q][i][0], i, q))
                        self.params[q][i] = new
                        if q != r:
                            yield None
                for q in range(n):
                    self.done[q][r] = ep + 1
                    if q != r:
                        yield None
                self.accepted[r] += 1
            else:
                self.dropped[r] += 1
            # ---- wait until every team member's shard has landed in MY arena ------------------------------------------------------
            for q in range(n):
                need = (ep + 1) if (mask >> q) & 1 else 0
                if late:
                    need = self.last_in_mask[q]                             # fast-forward: everything committed so far
                    yield None
                yield (lambda q=q, need=need: self.done[r][q] >= need)
            self.epoch[r] = target

    def run(self, choose) -> None:
        gens = {r: self.replica(r) for r in range(self.n)}
        waiting = {}
        for r in list(gens):
            BucketV2Model._advance(r, gens, waiting)
        while gens:
            runnable = [r for r in sorted(gens) if r not in waiting or waiting[r]()]
            if not runnable:
                raise ProtocolHazard("deadlock: replicas %s wait forever (epochs %s, global step %d)" % (sorted(gens), self.epoch, self.global_step))
            r = choose(runnable)
            waiting.pop(r, None)
            BucketV2Model._advance(r, gens, waiting)
        # serial replay of the committed means
        assert sorted(self.log) == list(range(self.steps)), "committed steps %s" % sorted(self.log)
        w = [0.0] * self.numel
        for s in range(self.steps):
            members = [q for q in range(self.n) if (self.log[s] >> q) & 1]
            if len(members) < self.k:
                raise ProtocolHazard("step %d committed with %d < K contributors" % (s, len(members)))
            for i in range(self.numel):
                w[i] -= self.lr / len(members) * sum(self.grad(q, s, i) for q in members)
        for r in range(self.n):
            if self.params[r] != [(self.steps, v) for v in w]:
                raise ProtocolHazard("replica %d ends with %s, expected versions %d values %s" % (r, self.params[r], self.steps, w))

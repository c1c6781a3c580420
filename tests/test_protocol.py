"""K-of-N commit protocol semantics with a software arrival bitmap."""
import threading
import time

import pytest

from distributedmnist_b200.parallel.protocol import CommitBoard, popcount


def _run(board, delays, local_steps=None):
    n = len(delays)
    out = [None] * n

    def worker(r):
        time.sleep(delays[r])
        out[r] = board.arrive(r, 0 if local_steps is None else local_steps[r], timeout=5.0)

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def test_full_barrier_accepts_everyone():
    out = _run(CommitBoard(4, 4), [0.0, 0.01, 0.02, 0.03])
    assert all(d.accepted and d.mask == 0b1111 and d.count == 4 and d.global_step == 1 for d in out)


def test_k_of_n_masks_late_replicas_and_divisor_is_accepted_count():
    board = CommitBoard(8, 6)
    delays = [0.0] * 6 + [0.3, 0.4]         # ranks 6, 7 are stragglers
    out = _run(board, delays)
    assert all(out[r].accepted for r in range(6))
    assert not out[6].accepted and not out[7].accepted
    masks = {d.mask for d in out}
    assert masks == {0b00111111}             # every replica sees the same contributor set
    assert all(d.count == 6 for d in out)
    # late ones arrived after the commit -> reported stale, still told the new global step
    assert out[6].stale and out[7].stale and out[6].global_step == 1


def test_first_gradient_wins_when_k_is_one():
    out = _run(CommitBoard(4, 1), [0.2, 0.0, 0.2, 0.2])
    assert out[1].accepted and out[1].mask == 0b0010 and out[1].count == 1
    assert [d.accepted for d in out] == [False, True, False, False]


def test_stale_gradient_is_dropped_and_does_not_block():
    board = CommitBoard(2, 1)
    assert board.arrive(0, 0).accepted          # commits step 0 alone
    assert board.arrive(0, 1).accepted          # and step 1
    d = board.arrive(1, 0)                      # replica 1 still holds a step-0 gradient
    assert d.stale and not d.accepted and d.global_step == 2 and d.mask == 0b01
    assert board.arrive(1, 2).accepted          # caught up, contributes again


def test_ahead_of_global_step_is_a_protocol_error():
    with pytest.raises(RuntimeError):
        CommitBoard(2, 2).arrive(0, 3)


def test_timeout_when_not_enough_arrivals():
    with pytest.raises(TimeoutError):
        CommitBoard(3, 3).arrive(0, 0, timeout=0.05)


def test_interval_commit_takes_whatever_arrived():
    board = CommitBoard(4, 4)
    # before the deadline nothing commits; the first arrival after it commits the present set
    t = threading.Thread(target=lambda: board.arrive(0, 0, deadline_passed=False, timeout=5))
    t.start()
    time.sleep(0.05)
    assert board.global_step == 0
    d = board.arrive(2, 0, deadline_passed=True)
    t.join()
    assert d.mask == 0b0101 and d.count == 2 and board.global_step == 1


def test_many_steps_many_threads_consistent():
    n, k, steps = 8, 5, 50
    board = CommitBoard(n, k)
    accepted = [[False] * steps for _ in range(n)]

    def worker(r):
        s = 0
        while s < steps:
            d = board.arrive(r, s, timeout=10)
            if not d.stale:
                accepted[r][d.step] = d.accepted
            s = d.global_step
    ts = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for s in range(steps):
        m = board.commit_mask(s)
        assert popcount(m) >= k
        assert all(accepted[r][s] == bool(m >> r & 1) for r in range(n) if accepted[r][s])


def test_bucketed_overlapped_exchange_model_is_race_free_and_matches_allreduce_sgd():
    """Host model of the bucketed aggregation (csrc/fused_sync.cu): replicas drift apart by random amounts between the
    phases of a step; the flag waits must keep every reader on the right step's gradients, and the result must equal
    plain allreduce -> /N -> SGD on every replica."""
    import random

    import numpy as np

    from distributedmnist_b200.parallel.protocol import BucketedExchange
    n, numel, steps = 4, 4096, 12
    ex = BucketedExchange(n, numel, (512, 3584), lr=0.1)
    w_ref = ex.params[0].copy()
    all_g = [[np.random.default_rng(100 * s + r).standard_normal(numel).astype(np.float32) for r in range(n)] for s in range(steps)]
    errors = []

    def replica(r):
        rnd = random.Random(r)
        try:
            for s in range(steps):
                time.sleep(rnd.random() * 0.004)            # forward
                ex.busy_early[r] = True                      # fc kernels read the early-range weights ...
                time.sleep(rnd.random() * 0.002)
                ex.write_grads(r, all_g[s][r])               # backward (overwrites the arena the peers read LAST step)
                ex.busy_early[r] = False                     # ... until fc1_dgrad is done; only then the early kernel starts
                ex.early(r)
                time.sleep(rnd.random() * 0.004)            # the rest of the backward pass runs next to the early exchange
                ex.late(r)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=replica, args=(r,)) for r in range(n)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not errors, errors
    for s in range(steps):
        acc = np.zeros(numel, np.float32)
        for r in range(n):
            acc += all_g[s][r]
        w_ref = w_ref - np.float32(0.1 / n) * acc
    for r in range(n):
        assert ex.epoch[r] == steps
        assert np.array_equal(ex.params[r], ex.params[0])           # replicas bit-identical
    assert np.allclose(ex.params[0], w_ref, atol=1e-5)


def test_device_interval_protocol_model_accounts_every_gradient_exactly_once():
    """Executable model of csrc/fused_interval.cu (parallel/protocol.py::IntervalBoard): free-running replicas, random
    deadlines, one slow replica.  Every tick has exactly one committer; every computed gradient is either part of exactly one
    tick's divisor or dropped as stale; the committer never reads an accumulator that is being added to; with a constant
    gradient the weights are w0 - lr * (#ticks) * g on every replica (mean of ANY subset of equal gradients is that gradient)."""
    import random
    import threading
    import time

    from distributedmnist_b200.parallel.protocol import IntervalBoard
    n, iters, g = 4, 300, 0.5
    board = IntervalBoard(n, lr=0.1)
    errors = []

    def run(r):
        rng = random.Random(100 + r)
        try:
            for _ in range(iters // (6 if r == n - 1 else 1)):
                def compute():
                    if r == n - 1:
                        time.sleep(0.0005)            # the slow replica: several ticks per iteration
                    elif rng.random() < 0.3:
                        time.sleep(0)
                board.iteration(r, g, deadline_passed=rng.random() < 0.25, compute=compute)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=run, args=(r,)) for r in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert not errors, errors
    for r in range(n):
        board.adopt(r)
    steps = [t[0] for t in board.ticks]
    assert sorted(steps) == list(range(len(steps))) and len(steps) > 10          # one committer per step, no gaps
    assert all(total >= 1 and (mask >> committer) & 1 for (_s, committer, mask, total) in board.ticks)
    # conservation: gradients accumulated = gradients counted in some tick + gradients left in never-committed accumulators
    counted = sum(t[3] for t in board.ticks)
    leftover = sum(c for (tag, c) in board.state)
    # (+ the rare gradient that completed between the committer's read of its count (0) and the commit word reaching it)
    assert sum(board.accumulated) == counted + leftover + sum(board.lost), (sum(board.accumulated), counted, leftover, board.lost)
    assert len(set(board.epoch)) == 1 and board.epoch[0] == len(steps)
    expect = -0.1 * len(steps) * g
    assert all(abs(w - expect) < 1e-9 for w in board.weights), (board.weights, expect)
    assert any(not (mask >> (n - 1)) & 1 for (_s, _c, mask, _t) in board.ticks)   # ticks committed without the slow replica


# ---- csrc/fused_bucket.cu (default K == N path): every interleaving of small instances, sampled large ones -----------------------
def test_bucket_v2_ll_inbox_is_safe_under_every_schedule_of_two_replicas():
    from distributedmnist_b200.parallel.protocol import BucketV2Model, explore_schedules
    runs = explore_schedules(lambda: BucketV2Model(n=2, early_len=0, late_len=1, steps=4))
    assert runs > 50, runs


def test_bucket_v2_early_and_late_buckets_every_schedule_one_step_two_replicas():
    from distributedmnist_b200.parallel.protocol import BucketV2Model, explore_schedules
    runs = explore_schedules(lambda: BucketV2Model(n=2, early_len=2, late_len=1, steps=1), max_runs=2000000)
    assert runs > 1000, runs


def test_bucket_v2_sampled_schedules_up_to_eight_replicas():
    import random

    from distributedmnist_b200.parallel.protocol import BucketV2Model
    for seed in range(60):
        rng = random.Random(seed)
        n = rng.choice([2, 3, 4, 8])
        m = BucketV2Model(n=n, early_len=rng.choice([n, n + 1, 2 * n]), late_len=rng.choice([1, 3]), steps=rng.choice([3, 4, 5]))
        mode = seed % 3
        if mode == 0:
            m.run(lambda runnable: rng.choice(runnable))
        elif mode == 1:
            m.run(lambda runnable: runnable[0])              # lowest rank runs as far ahead as the protocol lets it
        else:
            m.run(lambda runnable: runnable[-1])


def test_bucket_v2_model_detects_the_hazards_the_kernels_guard_against():
    from distributedmnist_b200.parallel.protocol import BucketV2Model, ProtocolHazard, explore_schedules
    # single-buffered inbox: the fast replica's step s + 1 lines overwrite unread step-s lines
    with pytest.raises(ProtocolHazard, match="overwrote line|overwritten|deadlock"):
        explore_schedules(lambda: BucketV2Model(n=2, early_len=0, late_len=1, steps=3, parities=1))
    # no wait for the peers' "my shard is out" flags: a replica applies un-reduced elements
    with pytest.raises(ProtocolHazard, match="applied element|reduced element|before shard"):
        BucketV2Model(n=2, early_len=2, late_len=1, steps=2, wait_done_e=False).run(lambda runnable: runnable[-1])
    # no wait for the peers' "gradient written" flags: a replica reduces a peer's previous-step buffer
    with pytest.raises(ProtocolHazard, match="reduced element"):
        BucketV2Model(n=2, early_len=2, late_len=1, steps=2, wait_arrive_e=False).run(lambda runnable: runnable[0])


# ---- csrc/fused_sync.cu (K < N): word-level model, sampled schedules with stragglers ---------------------------------------------
def _kofn_run(seed: int, ring: int, stall_at_commit_word: bool, tagged: bool = False):
    import random

    from distributedmnist_b200.parallel.protocol import KofNModel
    rng = random.Random(seed)
    n = rng.choice([2, 3, 4, 8])
    k = rng.randint(1, n - 1) if stall_at_commit_word else rng.randint(1, n)
    m = KofNModel(n, k, steps=rng.choice([3, 6, 12]), ring=ring, tagged=tagged)
    victim, p = n - 1, rng.choice([0.01, 0.05, 0.3, 1.0])

    def choose(runnable):
        others = [r for r in runnable if r != victim]
        if stall_at_commit_word:      # freeze the victim between its commit-word read and its atomicOr while the others run on
            if victim in runnable and m.at[victim] == "after_commit_word_read" and others and rng.random() < 0.97:
                return rng.choice(others)
            return rng.choice(runnable)
        if victim in runnable and others and rng.random() >= p:       # a straggler: scheduled with probability p
            return rng.choice(others)
        return rng.choice(runnable)

    m.run(choose)
    return m


def test_kofn_word_level_model_accepts_only_consistent_gradients_and_replicas_converge():
    accepted = dropped = 0
    for seed in range(500):
        m = _kofn_run(seed, ring=[4, 8, 64][seed % 3], stall_at_commit_word=False)
        accepted += sum(m.accepted)
        dropped += sum(m.dropped)
    assert accepted > 1000 and dropped > 1000, (accepted, dropped)      # both the team path and the late / stale path ran


def test_kofn_bitmap_ring_tolerates_stalls_below_half_a_lap_and_the_model_sees_longer_ones():
    from distributedmnist_b200.parallel.protocol import ProtocolHazard
    for seed in range(150):
        _kofn_run(seed, ring=64, stall_at_commit_word=True)              # 12 steps can never lap half of the shipped ring
    with pytest.raises(ProtocolHazard, match="phantom bit"):
        for seed in range(150):
            _kofn_run(seed, ring=2, stall_at_commit_word=True)
    # the designed fix (NEXT_STEPS item 7): tagged arrival words survive the same stalls even on a two-slot ring
    dropped = 0
    for seed in range(300):
        dropped += sum(_kofn_run(seed, ring=2, stall_at_commit_word=bool(seed % 2), tagged=True).dropped)
    assert dropped > 100

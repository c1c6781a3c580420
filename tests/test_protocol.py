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

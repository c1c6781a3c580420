"""Property tests (hypothesis) of the host-side building blocks -- SURVEY §4 item 1/3: schedule vs closed form, data pipeline
epoch semantics, checkpoint round trip, K-of-N commit protocol under random arrival orders, summary-file framing."""
import math
import os
import random
import threading

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from distributedmnist_b200.checkpoint import Saver
from distributedmnist_b200.data import DataSet
from distributedmnist_b200.parallel.protocol import CommitBoard, popcount
from distributedmnist_b200.schedule import decay_steps_for, exponential_decay
from distributedmnist_b200.utils.summary import _record, encode_event, read_tfevents


@settings(max_examples=60, deadline=None)
@given(st.floats(1e-5, 1.0), st.integers(0, 100000), st.integers(1, 5000), st.floats(0.5, 1.0))
def test_staircase_schedule_matches_closed_form(lr0, step, decay_steps, rate):
    got = exponential_decay(lr0, step, decay_steps, rate, staircase=True)
    assert got == lr0 * rate ** (step // decay_steps)
    assert 0.0 <= got <= lr0 + 1e-12                 # (underflows to 0 after ~1075 halvings)
    # the staircase is constant inside a window and never increases
    assert exponential_decay(lr0, step - step % decay_steps, decay_steps, rate) == got
    assert exponential_decay(lr0, step + decay_steps, decay_steps, rate) <= got


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 100000), st.integers(1, 4096), st.floats(0.1, 8.0), st.integers(1, 64))
def test_decay_steps_rule_of_the_reference(n, batch, epochs, k):
    got = decay_steps_for(n, batch, epochs, k)
    assert got == max(int(n / float(batch) * epochs / k), 1)       # src/distributed_train.py:143-146 (+ the clamp)


@settings(max_examples=25, deadline=None)
@given(st.integers(5, 200), st.integers(1, 50), st.integers(0, 2 ** 31 - 1))
def test_dataset_draws_every_example_exactly_once_per_epoch(n, batch, seed):
    batch = min(batch, n)
    imgs = np.arange(n, dtype=np.float32).reshape(n, 1, 1, 1) + np.zeros((n, 2, 2, 1), np.float32)
    ds = DataSet(imgs, np.arange(n, dtype=np.int64), seed=seed)
    per_epoch = n // batch                      # a batch never straddles an epoch: the remainder is dropped (reference :112-130)
    for epoch in range(3):
        seen = []
        for _ in range(per_epoch):
            bx, by = ds.next_batch(batch)
            assert bx.shape[0] == batch and np.all(bx[:, 0, 0, 0] == by)          # images and labels stay paired
            seen.extend(int(v) for v in by)
        assert len(seen) == len(set(seen))                                          # no example twice within an epoch
        assert ds.epochs_completed == epoch
        if n % batch == 0 and per_epoch * batch == n:
            assert sorted(seen) == list(range(n))
    ds.next_batch(batch)
    assert ds.epochs_completed == 3 if per_epoch * batch + batch > n else True


@settings(max_examples=15, deadline=None)
@given(shapes=st.lists(st.tuples(st.integers(1, 6), st.integers(1, 7)), min_size=1, max_size=4), step=st.integers(0, 10 ** 9))
def test_checkpoint_round_trip_is_bit_exact(tmp_path_factory, shapes, step):
    d = str(tmp_path_factory.mktemp("ckpt"))
    g = torch.Generator().manual_seed(step % 1000)
    tensors = {"Variable_%d" % i if i else "Variable": torch.randn(*shp, generator=g) for i, shp in enumerate(shapes)}
    prefix = Saver().save(d, tensors, step)
    assert os.path.basename(prefix) == "model.ckpt-%d" % step
    got, gstep = Saver.restore(prefix)
    assert gstep == step and set(got) == set(tensors)
    assert all(torch.equal(got[k], tensors[k]) for k in tensors)
    assert Saver.latest(d) == prefix


@settings(max_examples=20, deadline=None)
@given(st.integers(2, 8), st.data())
def test_k_of_n_commit_every_replica_sees_the_same_mask(n, data):
    k = data.draw(st.integers(1, n))
    order_seed = data.draw(st.integers(0, 10 ** 6))
    board = CommitBoard(n, k)
    out = [None] * n

    def arrive(r):
        out[r] = board.arrive(r, 0, timeout=10.0)

    order = list(range(n))
    random.Random(order_seed).shuffle(order)
    ts = [threading.Thread(target=arrive, args=(r,)) for r in order]
    for t in ts:
        t.start()
    for t in ts:
        t.join(20)
    masks = {d.mask for d in out}
    assert len(masks) == 1                                        # one commit word per step
    mask = masks.pop()
    assert k <= popcount(mask) <= n and all(d.count == popcount(mask) for d in out)
    assert all(d.accepted == bool((mask >> r) & 1) for r, d in enumerate(out))
    assert all(d.global_step == 1 for d in out)


@settings(max_examples=30, deadline=None)
@given(wall=st.floats(0, 2e9), step=st.integers(0, 2 ** 62),
       scalars=st.dictionaries(st.text("abcdefgh /_", min_size=1, max_size=12), st.floats(-1e6, 1e6, width=32), max_size=4))
def test_event_records_survive_framing(tmp_path_factory, wall, step, scalars):
    p = os.path.join(str(tmp_path_factory.mktemp("ev")), "events.out.tfevents.test")
    with open(p, "wb") as f:
        f.write(_record(encode_event(wall, step, scalars)))
    (ev,) = read_tfevents(p)
    assert ev["step"] == step and math.isclose(ev["wall_time"], wall, rel_tol=0, abs_tol=0)
    assert set(ev["scalars"]) == set(scalars)
    assert all(ev["scalars"][k] == np.float32(v) for k, v in scalars.items())

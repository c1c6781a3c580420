"""Host side of the fused backend that needs no GPU: decoding of the page-locked status mirror the closing kernel of a step
writes (csrc/fused_sync.cuh::publish_status: 8 words per step, ring of 4, last word = sequence number)."""
import numpy as np
import pytest


def _backend(mirror, fallback):
    from distributedmnist_b200.parallel.context import ReplicaContext
    from distributedmnist_b200.parallel.fused import FusedBackend
    import torch
    be = object.__new__(FusedBackend)                     # no device: only the decoding logic is under test
    be.ctx = ReplicaContext(rank=1, world_size=2, local_rank=1, device=torch.device("cpu"), backend="none", store=None)
    be._mirror_np = mirror
    be.debug_sync = False
    be.read_status = lambda: dict(fallback)
    return be


def _publish(mirror, seq, epoch, error=0, accepted=0, dropped=0, mask=0, count=0, late=0):
    mirror[8 * ((seq - 1) & 3):8 * ((seq - 1) & 3) + 8] = [epoch, error, accepted, dropped, mask, count, late, seq]


def test_status_mirror_decodes_the_ring_slot_of_a_step():
    mirror = np.zeros(32, np.uint32)
    be = _backend(mirror, {"epoch": 999, "error": 0, "accepted_steps": 0, "dropped_steps": 0, "last_mask": 0, "last_count": 0, "last_late": 0})
    for seq in range(1, 11):                              # the ring wraps twice
        late = int(seq % 3 == 0)
        _publish(mirror, seq, epoch=seq, accepted=seq, mask=0b01 if late else 0b11, count=1 if late else 2, late=late)
        info = be.mirror_info(seq, check=True)
        assert (info.global_step, info.accepted, info.mask, info.count, info.stale) == \
            (seq, not late, 0b01 if late else 0b11, 1 if late else 2, bool(late))
    # the host fell more than a ring behind: slot 7's words belong to step 7 + 4 -> device read instead of a stale decode
    _publish(mirror, 11, epoch=11, mask=0b11, count=2)
    assert be.mirror_info(7).global_step == 999


def test_watchdog_error_word_raises_with_the_phase_name():
    mirror = np.zeros(32, np.uint32)
    be = _backend(mirror, {})
    _publish(mirror, 1, epoch=1, error=2, mask=0b11, count=2)
    with pytest.raises(RuntimeError, match="rank 1: push-complete timeout"):
        be.mirror_info(1, check=True)
    _publish(mirror, 2, epoch=2, error=1)
    with pytest.raises(RuntimeError, match="arrival timeout"):
        be.mirror_info(2, check=True)
    assert be.mirror_info(2, check=False).global_step == 2

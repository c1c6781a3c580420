"""LR schedule vs closed form (reference distributed_train.py:143-156)."""
import math

import pytest

from distributedmnist_b200.schedule import LearningRateSchedule, decay_steps_for, exponential_decay


def test_decay_steps_rule():
    # 60000/128 * 2.0 / 50 = 18.75 -> 18 ; the K=50,B=128,epochs=1 case hits 9
    assert decay_steps_for(60000, 128, 2.0, 50) == 18
    assert decay_steps_for(60000, 128, 1.0, 50) == 9
    assert decay_steps_for(60000, 256, 2.0, 8) == 58
    assert decay_steps_for(10, 1024, 1.0, 50) == 1   # clamped, never 0


@pytest.mark.parametrize("step", [0, 1, 17, 18, 19, 36, 1000])
def test_staircase_closed_form(step):
    lr = exponential_decay(0.1, step, 18, 0.98, staircase=True)
    assert lr == pytest.approx(0.1 * 0.98 ** math.floor(step / 18))


def test_continuous():
    assert exponential_decay(0.1, 9, 18, 0.5, staircase=False) == pytest.approx(0.1 * 0.5 ** 0.5)


def test_constant_when_factor_one():
    s = LearningRateSchedule(0.0008, 9, 1.0)
    assert all(s(i) == pytest.approx(0.0008) for i in (0, 5, 500))

"""hotloop.wave_plan: how the gather of a step's L batches is cut into waves (host logic, no GPU)."""
import pytest

from agent0_b200.hotloop import wave_plan

F = 84 * 84


def _sizes(p):
    return None if p is None else [w[1] for w in p]


def test_auto_plan_doubles_for_small_batches_and_goes_per_batch_for_large_ones():
    assert _sizes(wave_plan(20, 32, F)) == [1, 1, 2, 4, 12]            # 8 + a tail of 4 joins the 8
    assert _sizes(wave_plan(6, 32, F)) == [1, 1, 2, 2]
    assert _sizes(wave_plan(3, 16, F)) == [1, 1, 1]
    assert _sizes(wave_plan(2, 32, F)) == [1, 1]
    assert _sizes(wave_plan(16, 32, F)) == [1, 1, 2, 4, 8]
    assert _sizes(wave_plan(20, 512, F)) == [1] * 20
    assert _sizes(wave_plan(4, 128, F)) == [1] * 4                     # 128 x 16 x 7056 = 14 MB per batch
    assert _sizes(wave_plan(4, 64, F)) == [1, 1, 2]
    assert wave_plan(1, 512, F) is None and wave_plan(20, 32, F, None) is None and wave_plan(20, 32, F, 0) is None


def test_plan_rows_cover_every_draw_once_in_order():
    for L, B, spec in ((20, 32, "auto"), (20, 512, "auto"), (7, 12, [2, 5]), (5, 3, [1, 1, 1, 1, 1])):
        p = wave_plan(L, B, F, spec)
        assert sum(w[1] for w in p) == L and sum(w[3] for w in p) == L * B
        b0 = 0
        for first, n, lo, cnt in p:
            assert first == b0 and lo == b0 * B and cnt == n * B and n > 0
            b0 += n


def test_explicit_plans_are_validated():
    assert wave_plan(20, 32, F, [20]) is None                         # a single wave is the plain launch
    with pytest.raises(AssertionError):
        wave_plan(20, 32, F, [1, 2, 3])
    with pytest.raises(AssertionError):
        wave_plan(4, 32, F, [4, 0])
    with pytest.raises(AssertionError):
        wave_plan(4, 32, F, "fast")

"""CPU tests of the host-side stand-ins for the un-vendored dphtools helpers (pyotf/utils.py:11 imports them).

UNPINNED against dphtools itself (not installed, not in /root/reference): these tests fix the documented rules and show
that the one ambiguous rule (how a window over the lower edge is clipped) cannot change `prep_data_for_PR`.
"""
import numpy as np
import pytest

from pyotf_b200.utils import fft_pad, prep_data_for_PR, slice_maker


def test_slice_maker_windows_that_fit():
    assert slice_maker((4, 4), 3) == (slice(3, 6), slice(3, 6))
    assert slice_maker((4, 4), 4) == (slice(2, 6), slice(2, 6))
    assert slice_maker((3, 2), (3, 4)) == (slice(2, 5), slice(0, 4))
    assert slice_maker((10.4, 7.6), 5.2) == (slice(8, 13), slice(6, 11))      # centres and widths are rounded


def test_slice_maker_clips_only_the_start():
    assert slice_maker((1,), 6) == (slice(0, 4),)          # [-2, 4) -> [0, 4): shorter, not shifted
    assert slice_maker((0,), 3) == (slice(0, 2),)
    assert slice_maker((-5,), 4) == (slice(0, 0),)         # ends before 0: empty
    with pytest.raises(ValueError):
        slice_maker((3,), -1)


@pytest.mark.parametrize("n", range(1, 12))
@pytest.mark.parametrize("w", range(1, 14))
def test_prep_crop_window_is_the_same_under_either_clipping_rule(n, w):
    """prep_data_for_PR centres its crop on (n + 1) // 2 (pyotf/utils.py:197): whenever the start is clipped the stop
    lies beyond the array, so `slice(0, stop)` and `slice(0, w)` select the same elements."""
    c = (n + 1) // 2
    (s,) = slice_maker((c,), w)
    lo = max(c - w // 2, 0)
    a = np.arange(n)
    np.testing.assert_array_equal(a[s], a[lo:lo + w])


@pytest.mark.parametrize("shape_in,xysize,shape_out", [((5, 5), None, (5, 5)), ((5, 4), None, (5, 5)),
                                                       ((5, 4), 10, (10, 10)), ((5, 4), 3, (3, 3))])
def test_prep_pr_size_positive(shape_in, xysize, shape_out):
    """The reference's own test (tests/utils_test.py:53-72), restated."""
    data = np.ones((10,) + shape_in, dtype=int)
    data[0, 0, 0] = 2
    data[-1, -1, -1] = 0
    out = prep_data_for_PR(data, xysize=xysize)
    assert out.shape == (10,) + shape_out
    assert np.all(out >= 0)


@pytest.mark.parametrize("size", [3, 4, 5, 6, 7])
def test_prep_pr_center(size):
    """tests/utils_test.py:75-91, restated: the brightest voxel lands on the FFT centre of the crop."""
    data = np.ones((10, 8, 8), dtype=int)
    data[0, 0, 0] = 2
    data[-1, -1, -1] = 0
    out = prep_data_for_PR(data, xysize=size)
    assert out.shape == (10, size, size)
    nz, ny, nx = out.shape
    assert np.unravel_index(out.argmax(), out.shape) == (nz // 2, ny // 2, nx // 2)


def test_fft_pad_keeps_the_fft_centre():
    for n in (4, 5, 10, 11):
        for m in (n, n + 1, n + 4, n + 5):
            a = np.zeros(n)
            a[n // 2] = 1.0
            b = fft_pad(a, (m,))
            assert b.shape == (m,) and b[m // 2] == 1.0 and b.sum() == 1.0
    with pytest.raises(ValueError):
        fft_pad(np.zeros(8), (6,))

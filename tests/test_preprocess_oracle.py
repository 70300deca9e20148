"""The bicubic uint8 resize of the pre-processing (model/decode_np.py:125-134 -> cv2.resize INTER_CUBIC): the oracle's
restatement of OpenCV's own code against the reference's output (tests/golden/resize.npz, make_golden_resize.py):
bit-exact against the run with OpenCV's own code, within +-1 level on < 6 % of the pixels against the run with the wheel's
Intel IPP primitive (whose arithmetic is closed source)."""
import os

import numpy as np
import pytest

from oracle import ppyolo_ref as ref

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_oracle_resize_vs_reference(tag):
    z = np.load(os.path.join(HERE, 'golden', 'resize.npz'))
    img, size = z['img_' + tag], int(z['size_' + tag])
    got = ref.resize_cubic_u8(img[:, :, ::-1], size)                 # the reference converts BGR -> RGB first
    np.testing.assert_array_equal(got, z['ocv_' + tag])
    d = np.abs(got.astype(int) - z['ipp_' + tag].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.06

"""model.decode_np.Decode (reference model/decode_np.py:21-150): pre-processing bit-identical to the reference class (golden),
and the predict / detect_image / detect_batch plumbing around the CUDA model."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (os.path.join(REPO, 'pytorch-ppyolo_b200'), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

cv2 = pytest.importorskip('cv2')

import config as cfgs  # noqa: E402
from model.decode_np import Decode  # noqa: E402

CLASSES = ['c%d' % i for i in range(80)]


def test_process_image_matches_reference_golden():
    z = np.load(os.path.join(HERE, 'golden', 'preprocess.npz'))
    for tag in ('a', 'b'):
        cfg = cfgs.PPYOLO_2x_Config()
        cfg.test_cfg['target_size'] = int(z['size_' + tag])
        d = Decode(None, CLASSES, False, cfg, for_test=True)
        img = z['img_' + tag]
        pimage, im_size = d.process_image(img.copy())
        assert pimage.dtype == np.float32 and pimage.shape == z['pimage_' + tag].shape
        np.testing.assert_array_equal(pimage, z['pimage_' + tag])
        np.testing.assert_array_equal(im_size, z['im_size_' + tag])
        assert im_size.dtype == np.int32


def test_unpack_sentinel_and_rows():
    none = np.zeros((1, 6), dtype=np.float32) - 1.0
    b, s, c = Decode._unpack(none)
    assert len(b) == 0 and len(s) == 0 and len(c) == 0
    rows = np.array([[3., .9, 1, 2, 3, 4], [7., .5, 5, 6, 7, 8]], dtype=np.float32)
    b, s, c = Decode._unpack(rows)
    assert b.shape == (2, 4) and c.dtype == np.int32 and list(c) == [3, 7] and np.allclose(s, [.9, .5])


@pytest.mark.gpu
def test_detect_image_and_batch_on_gpu():
    from tests.helpers import build_model
    model, cfg = build_model('r18vd')
    model = model.cuda().eval()
    cfg.test_cfg['target_size'] = 128
    d = Decode(model, CLASSES, True, cfg, for_test=True)
    rng = np.random.RandomState(3)
    imgs = [rng.randint(0, 256, size=(90 + 10 * i, 120, 3)).astype(np.uint8) for i in range(2)]
    pairs = [d.process_image(im.copy()) for im in imgs]
    image, boxes, scores, classes = d.detect_image(imgs[0].copy(), pairs[0][0], pairs[0][1], draw_image=True, draw_thresh=0.0)
    assert image.shape == imgs[0].shape and len(boxes) == len(scores) == len(classes)
    batch_p = np.concatenate([p for p, _ in pairs], axis=0)
    batch_s = np.concatenate([s for _, s in pairs], axis=0)
    out_img, out_boxes, out_scores, out_classes = d.detect_batch([im.copy() for im in imgs], batch_p, batch_s, draw_image=True)
    assert len(out_img) == 2
    # image 0 alone and inside the batch: same detections (per-image independence of the whole path)
    assert len(out_scores[0]) == len(scores)
    if len(scores):
        np.testing.assert_allclose(out_scores[0], scores, rtol=2e-2, atol=2e-3)
        assert out_boxes[0].shape[1] == 4 and out_classes[0].dtype == np.int32


def test_process_image_u8_plus_table_equals_process_image():
    """The uint8 fast path splits process_image in two: cv2 steps on the host (process_image_u8), NormalizeImage + Permute as a
    3 x 256 table applied on the GPU.  Table(process_image_u8(img)) must equal process_image(img) bit for bit."""
    import numpy as np
    from ppyolo_b200.engine import normalize_lut
    from model.decode_np import Decode
    import config as cfgs
    cfg = cfgs.PPYOLO_r18vd_Config()
    d = Decode(None, ['c%d' % i for i in range(80)], False, cfg, for_test=True)
    img = np.random.RandomState(7).randint(0, 256, (240, 320, 3)).astype(np.uint8)
    want, im_size = d.process_image(img.copy())
    u8, im_size2 = d.process_image_u8(img.copy())
    assert u8.dtype == np.uint8 and u8.shape == (1, d.target_size, d.target_size, 3) and np.array_equal(im_size, im_size2)
    lut = normalize_lut([float(v) for v in d.mean], [float(v) for v in d.std], d.is_scale)
    got = np.stack([lut[c][u8[0, :, :, c]] for c in range(3)], 0)[None]
    assert np.array_equal(got, want)

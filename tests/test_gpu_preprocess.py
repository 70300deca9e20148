"""GPU pre-processing: the batched bicubic uint8 resize kernel (csrc/preprocess.cu) against the reference's
pre-processing output (tests/golden/resize.npz: Decode.process_image of the unmodified reference, model/decode_np.py:125-140).

Bar: BIT-EXACT against the reference run with OpenCV's own resize code; against the run with the wheel's Intel IPP primitive
(closed-source arithmetic) at most +-1 level on < 6 % of the pixels -- the same gap OpenCV's own code has to IPP.  At full size
(480x640 / 427x640 / 1080x1920 -> 608) the kernel is compared with the oracle restatement, and the end-to-end effect of the
+-1 levels on the head outputs is measured."""
import os

import numpy as np
import pytest
import torch

from oracle import ppyolo_ref as ref
from tests.helpers import build_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'
HERE = os.path.dirname(os.path.abspath(__file__))


def gpu_resize(images, size, swap_rb=True):
    from ppyolo_b200 import ops
    blob, meta = ops.pack_images(images)
    return ops.resize_cubic_u8(blob.to(DEV), meta.to(DEV), size, swap_rb=swap_rb).cpu().numpy()


def test_resize_kernel_vs_reference_golden():
    z = np.load(os.path.join(HERE, 'golden', 'resize.npz'))
    for tags in (['a'], ['b'], ['c'], ['d']):
        imgs = [z['img_' + t] for t in tags]
        size = int(z['size_' + tags[0]])
        got = gpu_resize(imgs, size)
        for i, t in enumerate(tags):
            np.testing.assert_array_equal(got[i], z['ocv_' + t])
            d = np.abs(got[i].astype(int) - z['ipp_' + t].astype(int))
            print('resize %s: vs the IPP run %.2f %% of the pixels differ, max %d' % (t, 100.0 * (d > 0).mean(), d.max()))
            assert d.max() <= 1 and (d > 0).mean() < 0.06


def test_resize_kernel_batch_of_mixed_sizes_vs_oracle():
    """One launch over a batch of differently sized images at the real target size: bit-exact against the oracle, per image."""
    rng = np.random.RandomState(5)
    shapes = [(480, 640), (427, 640), (640, 480), (1080, 1920), (333, 500), (608, 608), (97, 131), (720, 1280)]
    imgs = []
    for h, w in shapes:
        base = rng.randint(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.uint8)
        im = np.kron(base, np.ones((8, 8, 1), np.uint8))[:h, :w]                      # blocky image with hard edges (overshoot)
        imgs.append(np.ascontiguousarray(im + rng.randint(0, 8, im.shape).astype(np.uint8) // 2))
    got = gpu_resize(imgs, 608)
    for i, im in enumerate(imgs):
        np.testing.assert_array_equal(got[i], ref.resize_cubic_u8(im[:, :, ::-1], 608), err_msg=str(shapes[i]))
    same = gpu_resize(imgs, 608, swap_rb=False)
    np.testing.assert_array_equal(same[..., ::-1], got)


def test_decode_gpu_preprocessing_end_to_end():
    """Decode.process_batch_gpu -> predict: same detections as the host path (cv2.resize + the uint8 engine input) up to the
    +-1-level difference between OpenCV's own resize code and the wheel's IPP primitive; head-output drift measured and bounded
    (5e-3 of scale: a +-1/255 input perturbation on ~1-3 % of the pixels), >= 80 % of the top detections matched at IoU >= 0.9."""
    import config as cfgs
    from model.decode_np import Decode
    model, cfg = build_model('r18vd')
    model = model.to(DEV).eval()
    cfg.test_cfg['target_size'] = 320
    dec = Decode(model, ['c%d' % i for i in range(80)], True, cfg, for_test=True)
    rng = np.random.RandomState(9)
    imgs = []
    for h, w in ((240, 320), (375, 500)):
        base = rng.randint(0, 256, (h // 16 + 2, w // 16 + 2, 3)).astype(np.uint8)
        imgs.append(np.ascontiguousarray(np.kron(base, np.ones((16, 16, 1), np.uint8))[:h, :w]))
    dev_batch, im_size = dec.process_batch_gpu(imgs)
    host = [dec.process_image_u8(im.copy())[0][0] for im in imgs]
    d = np.abs(dev_batch.cpu().numpy().astype(int) - np.stack(host).astype(int))
    print('GPU resize vs cv2 (as installed): %.2f %% of the pixels differ, max %d' % (100.0 * (d > 0).mean(), d.max()))
    assert d.max() <= 1
    p_gpu = dec.predict(dev_batch, im_size)
    o_gpu = [o.clone() for o in model.engine(2, 320, 320, input_u8=True).head_outputs_nchw()]
    p_host = dec.predict(np.stack(host), im_size)
    o_host = model.engine(2, 320, 320, input_u8=True).head_outputs_nchw()
    for a, b in zip(o_gpu, o_host):
        drift = float((a - b).abs().max()) / float(b.abs().max())
        print('head-output drift from the +-1 levels: %.2e of scale' % drift)
        assert drift < 5e-3
    for a, b in zip(p_gpu, p_host):
        assert a.shape[1] == 6 and b.shape[1] == 6
        if a[0][0] < 0 or b[0][0] < 0:
            continue
        n = min(len(a), len(b), 20)                      # the strongest detections must (mostly: a seeded random net) agree
        hits = 0
        for row in b[:n]:
            ious = _iou_1toN(row[2:], a[:, 2:]) * (a[:, 0] == row[0])
            hits += int(ious.max() >= 0.9)
        print('top-%d detections matched at IoU >= 0.9 with the same label: %d' % (n, hits))
        assert hits >= 0.8 * n


def _iou_1toN(box, boxes):
    x0, y0 = np.maximum(box[0], boxes[:, 0]), np.maximum(box[1], boxes[:, 1])
    x1, y1 = np.minimum(box[2], boxes[:, 2]), np.minimum(box[3], boxes[:, 3])
    inter = np.clip(x1 - x0, 0, None) * np.clip(y1 - y0, 0, None)
    area = lambda b: (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    return inter / (area(box) + area(boxes) - inter + 1e-9)

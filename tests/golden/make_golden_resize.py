"""tests/golden/resize.npz: the uint8 image the UNMODIFIED reference feeds its normalisation -- Decode.process_image
(model/decode_np.py:125-140: BGR->RGB, cv2.resize(..., interpolation=cv2.INTER_CUBIC)) run twice per image: with OpenCV's own code
(cv2.ipp.setUseIPP(False): `ocv_*`, what csrc/preprocess.cu restates bit for bit) and with the wheel's default Intel IPP primitive
(`ipp_*`: differs from OpenCV's own code by +-1 on a few per cent of the pixels).  The uint8 image is recovered from the
reference's normalised float output by inverting (x / 255 - mean) / std and rounding (exact: the float error is ~1e-5 of a level).
Run in the authoring container:  python tests/golden/make_golden_resize.py"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from config import PPYOLO_2x_Config  # noqa: E402
from model.decode_np import Decode  # noqa: E402

out = {}
rng = np.random.RandomState(1)
for tag, (h, w), size in (('a', (97, 131), 64), ('b', (60, 45), 96), ('c', (300, 400), 128), ('d', (213, 160), 160)):
    img = cv2.GaussianBlur(rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8), (0, 0), 1.5)
    img[5:35, 5:30] = (np.linspace(0, 255, 25)[None, :, None] * np.ones((30, 1, 3))).astype(np.uint8)
    img[40:50, 10:40] = rng.randint(0, 256, size=(10, 30, 3))          # a patch of noise: overshoot -> saturation
    cfg = PPYOLO_2x_Config()
    cfg.test_cfg['target_size'] = size
    d = Decode(None, ['c%d' % i for i in range(80)], False, cfg, for_test=True)
    mean, std = np.array(cfg.normalizeImage['mean']), np.array(cfg.normalizeImage['std'])
    out['img_' + tag], out['size_' + tag] = img, np.int64(size)
    for name, use_ipp in (('ocv', False), ('ipp', True)):
        cv2.ipp.setUseIPP(use_ipp)
        pimage, _ = d.process_image(img.copy())
        u = (pimage[0].transpose(1, 2, 0).astype(np.float64) * std + mean) * 255.0
        r = np.rint(u)
        assert np.abs(u - r).max() < 1e-3, np.abs(u - r).max()
        out['%s_%s' % (name, tag)] = r.astype(np.uint8)
    diff = np.abs(out['ocv_' + tag].astype(int) - out['ipp_' + tag].astype(int))
    print(tag, (h, w), size, 'IPP vs OpenCV code: %.2f %% of pixels differ, max %d' % (100.0 * (diff > 0).mean(), diff.max()))
np.savez_compressed(os.path.join(HERE, 'resize.npz'), **out)

"""tests/golden/preprocess.npz from the UNMODIFIED reference Decode.process_image (model/decode_np.py:125-140).
Run in the authoring container:  python tests/golden/make_golden_preprocess.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from config import PPYOLO_2x_Config  # noqa: E402
from model.decode_np import Decode  # noqa: E402

out = {}
rng = np.random.RandomState(0)
for tag, (h, w), size in (('a', (97, 131), 64), ('b', (60, 45), 96)):
    img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    # smooth it a little so the bicubic resize sees image-like content too
    img[10:40, 5:30] = (np.linspace(0, 255, 25)[None, :, None] * np.ones((30, 1, 3))).astype(np.uint8)
    cfg = PPYOLO_2x_Config()
    cfg.test_cfg['target_size'] = size
    d = Decode(None, ['c%d' % i for i in range(80)], False, cfg, for_test=True)
    pimage, im_size = d.process_image(img.copy())
    out['img_' + tag], out['pimage_' + tag], out['im_size_' + tag], out['size_' + tag] = img, pimage, im_size, np.int64(size)
np.savez_compressed(os.path.join(HERE, 'preprocess.npz'), **out)
print({k: v.shape for k, v in out.items()})

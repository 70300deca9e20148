"""tests/golden/dcn_far.npz: the UNMODIFIED reference DCNv2 (model/custom_layers.py:486-677, inside a Conv2dUnit) with LARGE
offsets -- sigma ~6 px plus an N(0,1) bias, i.e. samples up to ~+-20 px from their tap on a 20x20 map, most of them far outside
the image and beyond the reference's one-pixel zero border (custom_layers.py:571-574, clamp at :614-615).  Pins the oracle's
"outside the image -> 0" rule to the reference's clamp-into-the-border trick for far offsets too.

Run in the authoring container:  python tests/golden/make_golden_dcn_far.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
torch.Tensor.cuda = lambda self, *a, **k: self        # the reference hard-codes .cuda()
torch.set_num_threads(1)
sys.path.insert(0, '/root/reference')
from model import custom_layers as ref_layers  # noqa: E402

spec = importlib.util.spec_from_file_location('synth', os.path.join(REPO, 'pytorch-ppyolo_b200', 'ppyolo_b200', 'synth.py'))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

out = {}
g = torch.Generator().manual_seed(41)
for tag, stride in (('s1', 1), ('s2', 2)):
    u = ref_layers.Conv2dUnit(64, 24, 3, stride=stride, bn=1, act='relu', use_dcn=True)
    synth.randomize_(u, seed=33, offset_scale=0.25)       # offsets = sum of 576 terms * N(0, .25): sigma ~6 px
    u.eval()
    x = torch.randn((2, 64, 20, 20), generator=g)
    with torch.no_grad():
        om = u.conv.conv_offset(x)
        out[tag + '_in'] = x.numpy()
        out[tag + '_offsetmask'] = om.numpy()
        out[tag + '_raw'] = u.conv(x).numpy()
        out[tag + '_out'] = u(x).numpy()
    print(tag, 'offset abs max %.1f px, std %.1f px' % (float(om[:, :18].abs().max()), float(om[:, :18].std())))
np.savez_compressed(os.path.join(HERE, 'dcn_far.npz'), **out)
print('dcn_far.npz %.1f KB' % (os.path.getsize(os.path.join(HERE, 'dcn_far.npz')) / 1024.0))

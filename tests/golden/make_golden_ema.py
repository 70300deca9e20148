"""tests/golden/ema.npz from the UNMODIFIED reference class model/EMA.py (CPU-only methods register/update).
Run in the authoring container:  python tests/golden/make_golden_ema.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from model.EMA import ExponentialMovingAverage  # noqa: E402


def toy(seed):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.BatchNorm2d(5), torch.nn.Conv2d(5, 7, 1))
    m[2].bias.requires_grad = False          # frozen tensors are skipped (model/EMA.py:27)
    return m


out = {}
m = toy(0)
ema = ExponentialMovingAverage(m, 0.9998)
ema.register()
g = torch.Generator().manual_seed(1)
for step in range(12):
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    d = ema.update()
    out['decay_%d' % step] = np.float64(d)
    if step in (0, 1, 5, 11):
        for name, arr in ema._shadow.items():
            out['shadow_%d_%s' % (step, name)] = np.asarray(arr, dtype=np.float32)
for name, p in m.named_parameters():
    out['final_' + name] = p.detach().numpy().copy()
np.savez_compressed(os.path.join(HERE, 'ema.npz'), **out)
print('wrote ema.npz with', len(out), 'arrays')

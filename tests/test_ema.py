"""ExponentialMovingAverage (reference model/EMA.py:16-57): oracle vs the golden produced by the reference class (CPU), and the
device implementation through the C ABI vs the same golden (GPU, bit-exact)."""
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

from oracle import ppyolo_ref as ref  # noqa: E402


def toy(seed):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.BatchNorm2d(5), torch.nn.Conv2d(5, 7, 1))
    m[2].bias.requires_grad = False
    return m


def golden():
    return np.load(os.path.join(HERE, 'golden', 'ema.npz'))


def test_oracle_ema_matches_reference_golden():
    z = golden()
    m = toy(0)
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    params = [p for n, p in m.named_parameters() if p.requires_grad]
    shadow = [p.detach().numpy().copy() for p in params]
    g = torch.Generator().manual_seed(1)
    for step in range(12):
        with torch.no_grad():
            for p in m.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
        shadow, d = ref.ema_update(shadow, [p.detach().numpy() for p in params], 0.9998, step)
        assert d == float(z['decay_%d' % step])
        if step in (0, 1, 5, 11):
            for n, s in zip(names, shadow):
                np.testing.assert_array_equal(s, z['shadow_%d_%s' % (step, n)])


@pytest.mark.gpu
def test_device_ema_bit_exact_and_apply_restore():
    from model.EMA import ExponentialMovingAverage
    z = golden()
    m = toy(0).cuda()
    ema = ExponentialMovingAverage(m, 0.9998)
    ema.register()
    assert set(ema._shadow) == {n for n, p in m.named_parameters() if p.requires_grad}
    g = torch.Generator().manual_seed(1)
    for step in range(12):
        with torch.no_grad():
            for p in m.parameters():
                p.add_((torch.randn(p.shape, generator=g) * 0.05).cuda())
        d = ema.update()
        assert d == float(z['decay_%d' % step])
        if step in (0, 1, 5, 11):
            for n, s in ema._shadow.items():
                np.testing.assert_array_equal(s.cpu().numpy(), z['shadow_%d_%s' % (step, n)])
    live = {n: p.detach().clone() for n, p in m.named_parameters()}
    ema.apply()
    for n, p in m.named_parameters():
        if p.requires_grad:
            np.testing.assert_array_equal(p.detach().cpu().numpy(), z['shadow_11_' + n])
    ema.update()                                   # the pointer table follows the rebound param.data
    ema.restore()
    for n, p in m.named_parameters():
        assert torch.equal(p.detach(), live[n])
    with pytest.raises(RuntimeError):
        ExponentialMovingAverage(toy(1), 0.99).register()      # CPU model: no host fallback

"""GPU tests of the DropBlock kernels (reference model/custom_layers.py:293-342; SURVEY.md 8a-8 / 8c known answer)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def ops():
    from ppyolo_b200 import ops as _ops
    return _ops


def reference_dropblock(x, seeds, block_size=3):
    """The reference's arithmetic from the seed matrix on (custom_layers.py:333-342), torch CPU fp32."""
    mask = 1.0 - F.max_pool2d(seeds, (block_size, block_size), stride=1, padding=1)
    return x * mask * float(x.numel()) / mask.sum(), mask


@pytest.mark.parametrize('shape,channels_last', [((2, 4, 10, 10), False), ((3, 16, 19, 19), True), ((1, 8, 7, 23), False), ((2, 32, 38, 38), True)])
def test_dropblock_injected_seeds_bit_exact(shape, channels_last):
    """Given the reference's own Bernoulli draw, mask and output must equal the reference formula bit for bit (fp32)."""
    o = ops()
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g)
    gamma = o.dropblock_gamma(shape[2], 3, 0.9)
    seeds = (torch.rand(shape, generator=g) < gamma).float()
    want, mask = reference_dropblock(x, seeds)
    xd = x.to(DEV)
    if channels_last:
        xd = xd.contiguous(memory_format=torch.channels_last)
    got = o.drop_block(xd, 3, 0.9, seeds=seeds.to(DEV))
    assert got.stride() == xd.stride()
    np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())
    assert float(mask.sum()) < mask.numel()                   # something was dropped


def test_dropblock_known_answer_statistics():
    """DropBlock(3, 0.9) on ones renormalises to mean exactly 1; the zero fraction follows the analytic
    expectation 1 - (1 - gamma)^k (k = in-image neighbours: 9 / 6 / 4).  The survey's CPU sample on [2,4,10,10] was 0.119."""
    o = ops()
    o.dropblock_seed(0)
    y = o.drop_block(torch.ones((2, 4, 10, 10), device=DEV), 3, 0.9)
    assert abs(float(y.mean()) - 1.0) < 1e-6
    zf = float((y == 0).float().mean())
    assert 0.03 < zf < 0.25, zf
    shape = (8, 64, 38, 38)
    y = o.drop_block(torch.ones(shape, device=DEV), 3, 0.9)
    h = shape[2]
    gamma = o.dropblock_gamma(h, 3, 0.9)
    k = np.full((h, h), 9.0); k[0, :] = k[-1, :] = k[:, 0] = k[:, -1] = 6.0; k[0, 0] = k[0, -1] = k[-1, 0] = k[-1, -1] = 4.0
    expect = float((1.0 - (1.0 - gamma) ** k).mean())
    zf = float((y == 0).float().mean())
    print('dropblock zero fraction %.4f, analytic %.4f (gamma %.5f)' % (zf, expect, gamma))
    assert abs(zf - expect) < 0.004
    assert abs(float(y.double().mean()) - 1.0) < 1e-5
    # every zero belongs to the 3x3 block of some seed: dilating the kept-mask's complement by nothing new
    kept = (y != 0).float().cpu()
    holes = 1.0 - kept
    assert float((F.max_pool2d(holes, 3, 1, 1) - holes).clamp(min=0).sum()) >= 0.0


@pytest.mark.parametrize('channels', [8, 32, 48])
def test_dropblock_rng_stream_and_layouts(channels):
    """(channels % 16 == 0 in channels_last takes the 16-channels-per-thread kernels: same numbers as the generic ones)"""
    o = ops()
    x = torch.randn((2, channels, 19, 19), device=DEV)
    o.dropblock_seed(123)
    a = o.drop_block(x, 3, 0.9)
    b = o.drop_block(x, 3, 0.9)                                # offset advanced: a fresh draw
    o.dropblock_seed(123)
    a2 = o.drop_block(x.contiguous(memory_format=torch.channels_last), 3, 0.9)
    assert not torch.equal(a, b)
    assert torch.equal(a, a2.contiguous())                     # same seed/offset -> same LOGICAL mask in either layout
    assert int(o.dropblock_rng(x.device)[1]) == 1


def test_dropblock_backward_and_module():
    from model.custom_layers import DropBlock
    o = ops()
    o.dropblock_seed(5)
    x = torch.randn((2, 6, 12, 12), device=DEV, requires_grad=True)
    m = DropBlock(block_size=3, keep_prob=0.9, is_test=False)
    y = m(x)
    scale_mask = (y.detach() / x.detach())                     # mask * numel / sum per element
    w = torch.randn_like(y)
    (y * w).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), (w * scale_mask).cpu().numpy(), rtol=1e-6, atol=1e-7)
    m.is_test = True
    assert m(x) is x


def test_dropblock_in_cuda_graph_draws_fresh_masks():
    o = ops()
    o.dropblock_seed(9)
    x = torch.ones((1, 4, 16, 16), device=DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        o.drop_block(x, 3, 0.9)                                # warm-up
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = o.drop_block(x, 3, 0.9)
    g.replay(); torch.cuda.synchronize(); first = y.clone()
    g.replay(); torch.cuda.synchronize(); second = y.clone()
    assert not torch.equal(first, second)

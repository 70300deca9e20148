"""GPU parity of the fp32-grade tensor-core path (precision 'f16x2', ppy_conv_f16x2): activations and weights are fp16
hi/lo pairs (22 significant bits), every K block runs hi*hi + hi*lo + lo*hi on tcgen05 into an fp32 TMEM accumulator.

Tolerance (stated here, used below): every kernel within 1e-5 of the output scale of an fp64 evaluation of the same
operator on the same fp32 inputs -- the band an fp32 FMA chain itself occupies (K up to 4608); measured errors are printed.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ppyolo_ref as ref
from ppyolo_b200 import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-5


def ops():
    from ppyolo_b200 import ops as _ops
    return _ops


def scale_of(t):
    return float(np.abs(t).max()) + 1e-12


def to_pair(x_nchw):
    o = ops()
    from ppyolo_b200._lib import PPY_F32
    return o.split_pair(o.to_nhwc(x_nchw.to(DEV), PPY_F32, o.round_up(x_nchw.shape[1], 8)))


def from_pair(y, c):
    return ops().join_pair(y, c).permute(0, 3, 1, 2).contiguous().cpu()


def run_pair_conv(x, w, scale, shift, stride, act, residual=None, bias_map=None, upsample=False, out_f32=False, offset_mask=None,
                  overflow=None):
    o = ops()
    cout, cin, k, _ = w.shape
    y = o.conv_pair(to_pair(x), w.to(DEV), scale.to(DEV), shift.to(DEV), stride, (k - 1) // 2, act,
                    residual=to_pair(residual) if residual is not None else None,
                    bias_map=bias_map.to(DEV) if bias_map is not None else None, out_f32=out_f32, upsample2x=upsample,
                    offset_mask=offset_mask, c_count=cin, overflow=overflow)
    return o.from_nhwc(y, cout).cpu() if out_f32 else from_pair(y, cout)


def act64(y, act):
    if act == 1:
        return torch.relu(y)
    if act == 2:
        return torch.where(y > 0, y, 0.1 * y)
    return y


def check(got, want, what, tol=TOL):
    err = float((got.double() - want).abs().max()) / scale_of(want.numpy())
    print('%s: max err / scale = %.2e' % (what, err))
    assert err < tol, (what, err)


SHAPES = [  # n, cin, cout, k, stride, hw          kernel mode
    (2, 64, 256, 1, 1, 19),     # tma_a, CTA pair, BLOCK_N 256
    (1, 128, 512, 1, 1, 30),    # tma_a
    (3, 64, 96, 1, 1, 11),      # tma_a, BLOCK_N 128, cout off the tile
    (1, 2048, 512, 1, 1, 19),   # tma_a, long K (32 K blocks)
    (2, 64, 64, 3, 1, 32),      # slab, BLOCK_N 64
    (1, 128, 128, 3, 1, 24),    # slab, BLOCK_N 128
    (2, 256, 512, 3, 1, 16),    # patch
    (1, 512, 1024, 3, 1, 19),   # im2col (19x19 wastes the patch grid), K = 4608
    (1, 128, 256, 3, 2, 20),    # im2col stride 2
    (2, 512, 27, 3, 1, 19),     # BLOCK_N 32 (DCN offset conv shape)
    (1, 64, 32, 3, 1, 8),       # single tile
    (2, 32, 32, 3, 1, 32),      # 32-element K blocks (SWIZZLE_64B build), BLOCK_N 32 (stem conv1_2)
    (1, 32, 64, 3, 1, 48),      # 32-element K blocks, BLOCK_N 64, CTA pair (stem conv1_3)
    (3, 32, 128, 3, 1, 40),     # 32-element K blocks, BLOCK_N 128
    (1, 32, 48, 3, 1, 19),      # cin 32 on a map the slab grid wastes: gather mode
    (2, 8, 16, 3, 1, 9),        # gather mode (cin < 64)
    (1, 24, 40, 1, 1, 7),       # gather mode 1x1
]


@pytest.mark.parametrize('n,cin,cout,k,stride,hw', SHAPES)
def test_pair_conv_shapes(n, cin, cout, k, stride, hw):
    g = torch.Generator().manual_seed(cin * 131 + cout * 7 + k + hw)
    x = torch.randn((n, cin, hw, hw), generator=g)
    w = torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    want = F.conv2d(x.double(), w.double(), stride=stride, padding=(k - 1) // 2) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    for out_f32 in (False, True):
        got = run_pair_conv(x, w, scale, shift, stride, 0, out_f32=out_f32)
        check(got, want, 'conv %s out_f32=%d' % ((n, cin, cout, k, stride, hw), out_f32))


def test_pair_conv_epilogue_variants():
    g = torch.Generator().manual_seed(5)
    n, cin, cout, hw = 2, 128, 256, 19
    x = torch.randn((n, cin, hw, hw), generator=g)
    w = torch.randn((cout, cin, 1, 1), generator=g) / cin ** 0.5
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    res = torch.randn((n, cout, hw, hw), generator=g)
    bm = torch.randn((hw * hw, cout), generator=g) * 0.3
    base = F.conv2d(x.double(), w.double())
    aff = lambda t: t * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    check(run_pair_conv(x, w, scale, shift, 1, 1, residual=res), act64(aff(base) + res.double(), 1), 'residual+relu')
    check(run_pair_conv(x, w, scale, shift, 1, 2), act64(aff(base), 2), 'leaky')
    bm64 = bm.double().t().reshape(1, cout, hw, hw)
    check(run_pair_conv(x, w, scale, shift, 1, 2, bias_map=bm), act64(aff(base + bm64), 2), 'bias map')
    up = F.interpolate(act64(aff(base), 2), scale_factor=2, mode='nearest')
    check(run_pair_conv(x, w, scale, shift, 1, 2, upsample=True), up, 'upsample')
    # a 3x3 with residual on the slab / patch loaders, cout not a multiple of 8 through the slow epilogue
    w3 = torch.randn((128, cin, 3, 3), generator=g) / (cin * 9) ** 0.5
    r3 = torch.randn((n, 128, hw, hw), generator=g)
    check(run_pair_conv(x, w3, scale[:128], shift[:128], 1, 1, residual=r3),
          act64(F.conv2d(x.double(), w3.double(), padding=1) * scale[:128].double().view(1, -1, 1, 1) + shift[:128].double().view(1, -1, 1, 1) + r3.double(), 1),
          '3x3 residual')
    w5 = torch.randn((13, cin, 1, 1), generator=g) / cin ** 0.5
    check(run_pair_conv(x, w5, scale[:13], shift[:13], 1, 0, out_f32=True),
          F.conv2d(x.double(), w5.double()) * scale[:13].double().view(1, -1, 1, 1) + shift[:13].double().view(1, -1, 1, 1), 'cout 13 fp32')


DUAL_SHAPES = [  # n, cin, cout, hw, c2 (None = identity shortcut)        kernel path
    (2, 64, 256, 19, None),     # identity, N tile 256, 1 + 4 K blocks (stage-2 conv3)
    (1, 128, 512, 30, None),    # identity, 2 + 4 blocks, two N tiles
    (2, 256, 1024, 19, None),   # identity, 4 + 4 blocks (stage-4 conv3)
    (1, 512, 2048, 19, None),   # identity, N tile 128, K-chunked (stage-5 conv3)
    (1, 320, 384, 17, None),    # identity, N tile 128 (cout % 256 != 0), 5 + 2 blocks in one accumulator
    (3, 64, 128, 11, None),     # identity, N tile 128, rows off the tile
    (2, 64, 256, 19, 64),       # concat: conv3 + conv4 of stage2_0 (K = 64 + 64)
    (1, 128, 512, 24, 256),     # concat, 6 K blocks in one accumulator (stage3_0)
    (1, 512, 2048, 10, 1024),   # concat, K-chunked (stage5_0)
]


@pytest.mark.parametrize('n,cin,cout,hw,c2', DUAL_SHAPES)
def test_pair_conv_second_k_source(n, cin, cout, hw, c2):
    """Second K source of the 1x1 pair conv (ppy_conv_params.x2): the bottleneck's `conv3(y) + shortcut` -- identity shortcut
    added by the tensor core through identity weight blocks, or the projection shortcut conv4 K-concatenated -- against fp64."""
    o = ops()
    g = torch.Generator().manual_seed(cin * 13 + cout + hw + (c2 or 0))
    x = torch.randn((n, cin, hw, hw), generator=g)
    w1 = torch.randn((cout, cin, 1, 1), generator=g) / cin ** 0.5
    s1, b1 = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    y1 = F.conv2d(x.double(), w1.double()) * s1.double().view(1, -1, 1, 1) + b1.double().view(1, -1, 1, 1)
    if c2 is None:
        r = torch.randn((n, cout, hw, hw), generator=g) * 2.0
        want = torch.relu(y1 + r.double())
        got = o.conv_pair_dual(to_pair(x), w1.to(DEV), s1.to(DEV), b1.to(DEV), to_pair(r), act=1)
    else:
        x2 = torch.randn((n, c2, hw, hw), generator=g)
        w2 = torch.randn((cout, c2, 1, 1), generator=g) / c2 ** 0.5
        s2, b2 = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
        want = torch.relu(y1 + F.conv2d(x2.double(), w2.double()) * s2.double().view(1, -1, 1, 1) + b2.double().view(1, -1, 1, 1))
        got = o.conv_pair_dual(to_pair(x), w1.to(DEV), s1.to(DEV), b1.to(DEV), to_pair(x2), w2.to(DEV), s2.to(DEV), b2.to(DEV), act=1)
    check(from_pair(got, cout), want, 'second K source %s' % ((n, cin, cout, hw, c2),))


@pytest.mark.parametrize('n,cin,cout,k,hw', [(2, 128, 128, 1, 19),     # tma_a, one K chunk
                                              (1, 512, 256, 1, 38),     # tma_a, K-chunked
                                              (3, 64, 96, 1, 7),        # image smaller than a tile: rows wrap several times
                                              (2, 128, 256, 3, 32),     # patch loader (4-D box of the coordinate source)
                                              (2, 256, 512, 3, 19),     # im2col loader
                                              (1, 64, 256, 3, 5)])      # im2col, tiny map
def test_pair_conv_coordconv_as_k_block(n, cin, cout, k, hw):
    """CoordConv (model/custom_layers.py:256-272) + k x k conv on the pair path: the two coordinate channels are one extra K block
    from a batch-invariant second source; against fp64 of conv(cat([x, xs, ys]))."""
    o = ops()
    g = torch.Generator().manual_seed(cin + cout * 3 + k * 7 + hw)
    x = torch.randn((n, cin, hw, hw), generator=g)
    w = torch.randn((cout, cin + 2, k, k), generator=g) / (cin * k * k) ** 0.5
    w[:, cin:] *= 3.0                       # coordinate weights larger than the rest: they share the row scaling
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    xs = torch.arange(hw, dtype=torch.float32) / (hw - 1) * 2.0 - 1
    xc = torch.cat([x, xs.view(1, 1, 1, hw).expand(n, 1, hw, hw), xs.view(1, 1, hw, 1).expand(n, 1, hw, hw)], 1)
    want = act64(F.conv2d(xc.double(), w.double(), padding=(k - 1) // 2) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1), 2)
    y = o.conv_pair(to_pair(x), w.to(DEV), scale.to(DEV), shift.to(DEV), 1, (k - 1) // 2, 2, c_count=cin, coord=True)
    check(from_pair(y, cout), want, 'coordconv as K block %s' % ((n, cin, cout, k, hw),))


def test_pair_conv_identity_blocks_small_weights():
    """Identity blocks with channel weight scales spread over 2^-12 .. 2^6: chan_scale is capped at 2^15 (the diagonal must be an
    fp16 number); a channel whose weights are too small for the cap makes the packer decline (the engine then keeps the
    epilogue residual)."""
    o = ops()
    g = torch.Generator().manual_seed(21)
    n, cin, cout, hw = 1, 64, 256, 16
    x = torch.randn((n, cin, hw, hw), generator=g)
    mag = torch.ldexp(torch.ones(cout), torch.randint(-12, 7, (cout,), generator=g))
    w1 = torch.randn((cout, cin, 1, 1), generator=g) / 8.0 * mag.view(-1, 1, 1, 1)
    one, zero = torch.ones(cout), torch.zeros(cout)
    r = torch.randn((n, cout, hw, hw), generator=g)
    want = F.conv2d(x.double(), w1.double()) + r.double()
    got = o.conv_pair_dual(to_pair(x), w1.to(DEV), one.to(DEV), zero.to(DEV), to_pair(r), act=0)
    check(from_pair(got, cout), want, 'identity blocks, spread weight scales')
    assert o.pack_weight_pair_dual((w1 * 1e-7).to(DEV), one.to(DEV), ident_bn=256) is None


def test_pair_conv_small_and_large_magnitudes():
    """Activations of scale 1e-3 and 1e3, weights of scale 1e-4: the per-channel weight scaling keeps hi and lo parts of the
    weights normal at any magnitude.  Activations of scale 1e-3 have their lo parts in the fp16 subnormal range (absolute floor
    2^-25): precision degrades gracefully to ~2e-5 of the output scale (the engine stores activations times act_scale = 8 to stay
    clear of this; tolerance 5e-5 for that case, 1e-5 otherwise)."""
    g = torch.Generator().manual_seed(9)
    for xs, ws in ((1e-3, 1.0), (1e3, 1.0), (1.0, 1e-4), (30.0, 1e2)):
        x = torch.randn((1, 256, 16, 16), generator=g) * xs
        w = torch.randn((128, 256, 3, 3), generator=g) / 48.0 * ws
        one, zero = torch.ones(128), torch.zeros(128)
        want = F.conv2d(x.double(), w.double(), padding=1)
        check(run_pair_conv(x, w, one, zero, 1, 0, out_f32=True), want, 'magnitudes x*%g w*%g' % (xs, ws), tol=5e-5 if xs < 0.01 else TOL)


def test_pair_conv_overflow_flag():
    g = torch.Generator().manual_seed(11)
    x = torch.randn((1, 64, 8, 8), generator=g)
    w = torch.randn((64, 64, 1, 1), generator=g)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    run_pair_conv(x, w, torch.ones(64), torch.zeros(64), 1, 0, overflow=flag)
    assert int(flag.item()) == 0
    run_pair_conv(x, w, torch.full((64,), 1e5), torch.zeros(64), 1, 0, overflow=flag)
    assert int(flag.item()) == 1


@pytest.mark.parametrize('n,c,cout,stride,h,w,off_scale', [(1, 64, 64, 1, 13, 21, 0.03), (2, 128, 96, 1, 10, 17, 0.1),
                                                           (1, 256, 256, 2, 20, 20, 0.3), (1, 512, 512, 1, 19, 19, 0.05),
                                                           (2, 128, 128, 1, 38, 38, 1.0)])
def test_pair_dcn(n, c, cout, stride, h, w, off_scale):
    """Fused DCNv2 of the pair path (offset conv -> fp32 offsets, bilinear producer on joined corners, pair A tiles) vs the fp32
    CPU oracle given the kernel's own offsets (so the comparison isolates the sampling + GEMM)."""
    o = ops()
    g = torch.Generator().manual_seed(c * 7 + h * 3 + w + stride)
    x = torch.randn((n, c, h, w), generator=g)
    ow = torch.randn((27, c, 3, 3), generator=g) * off_scale / (c * 9) ** 0.5 * 10
    ob = torch.randn(27, generator=g)
    wt = torch.randn((cout, c, 3, 3), generator=g) / (c * 9) ** 0.5
    want = ref.dcnv2(x, ow, ob, wt, stride, 1)
    xp = to_pair(x)
    om = o.conv_pair(xp, ow.to(DEV), torch.ones(27), ob, stride, 1, 0, out_f32=True)
    om_ref = F.conv2d(x.double(), ow.double(), ob.double(), stride=stride, padding=1)
    check(o.from_nhwc(om, 27).cpu(), om_ref, 'offset conv')
    y = o.conv_pair(xp, wt.to(DEV), torch.ones(cout), torch.zeros(cout), stride, 1, 0, out_f32=True, offset_mask=om)
    got = o.from_nhwc(y, cout).cpu()
    # offsets differ from the oracle's by fp32 rounding (~1e-6 px) -> sample values by ~1e-6 of the local gradient
    check(got, want.double(), 'dcn', tol=5e-5)


@pytest.mark.parametrize('h,w,c', [(13, 17, 16), (12, 18, 64), (19, 19, 32)])
def test_pair_pools_spp(h, w, c):
    from ppyolo_b200._lib import lib, check as chk
    o = ops()
    g = torch.Generator().manual_seed(h * 100 + w)
    x = torch.randn((2, c, h, w), generator=g)
    xp = to_pair(x)
    xj = from_pair(xp, c)                      # the values the pair carries (fp32-exact)
    pl = xp.stride(0)
    for fn, want in ((lib.ppy_maxpool3x3s2_f16x2, F.max_pool2d(xj, 3, 2, 1)), (lib.ppy_avgpool2x2_f16x2, F.avg_pool2d(xj, 2, 2))):
        ho, wo = want.shape[2:]
        y = torch.zeros((2, 2, ho, wo, c), dtype=torch.float16, device=DEV)
        chk(fn(o.ptr(xp), c, pl, o.ptr(y), c, y.stride(0), 2, h, w, c, o.stream_ptr()), 'pool')
        got = from_pair(y, c)
        if fn is lib.ppy_maxpool3x3s2_f16x2:
            np.testing.assert_array_equal(got.numpy(), want.numpy())
        else:
            np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-6)
    y = torch.zeros((2, 2, h, w, 4 * c), dtype=torch.float16, device=DEV)
    chk(lib.ppy_spp_f16x2(o.ptr(xp), c, pl, o.ptr(y), 4 * c, y.stride(0), 2, h, w, c, o.stream_ptr()), 'spp')
    want = torch.cat([xj] + [F.max_pool2d(xj, k, 1, k // 2) for k in (5, 9, 13)], 1)
    np.testing.assert_array_equal(from_pair(y, 4 * c).numpy(), want.numpy())


def test_pair_stem():
    from ppyolo_b200._lib import lib, check as chk
    import ctypes
    o = ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 3, 64, 96), generator=g)
    w = torch.randn((32, 3, 3, 3), generator=g) * 0.3
    sc, sh = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    want = torch.relu(F.conv2d(x.double(), w.double(), stride=2, padding=1) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    y = torch.zeros((2, 2, 32, 48, 32), dtype=torch.float16, device=DEV)
    fp = ctypes.POINTER(ctypes.c_float)
    wn, scn, shn = [np.ascontiguousarray(t.numpy()) for t in (w, sc, sh)]
    chk(lib.ppy_stem_conv3x3s2_f16x2(o.ptr(x.to(DEV)), 2, 64, 96, wn.ctypes.data_as(fp), scn.ctypes.data_as(fp), shn.ctypes.data_as(fp),
                                     32, 1, o.ptr(y), 32, y.stride(0), o.stream_ptr()), 'stem')
    check(from_pair(y, 32), want, 'stem pair')


@pytest.mark.parametrize('mode', ['bf16', 'f16x2'])
@pytest.mark.parametrize('n,c,cout,stride,hw,res', [(4, 512, 512, 1, 19, True), (3, 512, 512, 2, 38, False), (2, 128, 256, 1, 13, True),
                                                    (1, 64, 512, 1, 40, False)])
def test_whole_layer_dcn_kernel(mode, n, c, cout, stride, hw, res):
    """dcn_umma.cu (CTA pairs, full-N accumulators, coalesced corner gather, sampling table) at the shapes of the ppyolo_2x
    stage-5 DCNs (512 -> 512 at 19x19 stride 1 and 38 -> 19 stride 2) and smaller ones, with residual + ReLU, against the
    fp32 CPU oracle fed the kernel's own offsets.  bf16: operands rounded to bf16 first, 4e-3 of scale (the A operand is rounded
    to bf16 after the blend); f16x2: 5e-5 of scale."""
    from ppyolo_b200._lib import PPY_F32, PPY_BF16
    o = ops()
    g = torch.Generator().manual_seed(c + cout + hw + stride)
    rnd = (lambda t: t.to(torch.bfloat16).float()) if mode == 'bf16' else (lambda t: t)
    x = rnd(torch.randn((n, c, hw, hw), generator=g))
    ow = rnd(torch.randn((27, c, 3, 3), generator=g) * 1.5 / (c * 9) ** 0.5)
    ob = torch.randn(27, generator=g)
    wt = rnd(torch.randn((cout, c, 3, 3), generator=g) / (c * 9) ** 0.5)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ho = (hw + 2 - 3) // stride + 1
    r = rnd(torch.randn((n, cout, ho, ho), generator=g)) if res else None
    want = ref.dcnv2(x, ow, ob, wt, stride, 1).double() * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    if res:
        want = want + r.double()
    want = torch.relu(want)
    if mode == 'f16x2':
        xp = to_pair(x)
        om = o.conv_pair(xp, ow.to(DEV), torch.ones(27), ob, stride, 1, 0, out_f32=True)
        y = o.conv_pair(xp, wt.to(DEV), scale, shift, stride, 1, 1, residual=to_pair(r) if res else None, offset_mask=om)
        got = from_pair(y, cout)
        tol = 5e-5
    else:
        xh = o.to_nhwc(x.to(DEV), PPY_BF16)
        om = o.conv_nhwc(xh, o.pack_weight(ow.to(DEV), PPY_BF16), c, 27, 3, stride, 1, torch.ones(27, device=DEV), ob.to(DEV), 0, PPY_BF16,
                         out_code=PPY_F32)
        y = o.conv_nhwc(xh, o.pack_weight(wt.to(DEV), PPY_BF16), c, cout, 3, stride, 1, scale.to(DEV), shift.to(DEV), 1, PPY_BF16,
                        residual=o.to_nhwc(r.to(DEV), PPY_BF16) if res else None, offset_mask=om)
        got = o.from_nhwc(y, cout).cpu()
        tol = 8e-3                       # output rounded to bf16 as well
    check(got, want, 'whole-layer dcn %s %s' % (mode, (n, c, cout, stride, hw, res)), tol=tol)

"""GPU parity tests (B200): every kernel through the C ABI against the golden fixtures of the reference and
against the CPU oracle on seeded inputs.

Tolerances (stated per the north_star):
  * Matrix-NMS: labels and row order bit-exact; linear-kernel scores exact to 1 ulp (rtol 2e-7); gaussian
    kernel rtol 2e-6 (expf ulp).
  * decode: boxes rtol 1e-5 + atol 1e-4 px, scores rtol 1e-5 (libm ulp of exp/log/pow).
  * fp32 SIMT convs / DCN: 1e-4 relative to the output scale (fp32 re-association only).
  * bf16 tcgen05 convs: compared with the oracle evaluated on the SAME bf16-rounded inputs/weights, fp32
    output: 2e-3 relative to the output scale (fp32 accumulation order; bilinear blend rounded to bf16 for DCN).
"""
import numpy as np
import pytest
import torch

from oracle import ppyolo_ref as ref
from ppyolo_b200 import synth
from tests.helpers import build_model, assert_preds_close

pytestmark = pytest.mark.gpu
NMS_CFG = dict(score_threshold=0.01, post_threshold=0.01, nms_top_k=500, keep_top_k=100)
DEV = 'cuda'


def ops():
    from ppyolo_b200 import ops as _ops
    return _ops


def scale_of(t):
    return float(np.abs(t).max()) + 1e-12


# ------------------------------------------------------------------ Matrix-NMS
@pytest.mark.parametrize('name', ['zeros', 'dups', 'chain', 'degenerate'])
def test_nms_known_answers(golden, name):
    from model.matrix_nms import matrix_nms
    z = golden('nms')
    for gauss in ((False,) if name == 'degenerate' else (False, True)):
        tag = name if name == 'degenerate' else '%s_%s' % (name, 'g' if gauss else 'l')
        out = matrix_nms(torch.from_numpy(z[tag + '_boxes']).to(DEV), torch.from_numpy(z[tag + '_scores']).to(DEV),
                         use_gaussian=gauss, **NMS_CFG)
        assert_preds_close(out.cpu().numpy(), z[tag + '_out'], rtol=2e-6 if gauss else 2e-7, atol=0)


@pytest.mark.parametrize('nb,nc,seed,topk,keep', [(400, 80, 0, 500, 100), (400, 80, 1, 100, 20), (3000, 80, 2, 500, 100),
                                                  (64, 3, 3, -1, 100), (1500, 20, 4, 300, 50)])
@pytest.mark.parametrize('gauss', [False, True])
def test_nms_golden_random(golden, nb, nc, seed, topk, keep, gauss):
    from model.matrix_nms import matrix_nms
    z = golden('nms')
    b, s = synth.nms_inputs(nb, nc, seed=seed)
    out = matrix_nms(b.to(DEV), s.to(DEV), 0.01, 0.01, topk, keep, use_gaussian=gauss, gaussian_sigma=2.0)
    want = z['rand_b%d_c%d_s%d_t%d_k%d_%s_out' % (nb, nc, seed, topk, keep, 'g' if gauss else 'l')]
    assert_preds_close(out.cpu().numpy(), want, rtol=2e-6 if gauss else 2e-7, atol=0)


def test_nms_post_threshold(golden):
    from model.matrix_nms import matrix_nms
    z = golden('nms')
    b, s = synth.nms_inputs(800, 80, seed=5)
    out = matrix_nms(b.to(DEV), s.to(DEV), 0.05, 0.2, 200, 30)
    assert_preds_close(out.cpu().numpy(), z['post05_out'], rtol=2e-7, atol=0)


def test_nms_c5_batched_vs_oracle():
    """BASELINE config C5 (10k boxes x 80 classes), batch of 4 different images in one launch sequence."""
    o = ops()
    pairs = [synth.nms_inputs(10000, 80, seed=100 + i) for i in range(4)]
    boxes = torch.stack([p[0] for p in pairs]).to(DEV)
    scores = torch.stack([p[1] for p in pairs]).to(DEV)
    for gauss in (False, True):
        got = o.matrix_nms_batched(boxes, scores, use_gaussian=gauss, **NMS_CFG)
        for i in range(4):
            want = ref.matrix_nms(pairs[i][0].numpy(), pairs[i][1].numpy(), use_gaussian=gauss, **NMS_CFG)
            assert_preds_close(got[i].cpu().numpy(), want, rtol=2e-6 if gauss else 2e-7, atol=0)


@pytest.mark.parametrize('nb,topk,gauss', [(400, 3000, False), (400, 2000, True), (200, -1, False), (330, 4000, False)])
def test_nms_beyond_1024_candidates_vs_oracle(nb, topk, gauss):
    """More than 1024 boxes in the n x n stage (per-box arrays in dynamic shared memory, up to 4000) and nms_top_k = -1 ("all
    candidates", model/matrix_nms.py:120-125): labels / order exact, scores 2e-7 (linear) / 2e-6 (gaussian) against the oracle."""
    from model.matrix_nms import matrix_nms
    b, s = synth.nms_inputs(nb, 80, seed=nb + 7)
    cand = int((s > 0.01).sum())
    assert cand > 1024 and (topk > 0 or cand <= 4000)
    out = matrix_nms(b.to(DEV), s.to(DEV), 0.01, 0.01, topk, 100, use_gaussian=gauss, gaussian_sigma=2.0)
    want = ref.matrix_nms(b.numpy(), s.numpy(), 0.01, 0.01, topk, 100, use_gaussian=gauss, gaussian_sigma=2.0)
    assert_preds_close(out.cpu().numpy(), want, rtol=2e-6 if gauss else 2e-7, atol=0)


def test_nms_properties_full_size():
    """Size-independent properties at the bench size (22743 boxes x 80, bs 8): sorted scores, idempotent
    re-run, labels in range, all boxes are input boxes."""
    o = ops()
    g = torch.Generator().manual_seed(5)
    n, nb = 8, 22743
    cxy = torch.rand((n, nb, 2), generator=g) * 608
    wh = torch.rand((n, nb, 2), generator=g) * 150 + 4
    boxes = torch.cat([cxy - wh / 2, cxy + wh / 2], -1).to(DEV)
    scores = (torch.sigmoid(torch.randn((n, nb, 1), generator=g) * 2 - 4) *
              torch.sigmoid(torch.randn((n, nb, 80), generator=g) * 2 - 3)).to(DEV)
    a = o.matrix_nms_batched(boxes, scores, **NMS_CFG)
    b = o.matrix_nms_batched(boxes, scores, **NMS_CFG)
    for i in range(n):
        assert torch.equal(a[i], b[i])
        p = a[i].cpu().numpy()
        assert p.shape == (100, 6)
        assert (np.diff(p[:, 1]) <= 0).all() and (p[:, 1] >= 0.01).all()
        assert ((p[:, 0] >= 0) & (p[:, 0] < 80) & (p[:, 0] == np.round(p[:, 0]))).all()
    # image 0 against the oracle as well
    want = ref.matrix_nms(boxes[0].cpu().numpy(), scores[0].cpu().numpy(), **NMS_CFG)
    assert_preds_close(a[0].cpu().numpy(), want, rtol=2e-7, atol=0)


def test_jaccard(golden):
    from model.matrix_nms import jaccard
    z = golden('nms')
    ba, _ = synth.nms_inputs(37, 1, seed=7)
    bb, _ = synth.nms_inputs(53, 1, seed=8)
    np.testing.assert_allclose(jaccard(ba.to(DEV), bb.to(DEV)).cpu().numpy(), z['jaccard_out'], rtol=2e-7, atol=0)


# ------------------------------------------------------------------ decode
@pytest.mark.parametrize('tag,stride,mask,iou_aware', [('s32', 32, [6, 7, 8], True), ('s8', 8, [0, 1, 2], True),
                                                       ('plain', 16, [3, 4, 5], False)])
def test_decode_golden(golden, tag, stride, mask, iou_aware):
    from model import head as H
    z = golden('decode')
    anchors = np.array(build_model('r50vd')[1].head['anchors'], np.float32)[mask]
    x = torch.from_numpy(z[tag + '_in']).to(DEV)
    im_size = torch.from_numpy(z[tag + '_im_size']).to(DEV)
    if iou_aware:
        y = H.get_iou_aware_score(x, 3, 80, 0.4)
        np.testing.assert_allclose(y.cpu().numpy(), z[tag + '_iouaware'], rtol=1e-5, atol=1e-5)
    else:
        y = x
    for clip in (True, False):
        boxes, scores = H.yolo_box(y, anchors, stride, 80, 1.05, im_size, clip, 0.01)
        want = z['%s_boxes_clip%d' % (tag, int(clip))]
        got = boxes.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(scores.cpu().numpy(), z[tag + '_scores'], rtol=1e-5, atol=1e-9)
    if iou_aware:   # fused path: iou-aware inside the decode kernel
        boxes, scores = ops().yolo_box(x, anchors, stride, 80, 1.05, im_size, True, iou_aware=True, iou_aware_factor=0.4)
        np.testing.assert_allclose(boxes.cpu().numpy(), z[tag + '_boxes_clip1'], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(scores.cpu().numpy(), z[tag + '_scores'], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize('n,size,an,nc,stride,sxy,iou_aware', [
    (1, 1, 3, 80, 32, 1.05, True), (3, 7, 3, 80, 32, 1.05, True), (2, 19, 3, 80, 32, 1.05, False), (5, 13, 3, 20, 16, 1.0, True),
    (2, 40, 2, 20, 8, 1.05, True), (1, 76, 3, 80, 8, 1.05, True), (4, 10, 4, 1, 32, 1.2, False), (2, 26, 1, 90, 16, 1.05, True)])
def test_decode_sweep_vs_oracle(n, size, an, nc, stride, sxy, iou_aware):
    """Fused (iou-aware +) yolo_box against the oracle over grid sizes, batch sizes, anchor counts and class counts the
    goldens do not hold (the generic kernel instantiation included), with per-image im_size and both clip settings."""
    g = torch.Generator().manual_seed(size * 1000 + nc * 10 + an)
    ch = an * (nc + (6 if iou_aware else 5))
    x = torch.randn((n, ch, size, size), generator=g) * 2.0
    anchors = (torch.rand((an, 2), generator=g) * 300 + 8).numpy().astype(np.float32)
    im_size = torch.stack([torch.randint(200, 900, (n,), generator=g), torch.randint(200, 900, (n,), generator=g)], 1).float()
    y = ref.iou_aware_score(x, an, nc, 0.4) if iou_aware else x
    for clip in (True, False):
        wb, ws = ref.yolo_box(y, anchors, stride, nc, sxy, im_size, clip)
        gb, gs = ops().yolo_box(x.to(DEV), anchors, stride, nc, sxy, im_size.to(DEV), clip, iou_aware=iou_aware,
                                iou_aware_factor=0.4)
        np.testing.assert_allclose(gb.cpu().numpy(), wb.numpy(), rtol=1e-5, atol=2e-4)
        np.testing.assert_allclose(gs.cpu().numpy(), ws.numpy(), rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize('obj_bias,gauss', [(-4.0, False), (0.5, False), (-2.0, True)])
def test_sparse_decode_nms_equals_dense(obj_bias, gauss):
    """ppy_yolo_decode_candidates + ppy_matrix_nms_candidates == ppy_yolo_decode + ppy_matrix_nms_batched, bit for bit,
    at the three ppyolo_2x scales of a 320x320 input (sparse objectness, and a dense-objectness stress case where most
    anchors pass the threshold)."""
    o = ops()
    from ppyolo_b200._lib import PPY_F32
    n, nc = 3, 80
    anchors = np.array(build_model('r50vd')[1].head['anchors'], np.float32)
    g = torch.Generator().manual_seed(11)
    sizes, strides, masks = (10, 20, 40), (32, 16, 8), ([6, 7, 8], [3, 4, 5], [0, 1, 2])
    total = sum(s * s * 3 for s in sizes)
    im_size = synth.im_sizes(n).to(DEV)
    boxes_d = torch.zeros((n, total, 4), device=DEV)
    scores_d = torch.zeros((n, total, nc), device=DEV)
    boxes_s = torch.zeros((n, total, 4), device=DEV)
    cap = total * nc
    ws = o.nms_candidate_workspace(n, cap, torch.device(DEV))
    o.nms_candidates_reset(ws, n, cap)
    off = 0
    for s, st, mk in zip(sizes, strides, masks):
        x = torch.randn((n, 258, s, s), generator=g) * 1.5
        x[:, 3:] += -3.0
        for a in range(3):
            x[:, 3 + a * 85 + 4] += obj_bias + 2.0
        xh = o.to_nhwc(x.to(DEV), PPY_F32, 264)
        o.yolo_decode_nhwc(xh, 264, n, s, anchors[mk], st, nc, 1.05, im_size, True, True, 0.4, boxes_d, scores_d, off, total)
        o.yolo_decode_candidates_nhwc(xh, 264, n, s, anchors[mk], st, nc, 1.05, im_size, True, True, 0.4, boxes_s, off, total,
                                      0.01, ws, cap)
        off += s * s * 3
    np.testing.assert_array_equal(boxes_d.cpu().numpy(), boxes_s.cpu().numpy())
    n_cand = int((scores_d > 0.01).sum())
    assert n_cand > 100
    out_d = torch.zeros((n, 100, 6), device=DEV); cnt_d = torch.zeros((n,), dtype=torch.int32, device=DEV)
    out_s = torch.zeros((n, 100, 6), device=DEV); cnt_s = torch.zeros((n,), dtype=torch.int32, device=DEV)
    o.matrix_nms_launch(boxes_d, scores_d, out_d, cnt_d, o.nms_workspace(n, total, nc, torch.device(DEV)), 0.01, 0.01, 500, 100,
                        gauss, 2.0)
    o.matrix_nms_candidates_launch(boxes_s, nc, out_s, cnt_s, ws, cap, 0.01, 0.01, 500, 100, gauss, 2.0)
    cd, cs = cnt_d.cpu().numpy(), cnt_s.cpu().numpy()
    print('candidates > 0.01: %d, kept per image: %s' % (n_cand, cd.tolist()))
    np.testing.assert_array_equal(cd, cs)
    assert cd.min() > 0
    for i in range(n):
        np.testing.assert_array_equal(out_d[i, :cd[i]].cpu().numpy(), out_s[i, :cs[i]].cpu().numpy())
    # and against the CPU oracle for one image (dense scores -> reference matrix_nms)
    want = ref.matrix_nms(boxes_d[0].cpu().numpy(), scores_d[0].cpu().numpy(), 0.01, 0.01, 500, 100, use_gaussian=gauss)
    got = out_s[0, :cs[0]].cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got[:, 0], want[:, 0])
    np.testing.assert_allclose(got, want, rtol=2e-6 if gauss else 2e-7, atol=0)


# ------------------------------------------------------------------ glue layers
def test_coord_spp_pool(golden):
    from model.custom_layers import CoordConv, SPP
    o = ops()
    z = golden('layers')
    np.testing.assert_allclose(CoordConv(True)(torch.zeros(1, 2, 3, 4, device=DEV)).cpu().numpy(), z['coord_out'], atol=1e-7)
    x = torch.from_numpy(z['spp_in'])
    x8 = torch.cat([x, x * 0.5], 1)                      # 8 channels (vector width)
    got = SPP()(x8.to(DEV)).cpu()
    np.testing.assert_array_equal(got.numpy(), ref.spp(x8).numpy())
    x64 = torch.randn((3, 64, 19, 19), generator=torch.Generator().manual_seed(4))      # separable shared-memory kernel
    np.testing.assert_array_equal(SPP()(x64.to(DEV)).cpu().numpy(), ref.spp(x64).numpy())
    from ppyolo_b200._lib import PPY_BF16, lib, check
    xb = o.to_nhwc(x64.to(DEV), PPY_BF16)
    yb = torch.empty((3, 19, 19, 256), dtype=torch.bfloat16, device=DEV)
    check(lib.ppy_spp(o.ptr(xb), 64, o.ptr(yb), 256, 3, 19, 19, 64, PPY_BF16, o.stream_ptr()), 'spp')
    np.testing.assert_array_equal(o.from_nhwc(yb, 256).cpu().numpy(), ref.spp(x64.to(torch.bfloat16).float()).numpy())
    g = torch.Generator().manual_seed(2)
    t = torch.randn((2, 16, 13, 17), generator=g)
    np.testing.assert_array_equal(o.max_pool3s2(t.to(DEV)).cpu().numpy(), torch.nn.functional.max_pool2d(t, 3, 2, 1).numpy())
    t = torch.randn((2, 16, 12, 18), generator=g)
    np.testing.assert_allclose(o.avg_pool2(t.to(DEV)).cpu().numpy(), torch.nn.functional.avg_pool2d(t, 2, 2).numpy(),
                               rtol=1e-6, atol=1e-7)
    r, f = torch.randn((2, 8, 5, 5), generator=g), torch.randn((2, 16, 10, 10), generator=g)
    want = torch.cat([torch.nn.functional.interpolate(r, scale_factor=2, mode='nearest'), f], 1)
    np.testing.assert_array_equal(o.upsample2x_concat(r.to(DEV), f.to(DEV)).cpu().numpy(), want.numpy())


# ------------------------------------------------------------------ conv units (golden from the reference)
UNIT_CASES = (('c3s1', 8, 16, 3, 1, 'leaky', False), ('c3s2', 8, 16, 3, 2, 'relu', False),
              ('c1s1', 16, 24, 1, 1, None, True), ('c1s2', 8, 8, 1, 2, 'relu', False))


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_conv_units_golden(golden, precision):
    from model.custom_layers import Conv2dUnit
    o = ops()
    z = golden('layers')
    o.set_precision(precision)
    try:
        for tag, cin, cout, k, stride, act, bias in UNIT_CASES:
            u = Conv2dUnit(cin, cout, k, stride=stride, bias_attr=bias, bn=0 if bias else 1, act=act)
            synth.randomize_(u, seed=30)
            u = u.to(DEV).eval()
            y = u(torch.from_numpy(z[tag + '_in']).to(DEV)).cpu().numpy()
            want = z[tag + '_out']
            tol = 1e-4 if precision == 'fp32' else 3e-2      # bf16: inputs, weights AND output rounded to bf16
            np.testing.assert_allclose(y, want, rtol=0, atol=tol * scale_of(want), err_msg=tag)
    finally:
        o.set_precision('fp32')


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_conv_unit_mish(precision):
    """Conv2dUnit(act='mish') (reference model/custom_layers.py:37-43, :131-132): conv + folded BN + x*tanh(softplus(x)) against
    torch fp32 -- in the SIMT epilogue for fp32, conv kernel + ppy_activation pass on the tcgen05 path.  1e-4 / 3e-2 of scale."""
    from model.custom_layers import Conv2dUnit
    o = ops()
    o.set_precision(precision)
    try:
        u = Conv2dUnit(64, 96, 3, stride=1, bn=1, act='mish')
        synth.randomize_(u, seed=35)
        u = u.to(DEV).eval()
        x = torch.randn((2, 64, 12, 12), generator=torch.Generator().manual_seed(5)).to(DEV)
        y = u(x)
        c = lambda t: t.detach().cpu()                  # reference on the CPU: cuDNN would use TF32
        t = torch.nn.functional.conv2d(c(x), c(u.conv.weight), None, 1, 1)
        t = torch.nn.functional.batch_norm(t, c(u.bn.running_mean), c(u.bn.running_var), c(u.bn.weight), c(u.bn.bias), False, 0.1, u.bn.eps)
        want = t * torch.tanh(torch.nn.functional.softplus(t))
        tol = 1e-4 if precision == 'fp32' else 3e-2
        np.testing.assert_allclose(y.cpu().numpy(), want.detach().cpu().numpy(), rtol=0, atol=tol * scale_of(want.detach().cpu().numpy()))
    finally:
        o.set_precision('fp32')


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('tag,stride', [('dcn_s1', 1), ('dcn_s2', 2)])
def test_dcn_golden(golden, precision, tag, stride):
    from model.custom_layers import Conv2dUnit
    o = ops()
    z = golden('layers')
    u = Conv2dUnit(16, 24, 3, stride=stride, bn=1, act='relu', use_dcn=True)
    synth.randomize_(u, seed=31, offset_scale=0.05)
    u = u.to(DEV).eval()
    x = torch.from_numpy(z[tag + '_in']).to(DEV)
    if precision == 'fp32':
        raw = u.conv(x).cpu().numpy()
        np.testing.assert_allclose(raw, z[tag + '_raw'], rtol=0, atol=1e-4 * scale_of(z[tag + '_raw']))
        y = u(x).cpu().numpy()
        np.testing.assert_allclose(y, z[tag + '_out'], rtol=0, atol=1e-4 * scale_of(z[tag + '_out']))
    else:
        pytest.skip('bf16 DCN needs cin % 64 == 0; covered by test_umma_dcn')


@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'f16x2'])
@pytest.mark.parametrize('tag,stride', [('s1', 1), ('s2', 2)])
def test_dcn_far_offsets_golden(golden, precision, tag, stride):
    """DCNv2 with offsets up to ~24 px on a 20x20 map (most samples far outside the image) against the REFERENCE's output
    (tests/golden/dcn_far.npz): fp32 SIMT kernel and f16x2 tensor-core kernel at 1e-4 of scale (measured 1e-5 / 1.6e-5); bf16
    recorded with a loose 1e-1 bound (measured 4-5e-2: a 20 px offset computed from bf16 operands is off by ~0.05 px, and a
    sample that moves by that much near the border changes visibly -- bf16 is the throughput mode, not the parity mode)."""
    from model.custom_layers import Conv2dUnit
    from ppyolo_b200._lib import PPY_F32, PPY_BF16
    o = ops()
    z = golden('dcn_far')
    u = Conv2dUnit(64, 24, 3, stride=stride, bn=1, act='relu', use_dcn=True)
    synth.randomize_(u, seed=33, offset_scale=0.25)
    u = u.to(DEV).eval()
    x = torch.from_numpy(z[tag + '_in']).to(DEV)
    want = z[tag + '_raw']
    d = u.conv
    if precision == 'fp32':
        got = d(x).cpu().numpy()
        tol = 1e-4
    elif precision == 'f16x2':
        xp = o.split_pair(o.to_nhwc(x, PPY_F32))
        om = o.conv_pair(xp, d.conv_offset.weight, torch.ones(27), d.conv_offset.bias, stride, 1, 0, out_f32=True)
        y = o.conv_pair(xp, d.dcn_weight, torch.ones(24), torch.zeros(24), stride, 1, 0, out_f32=True, offset_mask=om)
        got = o.from_nhwc(y, 24).cpu().numpy()
        tol = 1e-4
    else:
        xh = o.to_nhwc(x, PPY_BF16)
        one = torch.ones(27, device=DEV)
        om = o.conv_nhwc(xh, o.pack_weight(d.conv_offset.weight, PPY_BF16), 64, 27, 3, stride, 1, one, d.conv_offset.bias.detach().float(), 0,
                         PPY_BF16, out_code=PPY_F32)
        y = o.conv_nhwc(xh, o.pack_weight(d.dcn_weight, PPY_BF16), 64, 24, 3, stride, 1, torch.ones(24, device=DEV), torch.zeros(24, device=DEV),
                        0, PPY_BF16, out_code=PPY_F32, offset_mask=om)
        got = o.from_nhwc(y, 24).cpu().numpy()
        tol = 1e-1
    err = np.abs(got - want).max() / scale_of(want)
    print('dcn far offsets %s %s: max err / scale = %.2e' % (precision, tag, err))
    assert err < tol


# ------------------------------------------------------------------ tcgen05 conv kernel, tight check
def _umma_conv(x, w, scale, shift, stride, act, residual=None, bias_map=None, upsample=False, out_f32=True,
               offset_mask=None):
    """Run ppy_conv_bf16 on NHWC bf16 buffers; returns NCHW fp32."""
    from ppyolo_b200._lib import PPY_F32, PPY_BF16
    o = ops()
    cout, cin, k, _ = w.shape
    packed = o.pack_weight(w.to(DEV), PPY_BF16)
    xh = o.to_nhwc(x.to(DEV), PPY_BF16, packed[1])
    out_code = PPY_F32 if out_f32 else PPY_BF16
    res = o.to_nhwc(residual.to(DEV), out_code) if residual is not None else None
    y = o.conv_nhwc(xh, packed, cin, cout, k, stride, (k - 1) // 2, scale.to(DEV), shift.to(DEV), act, PPY_BF16,
                    residual=res, bias_map=bias_map.to(DEV) if bias_map is not None else None, out_code=out_code,
                    upsample2x=upsample, offset_mask=offset_mask)
    return o.from_nhwc(y, cout).cpu()


def bf16_round(t):
    return t.to(torch.bfloat16).float()


UMMA_SHAPES = [  # (n, cin, cout, k, stride, hw)
    (2, 64, 64, 1, 1, 16), (2, 64, 128, 3, 1, 16), (1, 128, 256, 3, 2, 16), (2, 32, 32, 3, 1, 24),
    (2, 3, 32, 3, 2, 32), (1, 256, 258, 1, 1, 12), (1, 512, 27, 3, 1, 10), (3, 64, 512, 1, 1, 9),
    (1, 1024, 256, 1, 1, 19), (1, 256, 512, 3, 1, 19), (2, 64, 96, 3, 1, 7),
    (2, 64, 64, 3, 1, 30), (1, 128, 128, 3, 1, 46), (3, 64, 32, 3, 1, 16),      # 4-D TMA patch mode (16x8 pixel tiles)
    (2, 128, 128, 3, 2, 19), (3, 64, 64, 3, 2, 38), (2, 256, 256, 3, 1, 38), (5, 64, 64, 3, 1, 5),   # im2col-mode TMA (tile walks rows/images)
    (1, 64, 64, 3, 2, 11), (2, 128, 320, 3, 1, 13),
    (2, 128, 64, 3, 1, 24), (1, 192, 128, 3, 1, 40), (5, 64, 128, 3, 1, 16), (1, 64, 64, 3, 1, 8),      # slab mode (8x16 tiles, 3 taps per stage)
]


@pytest.mark.parametrize('n,cin,cout,k,stride,hw', UMMA_SHAPES)
def test_umma_conv_shapes(n, cin, cout, k, stride, hw):
    g = torch.Generator().manual_seed(cin * 131 + cout)
    x = bf16_round(torch.randn((n, cin, hw, hw), generator=g))
    w = bf16_round(torch.randn((cout, cin, k, k), generator=g) * (1.0 / (cin * k * k) ** 0.5))
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    want = torch.nn.functional.conv2d(x, w, None, stride, (k - 1) // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    want = torch.nn.functional.leaky_relu(want, 0.1)
    got = _umma_conv(x, w, scale, shift, stride, 2)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-4 * scale_of(want.numpy()))


def test_umma_conv_epilogue_variants():
    g = torch.Generator().manual_seed(9)
    n, cin, cout, hw = 2, 64, 64, 10
    x = bf16_round(torch.randn((n, cin, hw, hw), generator=g))
    w = bf16_round(torch.randn((cout, cin, 3, 3), generator=g) * 0.05)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    base = torch.nn.functional.conv2d(x, w, None, 1, 1)
    # residual + relu, bf16 output (the block epilogue of resnet_vd.py:54-56)
    res = bf16_round(torch.randn((n, cout, hw, hw), generator=g))
    want = torch.relu(base * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    got = _umma_conv(x, w, scale, shift, 1, 1, residual=res, out_f32=False)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=8e-3, atol=8e-3 * scale_of(want.numpy()))
    # CoordConv bias map + leaky, fp32 out
    bm = torch.randn((hw * hw, cout), generator=g)
    want = torch.nn.functional.leaky_relu((base + bm.t().reshape(1, cout, hw, hw)) * scale.view(1, -1, 1, 1) +
                                          shift.view(1, -1, 1, 1), 0.1)
    got = _umma_conv(x, w, scale, shift, 1, 2, bias_map=bm)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-4 * scale_of(want.numpy()))
    # fused nearest x2 upsample
    want = torch.nn.functional.interpolate(base * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), scale_factor=2)
    got = _umma_conv(x, w, scale, shift, 1, 0, upsample=True)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-4 * scale_of(want.numpy()))


@pytest.mark.parametrize('stride,hw', [(1, 19), (2, 20)])
def test_umma_dcn(stride, hw):
    """Fused DCNv2 on tensor cores (cin=cout=128) vs the oracle on bf16-rounded operands."""
    from ppyolo_b200._lib import PPY_F32, PPY_BF16
    o = ops()
    g = torch.Generator().manual_seed(77 + stride)
    n, c, cout = 2, 128, 128
    x = bf16_round(torch.randn((n, c, hw, hw), generator=g))
    ow = bf16_round(torch.randn((27, c, 3, 3), generator=g) * 0.03)
    ob = torch.randn(27, generator=g)
    w = bf16_round(torch.randn((cout, c, 3, 3), generator=g) * 0.03)
    want = ref.dcnv2(x, ow, ob, w, stride, 1)
    # run the two kernels like ops.dcnv2 does, but with fp32 output of the main GEMM
    xh = o.to_nhwc(x.to(DEV), PPY_BF16)
    om = o.conv_nhwc(xh, o.pack_weight(ow.to(DEV), PPY_BF16), c, 27, 3, stride, 1, torch.ones(27, device=DEV),
                     ob.to(DEV), 0, PPY_BF16, out_code=PPY_F32)
    y = o.conv_nhwc(xh, o.pack_weight(w.to(DEV), PPY_BF16), c, cout, 3, stride, 1, torch.ones(cout, device=DEV),
                    torch.zeros(cout, device=DEV), 0, PPY_BF16, out_code=PPY_F32, offset_mask=om)
    got = o.from_nhwc(y, cout).cpu()
    # A operand (modulated bilinear sample) is rounded to bf16 before the MMA: ~2^-9 relative per element
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=4e-3 * scale_of(want.numpy()))


@pytest.mark.parametrize('n,c,cout,stride,h,w,off_scale', [
    (1, 64, 64, 1, 13, 21, 0.03), (2, 192, 96, 1, 10, 17, 0.1), (1, 256, 256, 2, 20, 20, 0.3), (3, 64, 40, 2, 10, 14, 0.3),
    (1, 512, 512, 1, 19, 19, 0.05), (2, 128, 128, 1, 38, 38, 1.0)])
def test_umma_dcn_sweep(n, c, cout, stride, h, w, off_scale):
    """Fused DCNv2 over non-square maps, channel counts off the tile size and offset magnitudes from sub-pixel to far outside
    the image (offset sigma = 10 * off_scale px: at 1.0 most corners fall outside a 38x38 map and must contribute zero).
    Stride 2 only on even maps: the reference's output-size rule (custom_layers.py:567-568) equals the offset conv's own
    output size only there, and every input that is a multiple of 32 gives the stride-2 DCN block an even map."""
    from ppyolo_b200._lib import PPY_F32, PPY_BF16
    o = ops()
    g = torch.Generator().manual_seed(c * 7 + h * 3 + w + stride)
    x = bf16_round(torch.randn((n, c, h, w), generator=g))
    ow = bf16_round(torch.randn((27, c, 3, 3), generator=g) * off_scale / (c * 9) ** 0.5 * 10)
    ob = torch.randn(27, generator=g)
    wt = bf16_round(torch.randn((cout, c, 3, 3), generator=g) / (c * 9) ** 0.5)
    want = ref.dcnv2(x, ow, ob, wt, stride, 1)
    xh = o.to_nhwc(x.to(DEV), PPY_BF16)
    om = o.conv_nhwc(xh, o.pack_weight(ow.to(DEV), PPY_BF16), c, 27, 3, stride, 1, torch.ones(27, device=DEV),
                     ob.to(DEV), 0, PPY_BF16, out_code=PPY_F32)
    y = o.conv_nhwc(xh, o.pack_weight(wt.to(DEV), PPY_BF16), c, cout, 3, stride, 1, torch.ones(cout, device=DEV),
                    torch.zeros(cout, device=DEV), 0, PPY_BF16, out_code=PPY_F32, offset_mask=om)
    got = o.from_nhwc(y, cout).cpu()
    assert scale_of(want.numpy()) > 0.05
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=4e-3 * scale_of(want.numpy()))


@pytest.mark.parametrize('n,c,stride,h,w,code', [(2, 256, 1, 11, 14, 'bf16'), (1, 512, 1, 19, 19, 'bf16'), (2, 1024, 2, 10, 8, 'bf16'),
                                                 (1, 192, 1, 9, 9, 'bf16'), (1, 128, 1, 7, 9, 'fp32'), (2, 64, 2, 8, 6, 'fp32')])
def test_dcn_gather_stage(n, c, stride, h, w, code):
    """ppy_dcn_gather (the sampling stage of the two-kernel DCNv2 the engine runs): its [M x 9C] matrix times the dcn weight
    in (tap, c) order must equal the oracle's DCNv2 given the same offset/mask conv output.  Covers the compile-time-C
    specialisation (C = 256/512/1024 bf16, 128 fp32) and the generic kernel (C = 192, 64)."""
    from ppyolo_b200._lib import PPY_F32, PPY_BF16, lib, check
    o = ops()
    dt = PPY_BF16 if code == 'bf16' else PPY_F32
    g = torch.Generator().manual_seed(c + h * 31 + w)
    x = bf16_round(torch.randn((n, c, h, w), generator=g))
    ow = bf16_round(torch.randn((27, c, 3, 3), generator=g) * 2.0 / (c * 9) ** 0.5)
    ob = torch.randn(27, generator=g)
    wt = torch.randn((8, c, 3, 3), generator=g) / (c * 9) ** 0.5
    want = ref.dcnv2(x, ow, ob, wt, stride, 1)
    ho, wo = want.shape[2:]
    om = torch.nn.functional.conv2d(x, ow, ob, stride=stride, padding=1).permute(0, 2, 3, 1).contiguous()      # NHWC [.., 27]
    om_pad = torch.zeros((n, ho, wo, 32)); om_pad[..., :27] = om
    xh = o.to_nhwc(x.to(DEV), dt)
    out = torch.empty((n * ho * wo, 9 * c), dtype=torch.bfloat16 if code == 'bf16' else torch.float32, device=DEV)
    omd = om_pad.to(DEV)
    check(lib.ppy_dcn_gather(o.ptr(xh), xh.shape[-1], n, h, w, c, o.ptr(omd), 32, 3, stride, 1, o.ptr(out), dt, o.stream_ptr()), 'gather')
    torch.cuda.synchronize()
    wk = wt.permute(0, 2, 3, 1).reshape(8, 9 * c)                       # (tap, c) K order of the gathered matrix
    got = (out.float().cpu() @ wk.t()).reshape(n, ho, wo, 8).permute(0, 3, 1, 2)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=(4e-3 if code == 'bf16' else 1e-5) * scale_of(want.numpy()))


@pytest.mark.parametrize('n,cin,cout,k,hw', [(2, 64, 256, 1, 19), (1, 128, 512, 1, 30), (3, 64, 96, 1, 11), (2, 64, 64, 3, 32),
                                             (1, 256, 128, 1, 24), (2, 64, 128, 3, 46), (3, 64, 128, 3, 19), (2, 128, 64, 3, 13)])
def test_umma_conv_tma_epilogue(n, cin, cout, k, hw):
    """bf16-output layers with K <= 512 take the TMA epilogue (residual boxes in by TMA, output boxes out by TMA store):
    residual + ReLU against the oracle on bf16-rounded operands; covers M tails, N tails (cout 96) and 16x8 patch tiles."""
    g = torch.Generator().manual_seed(cin + cout + hw)
    x = bf16_round(torch.randn((n, cin, hw, hw), generator=g))
    w = bf16_round(torch.randn((cout, cin, k, k), generator=g) * (1.0 / (cin * k * k) ** 0.5))
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    res = bf16_round(torch.randn((n, cout, hw, hw), generator=g))
    base = torch.nn.functional.conv2d(x, w, None, 1, (k - 1) // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    for residual, act, want in ((res, 1, torch.relu(base + res)), (None, 2, torch.nn.functional.leaky_relu(base, 0.1))):
        got = _umma_conv(x, w, scale, shift, 1, act, residual=residual, out_f32=False)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=8e-3, atol=8e-3 * scale_of(want.numpy()))


@pytest.mark.parametrize('n,cin,cout,hw', [(2, 64, 128, 19), (1, 256, 256, 38), (3, 128, 64, 11)])
def test_umma_conv_coord_fold(n, cin, cout, hw):
    """1x1 conv after CoordConv (model/custom_layers.py:256-272): the two coordinate channels enter as the rank-2 epilogue
    term wx*xc + wy*yc (ppy_conv_params.coord_w, TMA epilogue) instead of as input channels."""
    from ppyolo_b200._lib import PPY_BF16
    o = ops()
    g = torch.Generator().manual_seed(5 * cin + cout + hw)
    x = bf16_round(torch.randn((n, cin, hw, hw), generator=g))
    w = torch.randn((cout, cin + 2, 1, 1), generator=g) * (1.0 / cin ** 0.5)
    w[:, :cin] = bf16_round(w[:, :cin])
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    want = torch.nn.functional.conv2d(ref.coord_concat(x), w) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    want = torch.nn.functional.leaky_relu(want, 0.1)
    packed = o.pack_weight(w.to(DEV), PPY_BF16, c_begin=0, c_count=cin)
    xh = o.to_nhwc(x.to(DEV), PPY_BF16, packed[1])
    coord_w = w[:, cin:, 0, 0].t().contiguous().to(DEV)
    y = o.conv_nhwc(xh, packed, cin, cout, 1, 1, 0, scale.to(DEV), shift.to(DEV), 2, PPY_BF16, coord_w=coord_w)
    got = o.from_nhwc(y, cout).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=8e-3, atol=8e-3 * scale_of(want.numpy()))


@pytest.mark.parametrize('n,h,w,act', [(2, 64, 64, 1), (1, 37, 37, 2), (3, 52, 52, 1), (2, 36, 36, 2), (2, 38, 136, 1), (1, 70, 260, 2),
                                       (33, 32, 32, 1)])
def test_stem_conv_bf16_tensor_core(n, h, w, act):
    """ppy_stem_conv3x3s2, bf16 output: NCHW fp32 image -> conv 3x3/s2/p1 (3 -> 32, model/resnet_vd.py:100) + folded BN + act
    on tcgen05 tensor cores (w % 4 == 0; other widths take the fp32 SIMT kernel), against conv2d on the same bf16-rounded image and
    weights (odd sizes cover the borders, odd output heights the half-dead tiles, wide maps several tiles per row)."""
    import ctypes
    from ppyolo_b200._lib import lib, check, PPY_BF16
    o = ops()
    g = torch.Generator().manual_seed(100 + h + w)
    x = torch.randn((n, 3, h, w), generator=g)
    wt = torch.randn((32, 3, 3, 3), generator=g) * 0.2
    scale, shift = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    want = torch.nn.functional.conv2d(bf16_round(x), bf16_round(wt), None, 2, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    want = torch.relu(want) if act == 1 else torch.nn.functional.leaky_relu(want, 0.1)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    xd = x.to(DEV).contiguous()
    y = torch.zeros((n, ho, wo, 32), dtype=torch.bfloat16, device=DEV)
    fp = ctypes.POINTER(ctypes.c_float)
    wn, sn, hn = (np.ascontiguousarray(t.numpy()) for t in (wt, scale, shift))
    check(lib.ppy_stem_conv3x3s2(o.ptr(xd), n, h, w, wn.ctypes.data_as(fp), sn.ctypes.data_as(fp), hn.ctypes.data_as(fp), 32, act,
                                 ctypes.c_void_p(y.data_ptr()), 32, PPY_BF16, o.stream_ptr()), 'stem')
    got = o.from_nhwc(y, 32).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=8e-3, atol=8e-3 * scale_of(want.numpy()))


@pytest.mark.parametrize('h,w', [(13, 17), (12, 18), (7, 7), (20, 2)])
def test_pools_bf16(h, w):
    """bf16 fast path of MaxPool2d(3,2,1) (model/resnet_vd.py:103) and AvgPool2d(2,2) (:30): two output pixels per thread,
    odd widths / heights cover the window clamping; max is exact, the average is rounded once to bf16."""
    from ppyolo_b200._lib import PPY_BF16, lib, check
    o = ops()
    g = torch.Generator().manual_seed(h * 31 + w)
    t = bf16_round(torch.randn((3, 16, h, w), generator=g))
    xb = o.to_nhwc(t.to(DEV), PPY_BF16)
    ho, wo = (h + 1) // 2, (w + 1) // 2
    yb = torch.empty((3, ho, wo, 16), dtype=torch.bfloat16, device=DEV)
    check(lib.ppy_maxpool3x3s2(o.ptr(xb), 16, o.ptr(yb), 16, 3, h, w, 16, PPY_BF16, o.stream_ptr()), 'maxpool')
    np.testing.assert_array_equal(o.from_nhwc(yb, 16).cpu().numpy(), torch.nn.functional.max_pool2d(t, 3, 2, 1).numpy())
    if h >= 2 and w >= 2:
        ya = torch.empty((3, h // 2, w // 2, 16), dtype=torch.bfloat16, device=DEV)
        check(lib.ppy_avgpool2x2(o.ptr(xb), 16, o.ptr(ya), 16, 3, h, w, 16, PPY_BF16, o.stream_ptr()), 'avgpool')
        want = torch.nn.functional.avg_pool2d(t, 2, 2).to(torch.bfloat16).float()
        np.testing.assert_array_equal(o.from_nhwc(ya, 16).cpu().numpy(), want.numpy())


@pytest.mark.parametrize('n,cin,cout,k,stride,hw,splits', [(1, 512, 27, 3, 1, 19, 0), (2, 256, 64, 1, 1, 12, 4), (1, 128, 40, 3, 2, 21, 3),
                                                           (2, 64, 128, 3, 1, 9, 2)])
def test_umma_conv_split_k(n, cin, cout, k, stride, hw, splits):
    """Partial-sum launches (ppy_conv_params.accumulate / split_k): K splits of one output tile are added atomically into a
    zeroed fp32 output, the shift enters once -- the DCN offset conv shape (few tiles, K = 4608) is the use case."""
    from ppyolo_b200._lib import PPY_F32, PPY_BF16
    o = ops()
    g = torch.Generator().manual_seed(cin + 7 * cout + splits)
    x = bf16_round(torch.randn((n, cin, hw, hw), generator=g))
    w = bf16_round(torch.randn((cout, cin, k, k), generator=g) * (1.0 / (cin * k * k) ** 0.5))
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    want = torch.nn.functional.conv2d(x, w, None, stride, (k - 1) // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    packed = o.pack_weight(w.to(DEV), PPY_BF16)
    xh = o.to_nhwc(x.to(DEV), PPY_BF16, packed[1])
    ho = (hw + 2 * ((k - 1) // 2) - k) // stride + 1
    out = torch.zeros((n, ho, ho, (cout + 7) // 8 * 8), dtype=torch.float32, device=DEV)
    o.conv_nhwc(xh, packed, cin, cout, k, stride, (k - 1) // 2, scale.to(DEV), shift.to(DEV), 0, PPY_BF16, out=out, out_code=PPY_F32,
                accumulate=True, split_k=splits)
    got = o.from_nhwc(out, cout).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-4 * scale_of(want.numpy()))


def _random_conv_cases(count, seed):
    rng = np.random.RandomState(seed)
    cases = []
    while len(cases) < count:
        k = int(rng.choice([1, 3]))
        stride = int(rng.choice([1, 1, 2])) if k == 3 else 1
        cin = int(rng.choice([8, 24, 32, 64, 128, 192, 256, 320]))
        cout = int(rng.choice([16, 27, 32, 64, 72, 96, 128, 136, 256, 258, 320, 512]))
        hw = int(rng.randint(3, 41))
        n = int(rng.randint(1, 5))
        if n * hw * hw * max(cin, cout) > 3_000_000:
            continue
        cases.append((n, cin, cout, k, stride, hw, bool(rng.randint(2)), bool(rng.randint(2))))
    return cases


@pytest.mark.parametrize('case', _random_conv_cases(48, seed=20261017), ids=lambda c: 'n%d_c%d_o%d_k%d_s%d_hw%d_r%d_b%d' % c)
def test_umma_conv_random_sweep(case):
    """Seeded random sweep over the conv kernel's dispatch space (tma_a / slab / patch / im2col / gather modes, CTA pairs and
    single CTAs, TMA and slab epilogues, M / N tails, odd map sizes): bf16 conv + scale/shift (+ residual + ReLU | leaky) against
    torch on the same bf16-rounded operands."""
    n, cin, cout, k, stride, hw, with_res, out_bf16 = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = bf16_round(torch.randn((n, cin, hw, hw), generator=g))
    w = bf16_round(torch.randn((cout, cin, k, k), generator=g) * (1.0 / (cin * k * k) ** 0.5))
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    base = torch.nn.functional.conv2d(x, w, None, stride, (k - 1) // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = bf16_round(torch.randn(base.shape, generator=g)) if (with_res and out_bf16) else None
    want = torch.relu(base + res) if res is not None else torch.nn.functional.leaky_relu(base, 0.1)
    got = _umma_conv(x, w, scale, shift, stride, 1 if res is not None else 2, residual=res, out_f32=not out_bf16)
    tol = 8e-3 if out_bf16 else 2e-4
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=tol if out_bf16 else 0, atol=tol * scale_of(want.numpy()))

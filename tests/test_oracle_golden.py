"""Pins the CPU oracle (oracle/ppyolo_ref.py) to outputs of the unmodified reference (tests/golden/*.npz).

CPU-only.  Tolerances: Matrix-NMS is bit-exact on labels and 1e-6 on scores (pure fp32 arithmetic in the
same order); layer/decode/network outputs allow the fp32 re-association noise between torch builds/threads.
"""
import numpy as np
import pytest
import torch

from oracle import ppyolo_ref as ref
from ppyolo_b200 import synth
from tests.helpers import build_model, weight_checksum, assert_preds_close

NMS_CFG = dict(score_threshold=0.01, post_threshold=0.01, nms_top_k=500, keep_top_k=100)


@pytest.mark.parametrize('name', ['zeros', 'dups', 'chain'])
@pytest.mark.parametrize('gauss', [False, True])
def test_nms_known_answers(golden, name, gauss):
    z = golden('nms')
    tag = '%s_%s' % (name, 'g' if gauss else 'l')
    out = ref.matrix_nms(z[tag + '_boxes'], z[tag + '_scores'], use_gaussian=gauss, **NMS_CFG)
    assert_preds_close(out, z[tag + '_out'], rtol=1e-6, atol=1e-7)


def test_nms_known_answer_values():
    """Hand-derived values of SURVEY.md 8c (IoUs 1/3, 1/9, 7/13)."""
    b = np.array([[0, 0, 10, 10], [5, 0, 15, 10], [8, 0, 18, 10]], np.float32)
    s = np.zeros((3, 80), np.float32)
    s[0, 1], s[1, 1], s[2, 1] = 0.9, 0.8, 0.7
    lin = ref.matrix_nms(b, s, **NMS_CFG)
    np.testing.assert_allclose(lin[:, 1], [0.9, 0.8 * 2 / 3, 0.7 * (6 / 13) / (2 / 3)], rtol=1e-6)
    gau = ref.matrix_nms(b, s, use_gaussian=True, gaussian_sigma=2.0, **NMS_CFG)
    np.testing.assert_allclose(gau[:, 1], [0.9, 0.6406, 0.4895], atol=1e-4)
    assert (ref.matrix_nms(b, np.zeros((3, 80), np.float32), **NMS_CFG) == -1).all()


@pytest.mark.parametrize('nb,nc,seed,topk,keep', [(400, 80, 0, 500, 100), (400, 80, 1, 100, 20), (3000, 80, 2, 500, 100),
                                                  (64, 3, 3, -1, 100), (1500, 20, 4, 300, 50)])
@pytest.mark.parametrize('gauss', [False, True])
def test_nms_random(golden, nb, nc, seed, topk, keep, gauss):
    z = golden('nms')
    b, s = synth.nms_inputs(nb, nc, seed=seed)
    out = ref.matrix_nms(b.numpy(), s.numpy(), 0.01, 0.01, topk, keep, use_gaussian=gauss, gaussian_sigma=2.0)
    want = z['rand_b%d_c%d_s%d_t%d_k%d_%s_out' % (nb, nc, seed, topk, keep, 'g' if gauss else 'l')]
    assert_preds_close(out, want, rtol=2e-6, atol=1e-7)


def test_nms_post_threshold_and_degenerate(golden):
    z = golden('nms')
    b, s = synth.nms_inputs(800, 80, seed=5)
    out = ref.matrix_nms(b.numpy(), s.numpy(), 0.05, 0.2, 200, 30)
    assert_preds_close(out, z['post05_out'], rtol=2e-6, atol=1e-7)
    out = ref.matrix_nms(z['degenerate_boxes'], z['degenerate_scores'], **NMS_CFG)
    assert_preds_close(out, z['degenerate_out'], rtol=1e-6, atol=1e-7)
    ba, _ = synth.nms_inputs(37, 1, seed=7)
    bb, _ = synth.nms_inputs(53, 1, seed=8)
    np.testing.assert_allclose(ref.pairwise_iou(ba.numpy(), bb.numpy()), z['jaccard_out'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('tag,stride,mask,iou_aware', [('s32', 32, [6, 7, 8], True), ('s8', 8, [0, 1, 2], True),
                                                       ('plain', 16, [3, 4, 5], False)])
def test_decode(golden, tag, stride, mask, iou_aware):
    z = golden('decode')
    anchors = np.array(build_model('r50vd')[1].head['anchors'], np.float32)[mask]
    x = torch.from_numpy(z[tag + '_in'])
    im_size = torch.from_numpy(z[tag + '_im_size'])
    if iou_aware:
        x = ref.iou_aware_score(x, 3, 80, 0.4)
        np.testing.assert_allclose(x.numpy(), z[tag + '_iouaware'], rtol=1e-5, atol=1e-5)
    for clip in (True, False):
        boxes, scores = ref.yolo_box(x, anchors, stride, 80, 1.05, im_size, clip)
        want = z['%s_boxes_clip%d' % (tag, int(clip))]
        assert np.array_equal(np.isnan(boxes.numpy()), np.isnan(want))
        np.testing.assert_allclose(boxes.numpy(), want, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(scores.numpy(), z[tag + '_scores'], rtol=1e-5, atol=1e-7)


def test_layers(golden):
    z = golden('layers')
    np.testing.assert_array_equal(ref.coord_concat(torch.zeros(1, 2, 3, 4)).numpy(), z['coord_out'])
    np.testing.assert_array_equal(ref.spp(torch.from_numpy(z['spp_in'])).numpy(), z['spp_out'])
    from model.custom_layers import Conv2dUnit
    for tag, cin, cout, k, stride, act, bias in (('c3s1', 8, 16, 3, 1, 'leaky', False), ('c3s2', 8, 16, 3, 2, 'relu', False),
                                                 ('c1s1', 16, 24, 1, 1, None, True), ('c1s2', 8, 8, 1, 2, 'relu', False)):
        u = Conv2dUnit(cin, cout, k, stride=stride, bias_attr=bias, bn=0 if bias else 1, act=act)
        synth.randomize_(u, seed=30)
        bn = None if bias else (u.bn.weight, u.bn.bias, u.bn.running_mean, u.bn.running_var)
        with torch.no_grad():
            y = ref.conv_norm_act(torch.from_numpy(z[tag + '_in']), u.conv.weight, u.conv.bias, bn, stride, act)
        np.testing.assert_allclose(y.numpy(), z[tag + '_out'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('tag,stride', [('dcn_s1', 1), ('dcn_s2', 2)])
def test_dcn(golden, tag, stride):
    z = golden('layers')
    from model.custom_layers import Conv2dUnit
    u = Conv2dUnit(16, 24, 3, stride=stride, bn=1, act='relu', use_dcn=True)
    synth.randomize_(u, seed=31, offset_scale=0.05)
    x = torch.from_numpy(z[tag + '_in'])
    with torch.no_grad():
        raw = ref.dcnv2(x, u.conv.conv_offset.weight, u.conv.conv_offset.bias, u.conv.dcn_weight, stride, 1)
    assert np.abs(z[tag + '_offsetmask'][:, :18]).max() > 1.5      # offsets really deform the sampling grid
    np.testing.assert_allclose(raw.numpy(), z[tag + '_raw'], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize('tag,stride', [('s1', 1), ('s2', 2)])
def test_dcn_far_offsets(golden, tag, stride):
    """Offsets of sigma ~6 px (max ~24 px) on a 20x20 map: most samples fall far outside the image, beyond the reference's
    one-pixel zero border -- the oracle's "outside -> 0" rule must equal the reference's clamp-into-the-border trick there too
    (golden from the unmodified reference, tests/golden/make_golden_dcn_far.py)."""
    z = golden('dcn_far')
    from model.custom_layers import Conv2dUnit
    u = Conv2dUnit(64, 24, 3, stride=stride, bn=1, act='relu', use_dcn=True)
    synth.randomize_(u, seed=33, offset_scale=0.25)
    x = torch.from_numpy(z[tag + '_in'])
    assert np.abs(z[tag + '_offsetmask'][:, :18]).max() > 18.0
    with torch.no_grad():
        raw = ref.dcnv2(x, u.conv.conv_offset.weight, u.conv.conv_offset.bias, u.conv.dcn_weight, stride, 1)
    np.testing.assert_allclose(raw.numpy(), z[tag + '_raw'], rtol=1e-4, atol=2e-5 * np.abs(z[tag + '_raw']).max())


def test_dcn_zero_offset_identity():
    """external/DCNv2/test.py:32-67 recipe: zero offset conv => mask 0.5 => 2*DCN(x) == conv(x)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 6, 8, 8), generator=g)
    w = torch.randn((5, 6, 3, 3), generator=g)
    y = ref.dcnv2(x, torch.zeros(27, 6, 3, 3), torch.zeros(27), w, 1, 1)
    np.testing.assert_allclose(2 * y.numpy(), torch.nn.functional.conv2d(x, w, padding=1).numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('arch,name', [('r18vd', 'net_r18vd_128'), ('r50vd', 'net_r50vd_128')])
def test_network(golden, arch, name):
    z = golden(name)
    model, cfg = build_model(arch)
    np.testing.assert_allclose(weight_checksum(model), z['w_checksum'], rtol=1e-12)
    x = synth.images(2, 128, seed=1)
    np.testing.assert_allclose([float(x.double().sum()), float(x.double().abs().sum())], z['x_checksum'], rtol=1e-12)
    net = ref.Net(model.state_dict(), cfg)
    res = net.forward(x, torch.from_numpy(z['im_size']), return_all=True)
    np.testing.assert_allclose(res['feats'][-1].numpy(), z['feat_last'], rtol=1e-4, atol=1e-4)
    for i, o in enumerate(res['outs']):
        np.testing.assert_allclose(o.numpy(), z['out%d' % i], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(res['boxes'].numpy(), z['boxes'], rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(res['scores'][0].numpy(), z['scores_img0_f32'], rtol=1e-3, atol=1e-6)
    for i, p in enumerate(res['preds']):
        assert_preds_close(p, z['pred%d' % i], rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize('tag', ['s1', 's2'])
def test_dcn_backward_oracle_vs_reference(golden, tag):
    """torch autograd through the oracle's dcnv2 restatement against the reference DCNv2 module's own backward
    (tests/golden/dcn_bwd.npz from make_golden_unfrozen.py; offsets up to ~8 px on a 9x9 / 12x12 map, so clamped and
    out-of-image samples are in): y, dx and the three parameter gradients -- the checker the GPU backward tests lean on.
    Tolerance 2e-5 of each tensor's scale (fp32 summation order)."""
    z = golden('dcn_bwd')
    stride = int(z[tag + '_stride'][0])
    t = lambda k: torch.from_numpy(z['%s_%s' % (tag, k)]).clone()
    x, ow, ob, w = (t(k).requires_grad_(True) for k in ('x', 'offset_w', 'offset_b', 'dcn_w'))
    y = ref.dcnv2(x, ow, ob, w, stride, 1)
    y.backward(t('dy'))
    for got, name in ((y, 'y'), (x.grad, 'dx'), (ow.grad, 'd_offset_w'), (ob.grad, 'd_offset_b'), (w.grad, 'd_dcn_w')):
        want = z['%s_%s' % (tag, name)]
        np.testing.assert_allclose(got.detach().numpy(), want, rtol=0, atol=2e-5 * np.abs(want).max(), err_msg=name)

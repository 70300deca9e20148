"""Training losses and target construction against the unmodified reference (tests/golden/losses.npz): loss values and
the gradients w.r.t. the head outputs (model/losses.py:121-356, model/iou_losses.py, tools/transform.py:1318-1421).
CPU-only: the loss modules are plain differentiable tensor code."""
import numpy as np
import pytest
import torch

import config as cfgs
from ppyolo_b200 import targets as tg
from tests.helpers import CONFIGS


def build_loss(cfg):
    iou_loss = cfgs.select_loss(cfg.iou_loss_type)(**cfg.iou_loss)
    iou_aware = cfgs.select_loss(cfg.iou_aware_loss_type)(**cfg.iou_aware_loss) if cfg.head['iou_aware'] else None
    return cfgs.select_loss(cfg.yolo_loss_type)(iou_loss=iou_loss, iou_aware_loss=iou_aware, **cfg.yolo_loss)


@pytest.mark.parametrize('tag', ['r50vd', 'r18vd'])
def test_targets_and_losses(golden, tag):
    z = golden('losses')
    cfg = CONFIGS[tag]()
    size, batch = 128, 2
    gt_bbox, gt_class, gt_score = tg.synthetic_ground_truth(batch, seed=3)
    targets = tg.gt2yolo_target(gt_bbox, gt_class, gt_score, h=size, w=size, **cfg.gt2YoloTarget)
    for i, t in enumerate(targets):
        np.testing.assert_allclose(t, z['%s_target%d' % (tag, i)], rtol=1e-6, atol=1e-7)
    assert sum(float(t[:, :, 5].sum()) for t in targets) > 0
    g = torch.Generator().manual_seed(41)
    per = 86 if cfg.head['iou_aware'] else 85
    outs = [(torch.randn((batch, 3 * per, size // s, size // s), generator=g) * 1.2).requires_grad_(True)
            for s in cfg.head['downsample']]
    anchors, masks = cfg.head['anchors'], cfg.head['anchor_masks']
    mask_anchors = [[v for aid in m for v in anchors[aid]] for m in masks]
    losses = build_loss(cfg)(outs, torch.from_numpy(gt_bbox), torch.from_numpy(gt_class), torch.from_numpy(gt_score),
                             [torch.from_numpy(t) for t in targets], anchors, masks, mask_anchors, 80)
    want_keys = {'loss_xy', 'loss_wh', 'loss_obj', 'loss_cls', 'loss_iou'} | ({'loss_iou_aware'} if per == 86 else set())
    assert set(losses) == want_keys
    for k, v in losses.items():
        np.testing.assert_allclose(float(v), float(z['%s_%s' % (tag, k)]), rtol=2e-5)
    sum(losses.values()).backward()
    for i, o in enumerate(outs):
        gs = z['%s_gradsum%d' % (tag, i)]
        np.testing.assert_allclose([float(o.grad.double().sum()), float(o.grad.double().abs().sum())], gs, rtol=1e-4)
        full = z['%s_grad%d' % (tag, i)]
        if full.ndim == 4:
            np.testing.assert_allclose(o.grad.numpy(), full, rtol=1e-4, atol=1e-6)

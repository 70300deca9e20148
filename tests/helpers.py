"""Shared builders for the tests: seeded models (this repo's module surface) and config lookup."""
import numpy as np
import torch

import config as cfgs
from model.ppyolo import PPYOLO
from ppyolo_b200 import synth

CONFIGS = {'r18vd': cfgs.PPYOLO_r18vd_Config, 'r50vd': cfgs.PPYOLO_2x_Config}


def build_model(arch, seed=0):
    cfg = CONFIGS[arch]()
    backbone = cfgs.select_backbone(cfg.backbone_type)(**cfg.backbone)
    head = cfgs.select_head(cfg.head_type)(yolo_loss=None, nms_cfg=cfg.nms_cfg, **cfg.head)
    model = PPYOLO(backbone, head)
    synth.randomize_(model, seed=seed)
    model.eval()
    head.set_dropblock(is_test=True)
    return model, cfg


def weight_checksum(model):
    w = torch.cat([v.double().flatten() for k, v in sorted(model.state_dict().items())])
    return np.array([float(w.sum()), float(w.abs().sum())])


def assert_preds_close(got, want, rtol=1e-4, atol=1e-4):
    """[M,6] rows: label exact, score and box within tolerance."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    np.testing.assert_array_equal(got[:, 0], want[:, 0])
    np.testing.assert_allclose(got[:, 1:], want[:, 1:], rtol=rtol, atol=atol)


def assert_preds_match(got, want, rtol=2e-3, atol=5e-2, max_row_mismatch=0.03):
    """End-to-end comparison of NMS outputs that tolerates ranking flips.

    The kept scores, sorted, must agree within tolerance; individual rows (label, box) may differ for at most
    ``max_row_mismatch`` of the rows: two candidates whose decayed scores differ by less than the upstream
    fp32 noise (~1e-6 relative) can swap places or swap in/out at the top-k cut.
    """
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    np.testing.assert_allclose(np.sort(got[:, 1]), np.sort(want[:, 1]), rtol=rtol, atol=1e-6)
    ok = (got[:, 0] == want[:, 0]) & np.all(np.abs(got[:, 1:] - want[:, 1:]) <= atol + rtol * np.abs(want[:, 1:]), axis=1)
    bad = int((~ok).sum())
    assert bad <= max(1, int(max_row_mismatch * len(got))), '%d of %d rows differ' % (bad, len(got))

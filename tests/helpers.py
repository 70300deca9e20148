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


def detection_drift(got, want, iou_thr=0.9):
    """How far a list of per-image [M,6] detections (label, score, x0, y0, x1, y1) is from the reference's.

    Every reference row is paired with the unused row of ``got`` that has the same label and the highest IoU; a pair
    with IoU >= ``iou_thr`` is a match.  Returns the fraction of reference rows matched, the fraction of rows whose
    label agrees at the same rank, and the largest score / box-coordinate difference over the matched pairs."""
    total = matched = same_rank = 0
    max_box = max_score = 0.0
    for g, w in zip(got, want):
        g, w = np.asarray(g, dtype=np.float64), np.asarray(w, dtype=np.float64)
        if w.shape[0] == 1 and w[0, 0] < 0:
            w = w[:0]
        if g.shape[0] == 1 and g[0, 0] < 0:
            g = g[:0]
        total += len(w)
        k = min(len(g), len(w))
        same_rank += int((g[:k, 0] == w[:k, 0]).sum())
        used = np.zeros(len(g), dtype=bool)
        for row in w:
            cand = np.where((g[:, 0] == row[0]) & ~used)[0]
            if not len(cand):
                continue
            b = g[cand, 2:]
            ix = np.clip(np.minimum(b[:, 2], row[4]) - np.maximum(b[:, 0], row[2]), 0, None)
            iy = np.clip(np.minimum(b[:, 3], row[5]) - np.maximum(b[:, 1], row[3]), 0, None)
            inter = ix * iy
            union = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) + (row[4] - row[2]) * (row[5] - row[3]) - inter
            iou = inter / np.maximum(union, 1e-12)
            j = int(np.argmax(iou))
            if iou[j] >= iou_thr:
                used[cand[j]] = True
                matched += 1
                max_box = max(max_box, float(np.abs(b[j] - row[2:]).max()))
                max_score = max(max_score, abs(float(g[cand[j], 1] - row[1])))
    return {'reference_rows': total, 'matched_frac': matched / max(total, 1), 'same_label_at_rank_frac': same_rank / max(total, 1),
            'max_box_err_px': max_box, 'max_score_err': max_score}

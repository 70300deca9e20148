"""Training-target construction for the YOLOv3 loss: the per-image ``Gt2YoloTargetSingle`` of the reference
(tools/transform.py:1318-1421), vectorised over the batch, plus the synthetic ground truth of BASELINE config C4
(SURVEY.md 8d).  Host-side numpy: in the reference this runs in the data-loader threads."""
import numpy as np


def _wh_iou(w0, h0, w1, h1):
    inter = np.minimum(w0, w1) * np.minimum(h0, h1)
    return inter / (w0 * h0 + w1 * h1 - inter)


def gt2yolo_target(gt_bbox, gt_class, gt_score, anchors, anchor_masks, downsample_ratios, num_classes, h, w,
                   iou_thresh=1.):
    """gt_bbox [N,G,4] normalised (cx,cy,w,h); gt_class [N,G] int; gt_score [N,G] -> list of float32
    [N, A, 6+C, H/s, W/s] = (tx, ty, tw, th, 2 - gw*gh, score, one-hot class)."""
    anchors_py = [[v for v in a] for a in anchors]          # python numbers, like the reference's config lists
    anchors = np.asarray(anchors, dtype=np.float64)
    an_w, an_h = anchors[:, 0] / w, anchors[:, 1] / h
    gt_bbox = np.asarray(gt_bbox, dtype=np.float32)
    gt_score = np.asarray(gt_score, dtype=np.float32)
    n = gt_bbox.shape[0]
    targets = []
    for mask, ratio in zip(anchor_masks, downsample_ratios):
        gh_, gw_ = int(h / ratio), int(w / ratio)
        tgt = np.zeros((n, len(mask), 6 + num_classes, gh_, gw_), dtype=np.float32)
        for b in range(n):
            for g in range(gt_bbox.shape[1]):
                gx, gy, bw, bh = gt_bbox[b, g]          # np.float32 scalars: the arithmetic below is float32 like the reference's
                score = gt_score[b, g]
                if bw <= 0. or bh <= 0. or score <= 0.:
                    continue
                ious = _wh_iou(bw, bh, an_w, an_h)
                best = int(np.argmax(ious)) if ious.max() > 0 else -1      # first maximum, like the reference's scan
                gi, gj = int(gx * gw_), int(gy * gh_)
                cls = int(gt_class[b, g])
                for slot, an_idx in enumerate(mask):
                    hit = an_idx == best or (iou_thresh < 1 and ious[an_idx] > iou_thresh)
                    if not hit:
                        continue
                    t = tgt[b, slot]
                    t[0, gj, gi] = gx * gw_ - gi
                    t[1, gj, gi] = gy * gh_ - gj
                    t[2, gj, gi] = np.log(bw * w / anchors_py[an_idx][0])
                    t[3, gj, gi] = np.log(bh * h / anchors_py[an_idx][1])
                    t[4, gj, gi] = 2.0 - bw * bh
                    t[5, gj, gi] = score
                    t[6 + cls, gj, gi] = 1.
        targets.append(tgt)
    return targets


def synthetic_ground_truth(batch, num_classes=80, boxes_per_image=5, max_boxes=50, seed=0):
    """C4 inputs: 5 random GT boxes per image (cx,cy ~ U(.2,.8), w,h ~ U(.05,.35), class ~ U{0..C-1}, score 1), padded
    with zeros to 50 like the reference's PadBox."""
    rng = np.random.RandomState(seed)
    gt_bbox = np.zeros((batch, max_boxes, 4), np.float32)
    gt_class = np.zeros((batch, max_boxes), np.int32)
    gt_score = np.zeros((batch, max_boxes), np.float32)
    gt_bbox[:, :boxes_per_image, 0:2] = rng.uniform(0.2, 0.8, (batch, boxes_per_image, 2))
    gt_bbox[:, :boxes_per_image, 2:4] = rng.uniform(0.05, 0.35, (batch, boxes_per_image, 2))
    gt_class[:, :boxes_per_image] = rng.randint(0, num_classes, (batch, boxes_per_image))
    gt_score[:, :boxes_per_image] = 1.0
    return gt_bbox, gt_class, gt_score

"""Fine-grained YOLOv3 loss of PP-YOLO training (reference model/losses.py: YOLOv3Loss :84-356, its decode copy
``paddle_yolo_box`` :22-81), same constructor/call signatures and the same dict of 4-6 scalar losses.

Pure tensor code (autograd-differentiable, device agnostic).  Differences in *how*: the per-image Python loop over
``jaccard(pred, gt)`` (:309-325) is one batched IoU; the decode used for the ignore mask is specialised to what the
mask needs (boxes only, normalised image size 1)."""
import numpy as np
import torch

try:
    from collections.abc import Sequence
except ImportError:  # pragma: no cover
    from collections import Sequence


_DEVICE_CONSTANTS = {}


def device_constant(values, device):
    """Small host constant (anchor sizes) as a cached device tensor: one upload per (values, device) -- also what keeps the
    loss capturable in a CUDA graph (no host-to-device copy inside the step)."""
    key = (tuple(float(v) for v in np.asarray(values, dtype=np.float64).reshape(-1)), str(device))
    t = _DEVICE_CONSTANTS.get(key)
    if t is None:
        t = torch.tensor(key[0], dtype=torch.float32, device=device)
        _DEVICE_CONSTANTS[key] = t
    return t


def _bce(p, t):
    """t * -log(p) + (1 - t) * -log(1 - p) with the reference's 1e-9 guards."""
    return t * (0 - torch.log(p + 1e-9)) + (1 - t) * (0 - torch.log(1 - p + 1e-9))


def decode_boxes_anchor_major(output, anchors, stride, num_classes, scale_x_y):
    """Boxes of ``paddle_yolo_box`` (:22-81) with im_size = 1 and no clipping: [N, A*S*S, 4] xyxy in (a, h, w) order."""
    n, _, size, _ = output.shape
    anchors = device_constant(anchors, output.device).view(-1, 2)
    a = anchors.shape[0]
    t = output.reshape(n, a, 5 + num_classes, size, size)
    gx = torch.arange(size, dtype=torch.float32, device=output.device).view(1, 1, 1, size)
    gy = torch.arange(size, dtype=torch.float32, device=output.device).view(1, 1, size, 1)
    cx = (scale_x_y * torch.sigmoid(t[:, :, 0]) + gx - (scale_x_y - 1.0) * 0.5) * stride
    cy = (scale_x_y * torch.sigmoid(t[:, :, 1]) + gy - (scale_x_y - 1.0) * 0.5) * stride
    bw = torch.exp(t[:, :, 2]) * anchors[:, 0].view(1, a, 1, 1)
    bh = torch.exp(t[:, :, 3]) * anchors[:, 1].view(1, a, 1, 1)
    box = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], dim=-1)      # [N, A, S, S, 4]
    return box.reshape(n, a * size * size, 4) / size / stride


def batched_iou(a, b):
    """[N, P, 4] x [N, G, 4] -> [N, P, G] (jaccard of the reference's matrix_nms.py:33-47, batched)."""
    lo = torch.max(a[:, :, None, :2], b[:, None, :, :2])
    hi = torch.min(a[:, :, None, 2:], b[:, None, :, 2:])
    d = torch.clamp(hi - lo, min=0)
    inter = d[..., 0] * d[..., 1]
    area_a = ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1]))[:, :, None]
    area_b = ((b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]))[:, None, :]
    return inter / (area_a + area_b - inter)


class YOLOv3Loss(object):
    def __init__(self, ignore_thresh=0.7, label_smooth=True, use_fine_grained_loss=False, iou_loss=None,
                 iou_aware_loss=None, downsample=[32, 16, 8], scale_x_y=1., match_score=False):
        self._ignore_thresh = ignore_thresh
        self._label_smooth = label_smooth
        self._use_fine_grained_loss = use_fine_grained_loss
        self._iou_loss, self._iou_aware_loss = iou_loss, iou_aware_loss
        self.downsample, self.scale_x_y, self.match_score = downsample, scale_x_y, match_score

    def __call__(self, outputs, gt_box, gt_label, gt_score, targets, anchors, anchor_masks, mask_anchors, num_classes):
        if outputs[0].is_cuda and self.use_kernel and self._kernel_supported(gt_box, mask_anchors):
            return self._fused_loss(outputs, targets, gt_box, num_classes, mask_anchors)
        return self._get_fine_grained_loss(outputs, targets, gt_box, num_classes, mask_anchors, self._ignore_thresh)

    use_kernel = True      # CUDA tensors: fused forward/backward loss kernels (csrc/loss.cu); False = the tensor code below

    def _kernel_supported(self, gt_box, mask_anchors):
        from model.iou_losses import IouLoss, IouAwareLoss
        il, al = self._iou_loss, self._iou_aware_loss
        if il is not None and (type(il) is not IouLoss or il.ciou_term):
            return False
        if al is not None and type(al) is not IouAwareLoss:
            return False
        return gt_box.shape[1] <= 128 and all(len(a) // 2 <= 8 for a in mask_anchors)

    def _fused_loss(self, outputs, targets, gt_box, num_classes, mask_anchors):
        """Same dict as ``_get_fine_grained_loss`` from one forward launch per scale (and one backward launch per scale through
        autograd): model/losses.py:121-356 + model/iou_losses.py in csrc/loss.cu."""
        from ppyolo_b200 import ops
        assert len(outputs) == len(targets), "YOLOv3 output layer number not equal target number"
        il, al = self._iou_loss, self._iou_aware_loss
        scales = []
        for i, anchors in enumerate(mask_anchors):
            sxy = self.scale_x_y if not isinstance(self.scale_x_y, Sequence) else self.scale_x_y[i]
            scales.append(dict(anchors=list(anchors), stride=self.downsample[i], scale_x_y=float(sxy)))
        vec = ops.yolo_loss_fused(outputs, targets, gt_box, scales, num_classes, self._ignore_thresh, al is not None, il is not None,
                                  il._loss_weight if il is not None else 0.0, il.loss_square if il is not None else True,
                                  al._loss_weight if al is not None else 0.0, self.match_score)
        keys = ['loss_xy', 'loss_wh', 'loss_obj', 'loss_cls'] + (['loss_iou'] if il is not None else []) + \
               (['loss_iou_aware'] if al is not None else [])
        return {k: vec[ops.LOSS_NAMES.index(k)] for k in keys}

    def _get_fine_grained_loss(self, outputs, targets, gt_box, num_classes, mask_anchors, ignore_thresh, eps=1.e-10):
        assert len(outputs) == len(targets), "YOLOv3 output layer number not equal target number"
        batch_size = gt_box.shape[0]
        totals = {'loss_xy': 0.0, 'loss_wh': 0.0, 'loss_obj': 0.0, 'loss_cls': 0.0}
        if self._iou_loss is not None:
            totals['loss_iou'] = 0.0
        if self._iou_aware_loss is not None:
            totals['loss_iou_aware'] = 0.0
        gt_xyxy = torch.cat([gt_box[..., 0:2] - gt_box[..., 2:4] / 2., gt_box[..., 0:2] + gt_box[..., 2:4] / 2.], dim=-1)
        for i, (output, target, anchors) in enumerate(zip(outputs, targets, mask_anchors)):
            stride = self.downsample[i]
            an_num = len(anchors) // 2
            ioup = None
            if self._iou_aware_loss is not None:
                ioup = torch.sigmoid(output[:, :an_num])
                output = output[:, an_num:]
            n, _, size, _ = output.shape
            o = output.reshape(n, an_num, 5 + num_classes, size, size)
            x, y, w, h, obj = o[:, :, 0], o[:, :, 1], o[:, :, 2], o[:, :, 3], o[:, :, 4]
            cls = o[:, :, 5:].permute(0, 1, 3, 4, 2)
            tx, ty, tw, th = target[:, :, 0], target[:, :, 1], target[:, :, 2], target[:, :, 3]
            tscale, tobj = target[:, :, 4], target[:, :, 5]
            tcls = target[:, :, 6:].permute(0, 1, 3, 4, 2)
            tscale_tobj = tscale * tobj
            sxy = self.scale_x_y if not isinstance(self.scale_x_y, Sequence) else self.scale_x_y[i]

            if abs(sxy - 1.0) < eps:
                loss_x = (_bce(torch.sigmoid(x), tx) * tscale_tobj).sum((1, 2, 3))
                loss_y = (_bce(torch.sigmoid(y), ty) * tscale_tobj).sum((1, 2, 3))
            else:  # grid sensitive: L1 on the decoded offset
                dx = sxy * torch.sigmoid(x) - 0.5 * (sxy - 1.0)
                dy = sxy * torch.sigmoid(y) - 0.5 * (sxy - 1.0)
                loss_x = (torch.abs(dx - tx) * tscale_tobj).sum((1, 2, 3))
                loss_y = (torch.abs(dy - ty) * tscale_tobj).sum((1, 2, 3))
            loss_w = (torch.abs(w - tw) * tscale_tobj).sum((1, 2, 3))
            loss_h = (torch.abs(h - th) * tscale_tobj).sum((1, 2, 3))

            if self._iou_loss is not None:
                li = self._iou_loss(x, y, w, h, tx, ty, tw, th, anchors, stride, batch_size, sxy)
                totals['loss_iou'] = totals['loss_iou'] + (li * tscale_tobj).sum((1, 2, 3)).mean()
            if self._iou_aware_loss is not None:
                la = self._iou_aware_loss(ioup, x, y, w, h, tx, ty, tw, th, anchors, stride, batch_size, sxy)
                totals['loss_iou_aware'] = totals['loss_iou_aware'] + (la * tobj).sum((1, 2, 3)).mean()

            pos, neg = self._calc_obj_loss(output, obj, tobj, gt_xyxy, anchors, num_classes, stride, ignore_thresh, sxy)
            loss_cls = (_bce(torch.sigmoid(cls), tcls).sum(4) * tobj).sum((1, 2, 3))
            totals['loss_xy'] = totals['loss_xy'] + (loss_x + loss_y).mean()
            totals['loss_wh'] = totals['loss_wh'] + (loss_w + loss_h).mean()
            totals['loss_obj'] = totals['loss_obj'] + (pos + neg).mean()
            totals['loss_cls'] = totals['loss_cls'] + loss_cls.mean()
        return totals

    def _calc_obj_loss(self, output, obj, tobj, gt_xyxy, anchors, num_classes, stride, ignore_thresh, scale_x_y):
        """Objectness BCE; predictions whose best IoU with any GT exceeds ``ignore_thresh`` and that are not
        positives are ignored (reference :292-356)."""
        n, _, size, _ = output.shape
        an_num = len(anchors) // 2
        with torch.no_grad():
            boxes = decode_boxes_anchor_major(output, anchors, stride, num_classes, scale_x_y)
            max_iou = batched_iou(boxes, gt_xyxy).max(-1)[0]
            iou_mask = (max_iou <= ignore_thresh).float()
            if self.match_score:
                o = output.reshape(n, an_num, 5 + num_classes, size, size)
                prob = (torch.sigmoid(o[:, :, 4:5]) * torch.sigmoid(o[:, :, 5:])).amax(2).reshape(n, -1)
                iou_mask = iou_mask * (prob <= 0.25).float()
            iou_mask = iou_mask.reshape(n, an_num, size, size)
            noobj_mask = (1.0 - (tobj > 0.).float()) * iou_mask
        p = torch.sigmoid(obj)
        pos = (tobj * (0 - torch.log(p + 1e-9))).sum((1, 2, 3))
        neg = (noobj_mask * (0 - torch.log(1 - p + 1e-9))).sum((1, 2, 3))
        return pos, neg

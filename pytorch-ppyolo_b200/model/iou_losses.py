"""IoU loss and IoU-aware loss of PP-YOLO training (reference model/iou_losses.py: IouLoss :15-191,
IouAwareLoss :194-246), same constructor and call signatures.  Pure tensor code, differentiable through torch
autograd and device agnostic (the reference hard-codes ``.cuda()`` at :162,:167)."""
import math

import torch


class IouLoss(object):
    """loss = (1 - iou^2) * loss_weight between the decoded prediction and the decoded target box."""

    def __init__(self, loss_weight=2.5, max_height=608, max_width=608, ciou_term=False, loss_square=True):
        self._loss_weight = loss_weight
        self._MAX_HI, self._MAX_WI = max_height, max_width
        self.ciou_term, self.loss_square = ciou_term, loss_square

    def __call__(self, x, y, w, h, tx, ty, tw, th, anchors, downsample_ratio, batch_size, scale_x_y=1., ioup=None,
                 eps=1.e-10):
        pred = self._bbox_transform(x, y, w, h, anchors, downsample_ratio, batch_size, False, scale_x_y, eps)
        gt = self._bbox_transform(tx, ty, tw, th, anchors, downsample_ratio, batch_size, True, scale_x_y, eps)
        iouk = self._iou(pred, gt, ioup, eps)
        loss = 1. - iouk * iouk if self.loss_square else 1. - iouk
        return loss * self._loss_weight

    def _iou(self, pred, gt, ioup=None, eps=1.e-10):
        x1, y1, x2, y2 = pred
        x1g, y1g, x2g, y2g = gt
        x2, y2 = torch.max(x1, x2), torch.max(y1, y2)
        iw = torch.clamp(torch.min(x2, x2g) - torch.max(x1, x1g), min=0)
        ih = torch.clamp(torch.min(y2, y2g) - torch.max(y1, y1g), min=0)
        inter = iw * ih
        union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter + eps
        iouk = inter / union
        if self.ciou_term:
            iouk = iouk - self.get_ciou_term((x1, y1, x2, y2), gt, iouk, eps)
        return iouk

    def get_ciou_term(self, pred, gt, iouk, eps):
        x1, y1, x2, y2 = pred
        x1g, y1g, x2g, y2g = gt
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        w = (x2 - x1) + ((x2 - x1) == 0).float()
        h = (y2 - y1) + ((y2 - y1) == 0).float()
        cxg, cyg, wg, hg = (x1g + x2g) / 2, (y1g + y2g) / 2, x2g - x1g, y2g - y1g
        ex1, ey1, ex2, ey2 = torch.min(x1, x1g), torch.min(y1, y1g), torch.max(x2, x2g), torch.max(y2, y2g)
        centre_d2 = (cx - cxg) * (cx - cxg) + (cy - cyg) * (cy - cyg)
        diag_d2 = (ex2 - ex1) * (ex2 - ex1) + (ey2 - ey1) * (ey2 - ey1)
        diou = (centre_d2 + eps) / (diag_d2 + eps)
        dang = torch.atan(wg / hg) - torch.atan(w / h)
        ar = 4. / math.pi / math.pi * dang * dang
        alpha = (ar / (1 - iouk + ar + eps)).detach()
        return diou + alpha * ar

    def _bbox_transform(self, dcx, dcy, dw, dh, anchors, downsample_ratio, batch_size, is_gt, scale_x_y, eps):
        """Encoded (x, y, w, h) maps [N, A, S, S] -> normalised corner coordinates (reference :137-191)."""
        n_anchor, size = dcx.shape[1], dcx.shape[2]
        dev = dcx.device
        gx = torch.arange(size, dtype=torch.float32, device=dev).view(1, 1, 1, size)
        gy = torch.arange(size, dtype=torch.float32, device=dev).view(1, 1, size, 1)
        if is_gt:
            cx, cy = (dcx + gx) / size, (dcy + gy) / size
        else:
            sx, sy = torch.sigmoid(dcx), torch.sigmoid(dcy)
            if abs(scale_x_y - 1.0) > eps:
                sx = scale_x_y * sx - 0.5 * (scale_x_y - 1)
                sy = scale_x_y * sy - 0.5 * (scale_x_y - 1)
            cx, cy = (sx + gx) / size, (sy + gy) / size
        from model.losses import device_constant
        aw = device_constant(anchors[0::2], dev).view(1, n_anchor, 1, 1)
        ah = device_constant(anchors[1::2], dev).view(1, n_anchor, 1, 1)
        pw = (torch.exp(dw) * aw) / (size * downsample_ratio)
        ph = (torch.exp(dh) * ah) / (size * downsample_ratio)
        box = (cx - 0.5 * pw, cy - 0.5 * ph, cx + 0.5 * pw, cy + 0.5 * ph)
        return tuple(b.detach() for b in box) if is_gt else box


class IouAwareLoss(IouLoss):
    """-iou * log(ioup): trains the IoU-prediction channel.  As in the reference (:241-243) the map is summed over its
    LAST axis (W) and kept as [N, A, S, 1]; the caller broadcasts it against the objectness target."""

    def __init__(self, loss_weight=1.0, max_height=608, max_width=608):
        super(IouAwareLoss, self).__init__(loss_weight=loss_weight, max_height=max_height, max_width=max_width)

    def __call__(self, ioup, x, y, w, h, tx, ty, tw, th, anchors, downsample_ratio, batch_size, scale_x_y, eps=1.e-10):
        pred = self._bbox_transform(x, y, w, h, anchors, downsample_ratio, batch_size, False, scale_x_y, eps)
        gt = self._bbox_transform(tx, ty, tw, th, anchors, downsample_ratio, batch_size, True, scale_x_y, eps)
        iouk = self._iou(pred, gt, ioup, eps)
        loss = (iouk * (0 - torch.log(ioup + 1e-9))).sum(-1).unsqueeze(-1)
        return loss * self._loss_weight

"""CPU ORACLE -- test infrastructure only, never a product path.

A plain, loop-free restatement (torch CPU fp32 + numpy) of the reference's PP-YOLO inference path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package; nothing under ``pytorch-ppyolo_b200/`` does.

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports the unmodified reference
from ``/root/reference`` (with the ``.cuda()`` identity shim) and stores its outputs on seeded inputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function here against them.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# elementwise / glue ops
# ----------------------------------------------------------------------------------------------
def activation(x, act):
    """model/custom_layers.py:128-139 -- relu, leaky(0.1), mish, or None."""
    if act is None:
        return x
    if act == 'relu':
        return torch.relu(x)
    if act == 'leaky':
        return F.leaky_relu(x, 0.1)
    if act == 'mish':
        return x * torch.tanh(F.softplus(x))
    raise NotImplementedError(act)


def conv_norm_act(x, weight, bias=None, bn=None, stride=1, act=None):
    """Conv2dUnit.forward, model/custom_layers.py:243-253: conv(pad=(k-1)//2) -> eval BN -> act.

    ``bn`` = (gamma, beta, running_mean, running_var) or None.
    """
    k = weight.shape[-1]
    y = F.conv2d(x, weight, bias, stride=stride, padding=(k - 1) // 2)
    if bn is not None:
        gamma, beta, mean, var = bn
        y = F.batch_norm(y, mean, var, gamma, beta, training=False, eps=BN_EPS)
    return activation(y, act)


def coord_concat(x):
    """CoordConv, model/custom_layers.py:256-272: cat([x, xs, ys]) with xs=i/(w-1)*2-1 along W, ys along H."""
    b, _, h, w = x.shape
    xs = torch.arange(w, dtype=torch.float32, device=x.device) / (w - 1) * 2.0 - 1
    ys = torch.arange(h, dtype=torch.float32, device=x.device) / (h - 1) * 2.0 - 1
    xs = xs.view(1, 1, 1, w).expand(b, 1, h, w)
    ys = ys.view(1, 1, h, 1).expand(b, 1, h, w)
    return torch.cat([x, xs.to(x.dtype), ys.to(x.dtype)], dim=1)


def spp(x):
    """SPP, model/custom_layers.py:275-290 (seq='asc')."""
    return torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)], dim=1)


def dcnv2(x, offset_w, offset_b, dcn_w, stride=1, padding=1):
    """DCNv2.forward, model/custom_layers.py:551-677, restated with the standard sampling rule.

    offset/mask come from a plain conv; channel 2t is dy and 2t+1 is dx of tap t (:603-605), mask is the
    sigmoid of the last kH*kW channels (:560-561).  The sample of tap (i,j) for output (ho,wo) sits at
    (ho*stride - padding + i + dy, wo*stride - padding + j + dx) in unpadded input coordinates; each of
    the 4 bilinear corners contributes 0 when it lies outside the image.  (The reference reaches the
    same values by clamping into a zero border, :571-574 and :614-615, as long as |offset| keeps the
    clamped sample inside that border -- the golden test covers offsets up to +-3 px at 9x9.)
    Output size (H + 2p - (k-1)) // stride (:567-568).  K order of the contraction is (c, kh, kw) (:661-675).
    """
    n, c, h, w = x.shape
    cout, _, kh, kw = dcn_w.shape
    om = F.conv2d(x, offset_w, offset_b, stride=stride, padding=padding)
    ho = (h + 2 * padding - (kh - 1)) // stride
    wo = (w + 2 * padding - (kw - 1)) // stride
    om = om[:, :, :ho, :wo]
    taps = kh * kw
    off = om[:, :2 * taps].reshape(n, taps, 2, ho, wo)
    mask = torch.sigmoid(om[:, 2 * taps:])                                   # [n, taps, ho, wo]
    dev = x.device
    om = om.float()
    off = om[:, :2 * taps].reshape(n, taps, 2, ho, wo)
    mask = torch.sigmoid(om[:, 2 * taps:])
    base_y = (torch.arange(ho, dtype=torch.float32, device=dev) * stride - padding).view(1, 1, ho, 1)
    base_x = (torch.arange(wo, dtype=torch.float32, device=dev) * stride - padding).view(1, 1, 1, wo)
    tap_y = torch.arange(kh, dtype=torch.float32, device=dev).repeat_interleave(kw).view(1, taps, 1, 1)
    tap_x = torch.arange(kw, dtype=torch.float32, device=dev).repeat(kh).view(1, taps, 1, 1)
    py = base_y + tap_y + off[:, :, 0]                                       # [n, taps, ho, wo]
    px = base_x + tap_x + off[:, :, 1]
    y0, x0 = torch.floor(py), torch.floor(px)
    ly, lx = py - y0, px - x0
    flat = x.reshape(n, c, h * w)
    cols = torch.zeros(n, c, taps, ho, wo, dtype=x.dtype, device=dev)
    for dy, dx, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
        yy, xx = y0 + dy, x0 + dx
        ok = (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)
        idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).long()           # [n, taps, ho, wo]
        g = torch.gather(flat, 2, idx.reshape(n, 1, -1).expand(n, c, -1)).reshape(n, c, taps, ho, wo)
        cols = cols + g * (wgt * ok.float()).unsqueeze(1).to(x.dtype)
    cols = cols * mask.unsqueeze(1).to(x.dtype)
    cols = cols.reshape(n, c * taps, ho * wo)                                # K order (c, kh, kw)
    out = torch.matmul(dcn_w.reshape(cout, c * taps), cols)
    return out.reshape(n, cout, ho, wo)


# ----------------------------------------------------------------------------------------------
# head post-processing
# ----------------------------------------------------------------------------------------------
def _logit_clamped(p, eps=1e-7):
    """_de_sigmoid, model/head.py:97-109."""
    p = torch.clamp(p, eps, 1 / eps)
    p = torch.clamp(1.0 / p - 1.0, eps, 1 / eps)
    return -torch.log(p)


def iou_aware_score(output, an_num, num_classes, factor):
    """get_iou_aware_score, model/head.py:83-141: [N, A*(6+C), H, W] -> [N, A*(5+C), H, W]."""
    ioup = torch.sigmoid(output[:, :an_num])
    rest = output[:, an_num:]
    per = rest.shape[1] // an_num
    pieces = []
    for a in range(an_num):
        blk = rest[:, per * a: per * (a + 1)]
        obj = torch.sigmoid(blk[:, 4:5])
        fused = torch.pow(obj, 1 - factor) * torch.pow(ioup[:, a:a + 1], factor)
        pieces += [blk[:, :4], _logit_clamped(fused), blk[:, 5:5 + num_classes]]
    return torch.cat(pieces, dim=1)


def yolo_box(conv_output, anchors, stride, num_classes, scale_x_y, im_size, clip_bbox=True):
    """yolo_box, model/head.py:21-80.  Returns boxes [N, H*W*A, 4] (xyxy, image pixels) and scores
    [N, H*W*A, C]; box order (h, w, anchor) (:58); assumes square maps like the reference (:25-27)."""
    n, _, size, _ = conv_output.shape
    anchors = torch.as_tensor(np.asarray(anchors, dtype=np.float32)).reshape(-1, 2).to(conv_output.device)
    a = anchors.shape[0]
    t = conv_output.permute(0, 2, 3, 1).reshape(n, size, size, a, 5 + num_classes)
    gx = torch.arange(size, dtype=torch.float32, device=conv_output.device).view(1, 1, size, 1)
    gy = torch.arange(size, dtype=torch.float32, device=conv_output.device).view(1, size, 1, 1)
    grid = torch.stack([gx.expand(1, size, size, 1), gy.expand(1, size, size, 1)], dim=-1)   # (x_idx, y_idx) :33-37
    xy = (scale_x_y * torch.sigmoid(t[..., 0:2]) + grid - (scale_x_y - 1.0) * 0.5) * stride
    wh = torch.exp(t[..., 2:4]) * anchors
    xyxy = torch.cat([xy - wh / 2, xy + wh / 2], dim=-1).reshape(n, size * size * a, 4)
    scores = (torch.sigmoid(t[..., 4:5]) * torch.sigmoid(t[..., 5:])).reshape(n, size * size * a, num_classes)
    wh_img = torch.stack([im_size[:, 1], im_size[:, 0]], dim=1).unsqueeze(1)               # (w, h) :61-65
    p0 = xyxy[:, :, 0:2] / size / stride * wh_img
    p1 = xyxy[:, :, 2:4] / size / stride * wh_img
    if clip_bbox:
        p0 = torch.where(p0 < 0, p0 * 0, p0)                                               # :73-74
        p1 = torch.where(p1 > wh_img, wh_img.expand_as(p1), p1)                            # :75-76
    return torch.cat([p0, p1], dim=-1), scores


def pairwise_iou(a, b):
    """jaccard, model/matrix_nms.py:15-47 (numpy fp32; 0/0 stays NaN like torch)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    hi = np.minimum(a[:, None, 2:], b[None, :, 2:])
    lo = np.maximum(a[:, None, :2], b[None, :, :2])
    d = np.maximum(hi - lo, np.float32(0))
    inter = d[..., 0] * d[..., 1]
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None]
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :]
    with np.errstate(divide='ignore', invalid='ignore'):
        return inter / (area_a + area_b - inter)


def matrix_nms(bboxes, scores, score_threshold, post_threshold, nms_top_k, keep_top_k, use_gaussian=False,
               gaussian_sigma=2.):
    """matrix_nms + _matrix_nms, model/matrix_nms.py:51-151, numpy fp32.

    Sorts are stable (ties keep (box, class) order); the reference's ``torch.argsort`` leaves ties
    unspecified, so golden inputs are tie-free.  Returns float32 [M,6] or [[-1]*6].
    """
    boxes = np.asarray(bboxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    empty = np.full((1, 6), -1.0, dtype=np.float32)
    box_idx, labels = np.nonzero(scores > np.float32(score_threshold))          # row-major (box, class) :115-117
    cand = scores[box_idx, labels]
    if cand.size == 0:
        return empty
    order = np.argsort(-cand, kind='stable')
    if nms_top_k > 0 and order.size > nms_top_k:
        order = order[:nms_top_k]
    b, s, l = boxes[box_idx[order]], cand[order], labels[order]
    n = s.size
    iou = np.triu(pairwise_iou(b, b), k=1)                                      # :67-68
    same = np.triu((l[None, :] == l[:, None]).astype(np.float32), k=1)          # :71-73
    decay_iou = iou * same                                                      # NaN*0 stays NaN, as in torch
    comp = decay_iou.max(axis=0)                                                # :77  column-wise max
    comp_rows = np.broadcast_to(comp[:, None], (n, n))                          # :78  compensate of the *row* box
    with np.errstate(divide='ignore', invalid='ignore'):
        if use_gaussian:
            sig = np.float32(gaussian_sigma)
            dm = np.exp(-1 * sig * decay_iou ** 2) / np.exp(-1 * sig * comp_rows ** 2)
        else:
            dm = (1 - decay_iou) / (1 - comp_rows)
    s = s * dm.min(axis=0)                                                      # :96  (np.min propagates NaN)
    keep = s >= np.float32(post_threshold)                                      # NaN compares False :132
    if keep.sum() == 0:
        return empty
    b, s, l = b[keep], s[keep], l[keep]
    order = np.argsort(-s, kind='stable')[:keep_top_k]
    return np.concatenate([l[order, None].astype(np.float32), s[order, None], b[order]], axis=1).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# whole network, driven by a reference-format state_dict
# ----------------------------------------------------------------------------------------------
class Net(object):
    """Functional PP-YOLO forward from a state_dict with the reference's key names.

    ``cfg`` is a config object with ``backbone_type``, ``backbone``, ``head`` and ``nms_cfg`` attributes
    (the reference's ``config/ppyolo_*.py`` classes or this repo's mirrors).
    """

    def __init__(self, state_dict, cfg, device='cpu', dtype=torch.float32, channels_last=False):
        """``device`` / ``dtype`` / ``channels_last``: bench.py also runs this stock-PyTorch restatement on the GPU (cuDNN /
        cuBLAS kernels only) as the "reference on the same B200" baseline; tests and the CPU arm use the defaults."""
        self.device, self.dtype, self.channels_last = torch.device(device), dtype, channels_last
        self.sd = {k: v.detach().float().to(self.device) for k, v in state_dict.items()}
        if dtype != torch.float32 or channels_last:
            for k, v in self.sd.items():
                if v.dim() == 4:
                    v = v.to(dtype)
                    self.sd[k] = v.contiguous(memory_format=torch.channels_last) if channels_last else v
                elif k.endswith(('.bias', '.weight', 'running_mean', 'running_var')) and not k.endswith('conv_offset.bias'):
                    self.sd[k] = v.to(dtype) if '.bn.' not in k else v
        self.cfg = cfg

    # -- one Conv2dUnit -----------------------------------------------------------------------
    def unit(self, prefix, x, stride=1, act=None):
        sd = self.sd
        bn = None
        if prefix + '.bn.weight' in sd:
            bn = tuple(sd[prefix + '.bn.' + k] for k in ('weight', 'bias', 'running_mean', 'running_var'))
        if prefix + '.conv.dcn_weight' in sd:
            y = dcnv2(x, sd[prefix + '.conv.conv_offset.weight'], sd[prefix + '.conv.conv_offset.bias'].to(x.dtype),
                      sd[prefix + '.conv.dcn_weight'], stride=stride, padding=1)
            if bn is not None:
                y = F.batch_norm(y, bn[2], bn[3], bn[0], bn[1], training=False, eps=BN_EPS)
            return activation(y, act)
        return conv_norm_act(x, sd[prefix + '.conv.weight'], sd.get(prefix + '.conv.bias'), bn, stride, act)

    # -- backbone (model/resnet_vd.py) ---------------------------------------------------------
    def bottleneck(self, p, x, stride, first_of_stage, is_first):
        """ConvBlock :15-57 (first_of_stage) / IdentityBlock :60-87."""
        y = self.unit(p + '.conv1', x, 1, 'relu')
        y = self.unit(p + '.conv2', y, stride, 'relu')
        y = self.unit(p + '.conv3', y, 1, None)
        if first_of_stage:
            if is_first:
                sc = self.unit(p + '.conv4', x, stride, None)
            else:
                sc = self.unit(p + '.conv4', F.avg_pool2d(x, 2, 2), 1, None)
        else:
            sc = x
        return torch.relu(y + sc)

    def basic(self, p, x, stride, is_first):
        """BasicBlock :224-267."""
        y = self.unit(p + '.conv1', x, stride, 'relu')
        y = self.unit(p + '.conv2', y, 1, None)
        if stride == 2 or is_first:
            sc = self.unit(p + '.conv3', x if is_first else F.avg_pool2d(x, 2, 2), stride if is_first else 1, None)
        else:
            sc = x
        return torch.relu(y + sc)

    def backbone(self, x):
        """Resnet50Vd.forward :132-168 / Resnet18Vd.forward :302-330."""
        r50 = self.cfg.backbone_type == 'Resnet50Vd'
        depths = (3, 4, 6, 3) if r50 else (2, 2, 2, 2)
        x = self.unit('backbone.stage1_conv1_1', x, 2, 'relu')
        x = self.unit('backbone.stage1_conv1_2', x, 1, 'relu')
        x = self.unit('backbone.stage1_conv1_3', x, 1, 'relu')
        x = F.max_pool2d(x, 3, 2, 1)
        feats = []
        for stage, depth in zip((2, 3, 4, 5), depths):
            for i in range(depth):
                p = 'backbone.stage%d_%d' % (stage, i)
                stride = 2 if (i == 0 and stage > 2) else 1
                if r50:
                    x = self.bottleneck(p, x, stride, i == 0, stage == 2)
                else:
                    x = self.basic(p, x, stride, i == 0 and stage == 2)
            if stage in self.cfg.backbone['feature_maps']:
                feats.append(x)
        return feats

    # -- head (model/head.py) ------------------------------------------------------------------
    def detection_block(self, i, x):
        """DetectionBlock.__call__ :223-231 with the layer list built at :175-221."""
        h = self.cfg.head
        coord = h.get('coord_conv', True)
        nblk = h.get('conv_block_num', 2)
        use_spp = h.get('spp', True)
        drop = h.get('drop_block', True)
        first = i == 0
        p = 'head.detection_blocks.%d' % i
        j = 0

        def cc(t):
            return coord_concat(t) if coord else t
        for blk in range(nblk):
            x = self.unit('%s.layers.%d' % (p, j + 1), cc(x), 1, 'leaky')
            j += 2
            if use_spp and first and blk == 1:
                x = self.unit('%s.layers.%d' % (p, j + 1), spp(x), 1, 'leaky')
                x = self.unit('%s.layers.%d' % (p, j + 2), x, 1, 'leaky')
                j += 3
            else:
                x = self.unit('%s.layers.%d' % (p, j), x, 1, 'leaky')
                j += 1
            if drop and blk == 0 and not first:
                j += 1                                       # DropBlock is identity at inference
        if drop and first:
            j += 1
        route = self.unit('%s.layers.%d' % (p, j + 1), cc(x), 1, 'leaky')
        tip = self.unit('%s.tip_layers.1' % p, cc(route), 1, 'leaky')
        return route, tip

    def head_outputs(self, feats):
        """YOLOv3Head._get_outputs :381-398."""
        n_out = len(self.cfg.head['anchor_masks'])
        blocks = feats[-1:-n_out - 1:-1]
        outs, route = [], None
        for i, blk in enumerate(blocks):
            if i > 0:
                blk = torch.cat([route, blk], dim=1)
            route, tip = self.detection_block(i, blk)
            outs.append(self.unit('head.yolo_output_convs.%d' % i, tip, 1, None))
            if i < n_out - 1:
                route = self.unit('head.upsample_layers.%d' % (2 * i), route, 1, 'leaky')
                route = F.interpolate(route, scale_factor=2, mode='nearest')
        return outs

    def decode(self, outs, im_size):
        """YOLOv3Head.get_prediction :439-453."""
        h = self.cfg.head
        anchors = np.asarray(h['anchors'], dtype=np.float32)
        boxes, scores = [], []
        for i, out in enumerate(outs):
            mask = h['anchor_masks'][i]
            if h.get('iou_aware', True):
                out = iou_aware_score(out, len(mask), h['num_classes'], h.get('iou_aware_factor', 0.4))
            b, s = yolo_box(out, anchors[mask], h['downsample'][i], h['num_classes'], h.get('scale_x_y', 1.05),
                            im_size, h.get('clip_bbox', True))
            boxes.append(b)
            scores.append(s)
        return torch.cat(boxes, dim=1), torch.cat(scores, dim=1)

    def forward(self, x, im_size, return_all=False):
        """PPYOLO.forward(eval=True), model/ppyolo.py:19-22 -> list of [M,6] numpy arrays."""
        boxes, scores, feats, outs = self.forward_dense(x, im_size)
        nms = {k: v for k, v in self.cfg.nms_cfg.items() if k != 'nms_type'}
        preds = [matrix_nms(boxes[i].cpu().numpy(), scores[i].cpu().numpy(), **nms) for i in range(boxes.shape[0])]
        if return_all:
            return dict(feats=feats, outs=outs, boxes=boxes, scores=scores, preds=preds)
        return preds


def _forward_dense(self, x, im_size):
    """Backbone + head + box decode (everything before the per-image NMS loop), on ``self.device``."""
    with torch.no_grad():
        x = x.to(self.device, self.dtype)
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        feats = self.backbone(x)
        outs = [o.float() for o in self.head_outputs(feats)]
        boxes, scores = self.decode(outs, im_size.float().to(self.device))
    return boxes, scores, feats, outs


Net.forward_dense = _forward_dense


def ema_update(shadow, params, decay, update_step, thres_steps=True):
    """ExponentialMovingAverage.update, model/EMA.py:31-45, on numpy float32 arrays: returns (new shadows, decay used).
    ``decay * old + (1 - decay) * new`` with a Python-float decay keeps float32 arrays float32 (numpy weak scalars)."""
    d = min(decay, (1 + update_step) / (10 + update_step)) if thres_steps else decay
    return [np.asarray(d * s + (1 - d) * np.asarray(p, dtype=np.float32), dtype=np.float32) for s, p in zip(shadow, params)], d


# ----------------------------------------------------------------------------------------------
# pre-processing: the resize of model/decode_np.py:125-134 (cv2.resize(..., interpolation=cv2.INTER_CUBIC) of a uint8 image)
# ----------------------------------------------------------------------------------------------
def _cubic_axis(ssize, dsize):
    """Per destination index: the 4 clamped source indices and 16-bit fixed-point weights of OpenCV's bicubic resize
    (third-party dependency of the reference: opencv-python, imgproc/src/resize.cpp resizeGeneric_ / interpolateCubic,
    A = -0.75; float32 arithmetic in OpenCV's operation order, weights * 2048 rounded to nearest-even)."""
    f32 = np.float32
    scale = 1.0 / (float(dsize) / float(ssize))                       # cv2.resize(fx = dsize / ssize): scale = 1 / fx (double)
    f = ((np.arange(dsize, dtype=np.float64) + 0.5) * scale - 0.5).astype(f32)
    s = np.floor(f).astype(np.int64)
    x = (f - s.astype(f32)).astype(f32)
    a = f32(-0.75)
    one = f32(1)
    c0 = ((a * (x + one) - f32(5) * a) * (x + one) + f32(8) * a) * (x + one) - f32(4) * a
    c1 = ((a + f32(2)) * x - (a + f32(3))) * x * x + one
    c2 = ((a + f32(2)) * (one - x) - (a + f32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    w = np.clip(np.rint(np.stack([c0, c1, c2, c3], 1).astype(f32) * f32(2048)), -32768, 32767).astype(np.int32)
    idx = np.clip(s[:, None] - 1 + np.arange(4)[None, :], 0, ssize - 1)
    return idx, w


def resize_cubic_u8(img, size):
    """cv2.resize(img, None, None, fx=size/w, fy=size/h, interpolation=cv2.INTER_CUBIC) for an HWC uint8 image as OpenCV's own
    code computes it (HResizeCubic in 32-bit integers, VResizeCubicVec_32s8u in float32 with separately rounded products and
    sums -- the baseline code path has no fused multiply-add -- then round-to-nearest-even and saturation)."""
    h, w, _ = img.shape
    xi, xw = _cubic_axis(w, size)
    yi, yw = _cubic_axis(h, size)
    hor = np.einsum('hdkc,dk->hdc', img.astype(np.int32)[:, xi, :], xw)               # [h, size, c] int32
    scale = np.float32(1.0 / (2048.0 * 2048.0))
    b = yw.astype(np.float32) * scale                                                # [size, 4]
    rows = hor[yi].astype(np.float32)                                                # [size, 4, size, c]
    acc = rows[:, 3] * b[:, 3, None, None]
    for k in (2, 1, 0):
        acc = (rows[:, k] * b[:, k, None, None]).astype(np.float32) + acc
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)

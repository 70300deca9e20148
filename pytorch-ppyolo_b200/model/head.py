"""YOLOv3 head of PP-YOLO (CoordConv / SPP / DropBlock / IoU-aware) on the B200 kernel set.

Surface-compatible with the reference's ``model/head.py``: ``yolo_box`` :21-80,
``get_iou_aware_score`` :138-141, ``DetectionBlock`` :146-239, ``YOLOv3Head``
:242-469 (same kwargs, ModuleList indexing and state_dict keys).  Differences
in *how* it runs: IoU-aware fusion + box decode + score product are one kernel
per scale, and Matrix-NMS runs as one batched launch sequence over all images
instead of the per-image Python loop at reference :462-464.
"""
import copy

import numpy as np
import torch

from ppyolo_b200 import ops
from model.custom_layers import Conv2dUnit, CoordConv, SPP, DropBlock, get_norm
from model.matrix_nms import matrix_nms  # noqa: F401  (re-exported like the reference)


def yolo_box(conv_output, anchors, stride, num_classes, scale_x_y, im_size, clip_bbox, conf_thresh):
    """Decode one scale: NCHW head output -> (boxes [N,H*W*A,4] xyxy in image pixels, scores [N,H*W*A,C]).

    Box order is (h, w, anchor) as at reference :58.  ``conf_thresh`` is unused, as in the reference.
    """
    return ops.yolo_box(conv_output, np.asarray(anchors, dtype=np.float32).reshape(-1), stride, num_classes,
                        scale_x_y, im_size, clip_bbox, iou_aware=False, iou_aware_factor=0.0)


def get_iou_aware_score(output, an_num, num_classes, iou_aware_factor):
    """[N, A*(6+C), H, W] -> [N, A*(5+C), H, W] with obj logit replaced by logit(obj^(1-f) * ioup^f)."""
    return ops.iou_aware_score(output, an_num, num_classes, iou_aware_factor)


class DetectionBlock(torch.nn.Module):
    def __init__(self, in_c, channel, coord_conv=True, bn=0, gn=0, af=0, norm_decay=0., conv_block_num=2,
                 is_first=False, use_spp=True, drop_block=True, block_size=3, keep_prob=0.9, is_test=True, name=''):
        super().__init__()
        assert channel % 2 == 0, "channel {} cannot be divided by 2".format(channel)
        self.norm_decay, self.use_spp, self.coord_conv = norm_decay, use_spp, coord_conv
        self.is_first, self.is_test, self.drop_block = is_first, is_test, drop_block
        self.block_size, self.keep_prob = block_size, keep_prob
        extra = 2 if coord_conv else 0
        kw = dict(bn=bn, gn=gn, af=af, act='leaky', norm_decay=norm_decay)

        def unit(cin, cout, k, tag):
            return Conv2dUnit(cin, cout, k, stride=1, name='{}.{}'.format(name, tag), **kw)

        layers = []
        for j in range(conv_block_num):
            layers += [CoordConv(coord_conv), unit(in_c + extra, channel, 1, '%d.0' % j)]
            if use_spp and is_first and j == 1:
                layers += [SPP(), unit(channel * 4, 512, 1, '%d.spp.conv' % j), unit(512, channel * 2, 3, '%d.1' % j)]
            else:
                layers.append(unit(channel, channel * 2, 3, '%d.1' % j))
            if drop_block and j == 0 and not is_first:
                layers.append(DropBlock(block_size=block_size, keep_prob=keep_prob, is_test=is_test))
            in_c = channel * 2
        if drop_block and is_first:
            layers.append(DropBlock(block_size=block_size, keep_prob=keep_prob, is_test=is_test))
        route_in = in_c if conv_block_num == 0 else channel * 2
        layers += [CoordConv(coord_conv), unit(route_in + extra, channel, 1, '2')]
        self.layers = torch.nn.ModuleList(layers)
        self.tip_layers = torch.nn.ModuleList([CoordConv(coord_conv), unit(channel + extra, channel * 2, 3, 'tip')])

    def forward(self, x):
        for ly in self.layers:
            x = ly(x)
        tip = x
        for ly in self.tip_layers:
            tip = ly(tip)
        return x, tip

    def add_param_group(self, param_groups, base_lr, base_wd):
        for ly in list(self.layers) + list(self.tip_layers):
            if isinstance(ly, Conv2dUnit):
                ly.add_param_group(param_groups, base_lr, base_wd)


class YOLOv3Head(torch.nn.Module):
    def __init__(self, conv_block_num=2, num_classes=80,
                 anchors=[[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198],
                          [373, 326]],
                 anchor_masks=[[6, 7, 8], [3, 4, 5], [0, 1, 2]], norm_type="bn", norm_decay=0., coord_conv=True,
                 iou_aware=True, iou_aware_factor=0.4, block_size=3, scale_x_y=1.05, spp=True, drop_block=True,
                 keep_prob=0.9, clip_bbox=True, yolo_loss=None, downsample=[32, 16, 8],
                 in_channels=[2048, 1024, 512], nms_cfg=None, focalloss_on_obj=False, prior_prob=0.01,
                 is_train=False):
        super().__init__()
        self.conv_block_num, self.num_classes = conv_block_num, num_classes
        self.norm_type, self.norm_decay = norm_type, norm_decay
        self.coord_conv, self.iou_aware, self.iou_aware_factor = coord_conv, iou_aware, iou_aware_factor
        self.scale_x_y, self.use_spp, self.drop_block, self.keep_prob = scale_x_y, spp, drop_block, keep_prob
        self.clip_bbox, self.anchors, self.anchor_masks = clip_bbox, anchors, anchor_masks
        self.block_size, self.downsample, self.in_channels = block_size, downsample, in_channels
        self.yolo_loss, self.nms_cfg = yolo_loss, nms_cfg
        self.focalloss_on_obj, self.prior_prob, self.is_train = focalloss_on_obj, prior_prob, is_train
        self._anchors = np.array(copy.deepcopy(anchors)).astype(np.float32)  # [num_anchors, 2]
        self.mask_anchors = [[v for aid in m for v in anchors[aid]] for m in anchor_masks]

        assert norm_type in ['bn', 'sync_bn', 'gn', 'affine_channel']
        bn, gn, af = get_norm(norm_type)
        n_out = len(downsample)
        self.detection_blocks = torch.nn.ModuleList()
        self.yolo_output_convs = torch.nn.ModuleList()
        self.upsample_layers = torch.nn.ModuleList()
        for i in range(n_out):
            channel = 64 * (2 ** n_out) // (2 ** i)
            in_c = in_channels[i] + (512 // (2 ** i) if i > 0 else 0)
            self.detection_blocks.append(DetectionBlock(
                in_c=in_c, channel=channel, coord_conv=coord_conv, bn=bn, gn=gn, af=af, norm_decay=norm_decay,
                is_first=(i == 0), conv_block_num=conv_block_num, use_spp=spp, drop_block=drop_block,
                block_size=block_size, keep_prob=keep_prob, is_test=(not is_train), name="yolo_block.{}".format(i)))
            per_anchor = num_classes + (6 if iou_aware else 5)
            self.yolo_output_convs.append(Conv2dUnit(channel * 2, len(anchor_masks[i]) * per_anchor, 1, stride=1,
                                                     bias_attr=True, act=None,
                                                     name="yolo_output.{}.conv".format(i)))
            if i < n_out - 1:
                self.upsample_layers.append(Conv2dUnit(channel, 256 // (2 ** i), 1, stride=1, bn=bn, gn=gn, af=af,
                                                       act='leaky', norm_decay=norm_decay,
                                                       name="yolo_transition.{}".format(i)))
                self.upsample_layers.append(torch.nn.Upsample(scale_factor=2, mode='nearest'))  # container

    def add_param_group(self, param_groups, base_lr, base_wd):
        for blk in self.detection_blocks:
            blk.add_param_group(param_groups, base_lr, base_wd)
        for ly in list(self.yolo_output_convs) + list(self.upsample_layers):
            if isinstance(ly, Conv2dUnit):
                ly.add_param_group(param_groups, base_lr, base_wd)

    def set_dropblock(self, is_test):
        for blk in self.detection_blocks:
            for ly in blk.layers:
                if isinstance(ly, DropBlock):
                    ly.is_test = is_test

    def _get_outputs(self, body_feats):
        n_out = len(self.anchor_masks)
        feats = body_feats[-1:-n_out - 1:-1]
        outputs, route = [], None
        for i, feat in enumerate(feats):
            if i > 0:
                feat = ops.upsample2x_concat(route, feat)  # nearest x2 of route, then channel concat
            route, tip = self.detection_blocks[i](feat)
            outputs.append(self.yolo_output_convs[i](tip))
            if i < n_out - 1:
                route = self.upsample_layers[2 * i](route)
        return outputs

    def get_loss(self, input, gt_box, gt_label, gt_score, targets):
        """Reference model/head.py:400-422.  The module-level operators of ``_get_outputs`` run kernels on detached tensors, so the
        losses are computed on the differentiable evaluation of the head instead (same values, gradients for every head
        parameter) -- never a silently non-differentiable dict."""
        return self.get_loss_autograd(input, gt_box, gt_label, gt_score, targets)

    def get_loss_autograd(self, body_feats, gt_box, gt_label, gt_score, targets):
        """get_loss (reference :400-422) with the head evaluated as differentiable tensor code."""
        from ppyolo_b200 import autograd_head
        if self.yolo_loss is None:
            raise RuntimeError('YOLOv3Head was built without yolo_loss; pass the YOLOv3Loss object like train.py:246-252')
        impl = getattr(self, 'train_impl', None) or 'aten'         # 'kernels': convs (fwd, dgrad, wgrad) on the tcgen05 kernel
        outputs = autograd_head.head_outputs(self, body_feats, impl)
        return self.yolo_loss(outputs, gt_box, gt_label, gt_score, targets, self.anchors, self.anchor_masks,
                              self.mask_anchors, self.num_classes)

    def decode_outputs(self, outputs, im_size):
        """Fused IoU-aware + yolo_box over all scales -> (boxes [N,B,4], scores [N,B,C])."""
        boxes, scores = [], []
        for i, out in enumerate(outputs):
            anchors = self._anchors[self.anchor_masks[i]].reshape(-1)
            b, s = ops.yolo_box(out, anchors, self.downsample[i], self.num_classes, self.scale_x_y, im_size,
                                self.clip_bbox, iou_aware=self.iou_aware, iou_aware_factor=self.iou_aware_factor)
            boxes.append(b)
            scores.append(s)
        return torch.cat(boxes, dim=1), torch.cat(scores, dim=1)

    def get_prediction(self, body_feats, im_size):
        outputs = self._get_outputs(body_feats)
        yolo_boxes, yolo_scores = self.decode_outputs(outputs, im_size)
        nms_cfg = copy.deepcopy(self.nms_cfg)
        nms_type = nms_cfg.pop('nms_type')
        if nms_type != 'matrix_nms':
            raise NotImplementedError(nms_type)
        return ops.matrix_nms_batched(yolo_boxes, yolo_scores, **nms_cfg)

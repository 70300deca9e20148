"""PP-YOLO r18vd: ResNet18-vd, 2-scale head without CoordConv/SPP/IoU-aware (reference config/ppyolo_r18vd.py)."""
from ._base import _PPYOLOConfigBase, COCO_ANCHORS_6

__all__ = ['PPYOLO_r18vd_Config']


class PPYOLO_r18vd_Config(_PPYOLOConfigBase):
    model_file = 'ppyolo_r18vd.pt'
    target_size = 416

    def _model_section(self):
        masks = [[3, 4, 5], [0, 1, 2]]
        self.backbone_type = 'Resnet18Vd'
        self.backbone = dict(norm_type='bn', feature_maps=[4, 5], dcn_v2_stages=[], freeze_at=5, freeze_norm=False,
                             norm_decay=0.)
        self.head_type = 'YOLOv3Head'
        self.head = dict(num_classes=self.num_classes, conv_block_num=0, norm_type='bn', anchor_masks=masks,
                         anchors=[list(a) for a in COCO_ANCHORS_6], coord_conv=False, iou_aware=False,
                         iou_aware_factor=0.4, scale_x_y=1.05, spp=False, drop_block=True, keep_prob=0.9,
                         downsample=[32, 16], in_channels=[512, 256])
        self.gt2YoloTarget = dict(anchor_masks=masks, anchors=[list(a) for a in COCO_ANCHORS_6],
                                  downsample_ratios=[32, 16], num_classes=self.num_classes)
        # shorter schedule than the 2x model (reference config/ppyolo_r18vd.py:43,50-53)
        self.train_cfg['max_iters'] = 250000
        self.learningRate['PiecewiseDecay']['milestones'] = [150000, 200000]

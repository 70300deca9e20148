"""PP-YOLO 2x: ResNet50-vd + DCNv2 in stage 5, 3-scale head with CoordConv/SPP/IoU-aware.

Values follow the reference config/ppyolo_2x.py:94-151 (model section) and :183-217.
"""
from ._base import _PPYOLOConfigBase, COCO_ANCHORS_9

__all__ = ['PPYOLO_2x_Config']


class PPYOLO_2x_Config(_PPYOLOConfigBase):
    model_file = 'ppyolo_2x.pt'
    target_size = 608

    def _model_section(self):
        masks = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]
        self.backbone_type = 'Resnet50Vd'
        self.backbone = dict(norm_type='bn', feature_maps=[3, 4, 5], dcn_v2_stages=[5], downsample_in3x3=True,
                             freeze_at=5, freeze_norm=False, norm_decay=0.)
        self.head_type = 'YOLOv3Head'
        self.head = dict(num_classes=self.num_classes, norm_type='bn', anchor_masks=masks,
                         anchors=[list(a) for a in COCO_ANCHORS_9], coord_conv=True, iou_aware=True,
                         iou_aware_factor=0.4, scale_x_y=1.05, spp=True, drop_block=True, keep_prob=0.9,
                         downsample=[32, 16, 8], in_channels=[2048, 1024, 512])
        self.iou_aware_loss_type = 'IouAwareLoss'
        self.iou_aware_loss = dict(loss_weight=1.0, max_height=608, max_width=608)
        self.gt2YoloTarget = dict(anchor_masks=masks, anchors=[list(a) for a in COCO_ANCHORS_9],
                                  downsample_ratios=[32, 16, 8], num_classes=self.num_classes)

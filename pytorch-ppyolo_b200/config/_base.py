"""Shared defaults of the two PP-YOLO configs; attribute names follow the reference config classes."""

COCO_ANCHORS_9 = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
COCO_ANCHORS_6 = [[10, 14], [23, 27], [37, 58], [81, 82], [135, 169], [344, 319]]
MULTISCALE = list(range(320, 609, 32))


class _PPYOLOConfigBase(object):
    model_file = 'ppyolo.pt'
    target_size = 608

    def __init__(self):
        self.train_path = '../COCO/annotations/instances_train2017.json'
        self.val_path = '../COCO/annotations/instances_val2017.json'
        self.classes_path = 'data/coco_classes.txt'
        self.train_pre_path = '../COCO/train2017/'
        self.val_pre_path = '../COCO/val2017/'
        self.test_path = '../COCO/annotations/image_info_test-dev2017.json'
        self.test_pre_path = '../COCO/test2017/'
        self.num_classes = 80

        self.train_cfg = dict(batch_size=8, num_threads=5, max_batch=3, model_path=self.model_file, save_iter=1000,
                              eval_iter=5000, max_iters=500000, mixup_epoch=10, cutmix_epoch=-1)
        self.learningRate = dict(base_lr=0.0001,
                                 PiecewiseDecay=dict(gamma=0.1, milestones=[400000, 450000]),
                                 LinearWarmup=dict(start_factor=0., steps=4000))
        self.optimizerBuilder = dict(optimizer=dict(momentum=0.9, type='Momentum'),
                                     regularizer=dict(factor=0.0005, type='L2'))
        self.eval_cfg = dict(model_path=self.model_file, target_size=self.target_size, draw_image=False,
                             draw_thresh=0.15, eval_batch_size=4)
        self.test_cfg = dict(model_path=self.model_file, target_size=self.target_size, draw_image=True,
                             draw_thresh=0.15)
        self.use_ema = True
        self.ema_decay = 0.9998

        self.iou_loss_type = 'IouLoss'
        self.iou_loss = dict(loss_weight=2.5, max_height=608, max_width=608, ciou_term=False)
        self.yolo_loss_type = 'YOLOv3Loss'
        self.yolo_loss = dict(ignore_thresh=0.7, scale_x_y=1.05, label_smooth=False, use_fine_grained_loss=True)
        self.nms_cfg = dict(nms_type='matrix_nms', score_threshold=0.01, post_threshold=0.01, nms_top_k=500,
                            keep_top_k=100, use_gaussian=False, gaussian_sigma=2.)

        # pre-processing (host side; consumed by Decode / the training reader)
        self.context = {'fields': ['image', 'gt_bbox', 'gt_class', 'gt_score']}
        self.decodeImage = dict(to_rgb=True, with_mixup=True, with_cutmix=False)
        self.mixupImage = dict(alpha=1.5, beta=1.5)
        self.colorDistort = dict()
        self.randomExpand = dict(fill_value=[123.675, 116.28, 103.53])
        self.randomCrop = dict()
        self.randomFlipImage = dict(is_normalized=False)
        self.normalizeBox = dict()
        self.padBox = dict(num_max_boxes=50)
        self.bboxXYXY2XYWH = dict()
        self.randomShape = dict(sizes=list(MULTISCALE), random_inter=True)
        self.normalizeImage = dict(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], is_scale=True,
                                   is_channel_first=False)
        self.permute = dict(to_bgr=False, channel_first=True)
        self.resizeImage = dict(target_size=608, interp=2)
        self.sample_transforms_seq = ['decodeImage', 'mixupImage', 'colorDistort', 'randomExpand', 'randomCrop',
                                      'randomFlipImage', 'normalizeBox', 'padBox', 'bboxXYXY2XYWH']
        self.batch_transforms_seq = ['randomShape', 'normalizeImage', 'permute', 'gt2YoloTarget']
        self._model_section()

    def _model_section(self):
        raise NotImplementedError

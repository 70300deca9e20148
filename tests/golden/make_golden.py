"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the authoring container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The reference hard-codes ``.cuda()`` (model/head.py:43); the single shim below makes it an identity so
the reference runs on CPU, as described in SURVEY.md 0 / 8c.  Weights and inputs come from
``ppyolo_b200.synth`` so the tests can rebuild them bit-identically from the seed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'

torch.Tensor.cuda = lambda self, *a, **k: self        # the one shim
torch.set_num_threads(1)                               # fixed summation order

sys.path.insert(0, REF)
import importlib.util  # noqa: E402


def _load_synth():
    spec = importlib.util.spec_from_file_location('synth', os.path.join(REPO, 'pytorch-ppyolo_b200', 'ppyolo_b200',
                                                                        'synth.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synth = _load_synth()

from config import PPYOLO_2x_Config, PPYOLO_r18vd_Config, select_backbone, select_head  # noqa: E402
from model.ppyolo import PPYOLO  # noqa: E402
from model import custom_layers as ref_layers  # noqa: E402
from model import head as ref_head  # noqa: E402
from model.matrix_nms import matrix_nms as ref_matrix_nms, jaccard as ref_jaccard  # noqa: E402

NMS_CFG = dict(score_threshold=0.01, post_threshold=0.01, nms_top_k=500, keep_top_k=100)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-22s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


def gen_nms():
    arrays = {}
    t = torch.tensor
    # known-answer cases of SURVEY.md 8c
    kat = {
        'zeros': (t([[0., 0., 10., 10.], [1., 1., 5., 5.]]), torch.zeros(2, 80)),
        'dups': (t([[0., 0., 10., 10.]] * 3), None),
        'chain': (t([[0., 0., 10., 10.], [5., 0., 15., 10.], [8., 0., 18., 10.]]), None),
    }
    s = torch.zeros(3, 80); s[0, 5], s[1, 5], s[2, 7] = 0.9, 0.8, 0.7
    kat['dups'] = (kat['dups'][0], s)
    s = torch.zeros(3, 80); s[0, 1], s[1, 1], s[2, 1] = 0.9, 0.8, 0.7
    kat['chain'] = (kat['chain'][0], s)
    for name, (b, s) in kat.items():
        for gauss in (False, True):
            out = ref_matrix_nms(b, s, use_gaussian=gauss, gaussian_sigma=2.0, **NMS_CFG)
            arrays['%s_%s_boxes' % (name, 'g' if gauss else 'l')] = b
            arrays['%s_%s_scores' % (name, 'g' if gauss else 'l')] = s
            arrays['%s_%s_out' % (name, 'g' if gauss else 'l')] = out
    # random cases: (boxes, classes, seed, top_k, keep)
    for nb, nc, seed, topk, keep in ((400, 80, 0, 500, 100), (400, 80, 1, 100, 20), (3000, 80, 2, 500, 100),
                                     (64, 3, 3, -1, 100), (1500, 20, 4, 300, 50)):
        b, s = synth.nms_inputs(nb, nc, seed=seed)
        for gauss in (False, True):
            cfg = dict(score_threshold=0.01, post_threshold=0.01, nms_top_k=topk, keep_top_k=keep)
            out = ref_matrix_nms(b, s, use_gaussian=gauss, gaussian_sigma=2.0, **cfg)
            tag = 'rand_b%d_c%d_s%d_t%d_k%d_%s' % (nb, nc, seed, topk, keep, 'g' if gauss else 'l')
            arrays[tag + '_out'] = out
    # a post-threshold that actually bites, and a degenerate (zero-area) box pair
    b, s = synth.nms_inputs(800, 80, seed=5)
    arrays['post05_out'] = ref_matrix_nms(b, s, score_threshold=0.05, post_threshold=0.2, nms_top_k=200, keep_top_k=30)
    b = t([[3., 3., 3., 3.], [3., 3., 3., 3.], [0., 0., 4., 4.], [1., 1., 5., 5.]])
    s = torch.zeros(4, 80); s[0, 2], s[1, 2], s[2, 2], s[3, 4] = 0.9, 0.8, 0.7, 0.6
    arrays['degenerate_boxes'] = b
    arrays['degenerate_scores'] = s
    arrays['degenerate_out'] = ref_matrix_nms(b, s, **NMS_CFG)
    ba, _ = synth.nms_inputs(37, 1, seed=7)
    bb, _ = synth.nms_inputs(53, 1, seed=8)
    arrays['jaccard_out'] = ref_jaccard(ba, bb)
    save('nms', **arrays)


def gen_decode():
    arrays = {}
    g = torch.Generator().manual_seed(11)
    anchors9 = np.array(PPYOLO_2x_Config().head['anchors'], dtype=np.float32)
    for tag, size, stride, mask, iou_aware in (('s32', 5, 32, [6, 7, 8], True), ('s8', 12, 8, [0, 1, 2], True),
                                               ('plain', 7, 16, [3, 4, 5], False)):
        ch = 3 * (86 if iou_aware else 85)
        x = torch.randn((2, ch, size, size), generator=g) * 1.5
        x[0, :, 0, 0] *= 8.0     # drive a few logits into the clamp ranges of _de_sigmoid / exp
        im_size = torch.tensor([[480., 640.], [375., 500.]])
        y = ref_head.get_iou_aware_score(x, 3, 80, 0.4) if iou_aware else x
        for clip in (True, False):
            boxes, scores = ref_head.yolo_box(y, anchors9[mask], stride, 80, 1.05, im_size, clip, 0.01)
            arrays['%s_boxes_clip%d' % (tag, int(clip))] = boxes
        arrays[tag + '_in'] = x
        arrays[tag + '_iouaware'] = y if iou_aware else np.zeros(1, np.float32)
        arrays[tag + '_scores'] = scores
        arrays[tag + '_im_size'] = im_size
    save('decode', **arrays)


def gen_layers():
    arrays = {}
    g = torch.Generator().manual_seed(21)
    arrays['coord_out'] = ref_layers.CoordConv(True)(torch.zeros(1, 2, 3, 4))
    x = torch.randn((2, 4, 15, 15), generator=g)
    arrays['spp_in'] = x
    arrays['spp_out'] = ref_layers.SPP()(x)
    for tag, cin, cout, k, stride, act, bias in (('c3s1', 8, 16, 3, 1, 'leaky', False), ('c3s2', 8, 16, 3, 2, 'relu', False),
                                                 ('c1s1', 16, 24, 1, 1, None, True), ('c1s2', 8, 8, 1, 2, 'relu', False)):
        u = ref_layers.Conv2dUnit(cin, cout, k, stride=stride, bias_attr=bias, bn=0 if bias else 1, act=act)
        synth.randomize_(u, seed=30)
        u.eval()
        x = torch.randn((2, cin, 10, 10), generator=g)
        arrays[tag + '_in'] = x
        arrays[tag + '_out'] = u(x)
    for tag, stride, hw in (('dcn_s1', 1, 9), ('dcn_s2', 2, 10)):
        u = ref_layers.Conv2dUnit(16, 24, 3, stride=stride, bn=1, act='relu', use_dcn=True)
        synth.randomize_(u, seed=31, offset_scale=0.05)
        u.eval()
        x = torch.randn((2, 16, hw, hw), generator=g)
        arrays[tag + '_in'] = x
        arrays[tag + '_raw'] = u.conv(x)
        arrays[tag + '_out'] = u(x)
        arrays[tag + '_offsetmask'] = u.conv.conv_offset(x)
    save('layers', **arrays)


def build(cfg):
    backbone = select_backbone(cfg.backbone_type)(**cfg.backbone)
    head = select_head(cfg.head_type)(yolo_loss=None, nms_cfg=cfg.nms_cfg, **cfg.head)
    model = PPYOLO(backbone, head)
    synth.randomize_(model, seed=0)
    model.eval()
    head.set_dropblock(is_test=True)
    return model


def gen_net(name, cfg, size, batch=2):
    model = build(cfg)
    x = synth.images(batch, size, seed=1)
    im_size = torch.tensor([[480., 640.], [333., 500.]])[:batch]
    with torch.no_grad():
        feats = model.backbone(x)
        outs = model.head._get_outputs(feats)
        boxes, scores = [], []
        for i, out in enumerate(outs):
            o = out
            if model.head.iou_aware:
                o = ref_head.get_iou_aware_score(o, 3, 80, model.head.iou_aware_factor)
            b, s = ref_head.yolo_box(o, model.head._anchors[model.head.anchor_masks[i]], model.head.downsample[i], 80,
                                     model.head.scale_x_y, im_size, model.head.clip_bbox, 0.01)
            boxes.append(b)
            scores.append(s)
        boxes, scores = torch.cat(boxes, 1), torch.cat(scores, 1)
        preds = model(x, im_size)
    arrays = {'out%d' % i: o for i, o in enumerate(outs)}
    arrays['feat_last'] = feats[-1]
    arrays['boxes'] = boxes
    arrays['scores'] = scores.half()  # storage only: full-precision scores are re-derivable from out*
    arrays['scores_img0_f32'] = scores[0]
    arrays['im_size'] = im_size
    arrays['x_checksum'] = np.array([float(x.double().sum()), float(x.double().abs().sum())])
    w = torch.cat([v.double().flatten() for k, v in sorted(model.state_dict().items())])
    arrays['w_checksum'] = np.array([float(w.sum()), float(w.abs().sum())])
    for i, p in enumerate(preds):
        arrays['pred%d' % i] = p
    save(name, **arrays)
    frac = float((scores > 0.01).float().mean())
    print('   %s: %d boxes/img, %.2f%% of scores > 0.01, preds %s' % (name, boxes.shape[1], 100 * frac,
                                                                      [tuple(p.shape) for p in preds]))


def gen_losses():
    """Training losses (model/losses.py, model/iou_losses.py) and targets (tools/transform.py:1318-1421) of the reference."""
    from config import select_loss
    from tools.transform import Gt2YoloTargetSingle
    spec = importlib.util.spec_from_file_location('targets', os.path.join(REPO, 'pytorch-ppyolo_b200', 'ppyolo_b200', 'targets.py'))
    tg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tg)
    arrays = {}
    for tag, cfg in (('r50vd', PPYOLO_2x_Config()), ('r18vd', PPYOLO_r18vd_Config())):
        size, batch = 128, 2
        gt_bbox, gt_class, gt_score = tg.synthetic_ground_truth(batch, seed=3)
        op = Gt2YoloTargetSingle(**cfg.gt2YoloTarget)
        per_image = []
        for b in range(batch):
            sample = {'image': np.zeros((3, size, size), np.float32), 'gt_bbox': gt_bbox[b], 'gt_class': gt_class[b],
                      'gt_score': gt_score[b]}
            per_image.append(op(sample))
        n_scale = len(cfg.head['anchor_masks'])
        targets = [np.stack([s['target%d' % i] for s in per_image]) for i in range(n_scale)]
        iou_loss = select_loss(cfg.iou_loss_type)(**cfg.iou_loss)
        iou_aware = select_loss(cfg.iou_aware_loss_type)(**cfg.iou_aware_loss) if cfg.head['iou_aware'] else None
        yolo = select_loss(cfg.yolo_loss_type)(iou_loss=iou_loss, iou_aware_loss=iou_aware, **cfg.yolo_loss)
        g = torch.Generator().manual_seed(41)
        per = 86 if cfg.head['iou_aware'] else 85
        outs = [(torch.randn((batch, 3 * per, size // s, size // s), generator=g) * 1.2).requires_grad_(True)
                for s in cfg.head['downsample']]
        anchors, masks = cfg.head['anchors'], cfg.head['anchor_masks']
        mask_anchors = [[v for aid in m for v in anchors[aid]] for m in masks]
        losses = yolo(outs, torch.from_numpy(gt_bbox), torch.from_numpy(gt_class), torch.from_numpy(gt_score),
                      [torch.from_numpy(t) for t in targets], anchors, masks, mask_anchors, 80)
        total = sum(losses.values())
        total.backward()
        for k, v in losses.items():
            arrays['%s_%s' % (tag, k)] = v.detach()
        for i, (o, t) in enumerate(zip(outs, targets)):
            arrays['%s_target%d' % (tag, i)] = t
            arrays['%s_grad%d' % (tag, i)] = o.grad if i < n_scale - 1 or n_scale == 2 else np.zeros(1, np.float32)
            arrays['%s_gradsum%d' % (tag, i)] = np.array([float(o.grad.double().sum()), float(o.grad.double().abs().sum())])
        print('   losses %s:' % tag, {k: round(float(v), 4) for k, v in losses.items()})
    save('losses', **arrays)


def gen_train():
    """One training forward+backward of the reference (train.py:427-441) with its default frozen backbone in TRAIN mode
    (batch-statistic BN), DropBlock disabled for determinism."""
    from config import select_loss
    spec = importlib.util.spec_from_file_location('targets', os.path.join(REPO, 'pytorch-ppyolo_b200', 'ppyolo_b200', 'targets.py'))
    tg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tg)
    arrays = {}
    for tag, cfg in (('r50vd', PPYOLO_2x_Config()), ('r18vd', PPYOLO_r18vd_Config())):
        size, batch = 128, 2
        iou_loss = select_loss(cfg.iou_loss_type)(**cfg.iou_loss)
        iou_aware = select_loss(cfg.iou_aware_loss_type)(**cfg.iou_aware_loss) if cfg.head['iou_aware'] else None
        yolo = select_loss(cfg.yolo_loss_type)(iou_loss=iou_loss, iou_aware_loss=iou_aware, **cfg.yolo_loss)
        head_kw = dict(cfg.head); head_kw['drop_block'] = False
        backbone = select_backbone(cfg.backbone_type)(**cfg.backbone)
        head = select_head(cfg.head_type)(yolo_loss=yolo, is_train=True, nms_cfg=cfg.nms_cfg, **head_kw)
        model = PPYOLO(backbone, head)
        synth.randomize_(model, seed=0)
        model.train()
        backbone.freeze()
        x = synth.images(batch, size, seed=1)
        gt_bbox, gt_class, gt_score = tg.synthetic_ground_truth(batch, seed=3)
        targets = tg.gt2yolo_target(gt_bbox, gt_class, gt_score, h=size, w=size, **cfg.gt2YoloTarget)
        losses = model(x, None, False, torch.from_numpy(gt_bbox), torch.from_numpy(gt_class), torch.from_numpy(gt_score),
                       [torch.from_numpy(t) for t in targets])
        sum(losses.values()).backward()
        for k, v in losses.items():
            arrays['%s_%s' % (tag, k)] = v.detach()
        sd = model.state_dict()
        arrays[tag + '_stem_running_mean'] = sd['backbone.stage1_conv1_1.bn.running_mean']
        arrays[tag + '_stem_running_var'] = sd['backbone.stage1_conv1_1.bn.running_var']
        last = 'backbone.stage5_%d.conv%d.bn.running_var' % ((2, 3) if tag == 'r50vd' else (1, 2))
        arrays[tag + '_last_running_var'] = sd[last]
        g = head.yolo_output_convs[0].conv.bias.grad
        arrays[tag + '_out0_bias_grad'] = g
        w = head.detection_blocks[0].layers[1].conv.weight.grad
        arrays[tag + '_blk0_w_gradsum'] = np.array([float(w.double().sum()), float(w.double().abs().sum())])
        print('   train %s:' % tag, {k: round(float(v.detach()), 4) for k, v in losses.items()})
    save('train', **arrays)


def gen_train_c4():
    """BASELINE config C4 at its FULL shape: ppyolo_2x 608x608, bs 8, one training forward+backward of the reference (frozen
    backbone in train mode, DropBlock disabled for determinism) -> the six losses and gradient probes."""
    from config import select_loss
    spec = importlib.util.spec_from_file_location('targets', os.path.join(REPO, 'pytorch-ppyolo_b200', 'ppyolo_b200', 'targets.py'))
    tg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tg)
    torch.set_num_threads(8)
    cfg = PPYOLO_2x_Config()
    size, batch = 608, 8
    iou_loss = select_loss(cfg.iou_loss_type)(**cfg.iou_loss)
    iou_aware = select_loss(cfg.iou_aware_loss_type)(**cfg.iou_aware_loss)
    yolo = select_loss(cfg.yolo_loss_type)(iou_loss=iou_loss, iou_aware_loss=iou_aware, **cfg.yolo_loss)
    head_kw = dict(cfg.head); head_kw['drop_block'] = False
    backbone = select_backbone(cfg.backbone_type)(**cfg.backbone)
    head = select_head(cfg.head_type)(yolo_loss=yolo, is_train=True, nms_cfg=cfg.nms_cfg, **head_kw)
    model = PPYOLO(backbone, head)
    synth.randomize_(model, seed=0)
    model.train()
    backbone.freeze()
    x = synth.images(batch, size, seed=20)
    gt_bbox, gt_class, gt_score = tg.synthetic_ground_truth(batch, seed=30)
    targets = tg.gt2yolo_target(gt_bbox, gt_class, gt_score, h=size, w=size, **cfg.gt2YoloTarget)
    losses = model(x, None, False, torch.from_numpy(gt_bbox), torch.from_numpy(gt_class), torch.from_numpy(gt_score),
                   [torch.from_numpy(t) for t in targets])
    sum(losses.values()).backward()
    arrays = {k: v.detach() for k, v in losses.items()}
    arrays['out0_bias_grad'] = head.yolo_output_convs[0].conv.bias.grad
    arrays['out2_bias_grad'] = head.yolo_output_convs[2].conv.bias.grad
    w = head.detection_blocks[0].layers[1].conv.weight.grad
    arrays['blk0_w_gradabs'] = np.array([float(w.double().abs().sum())])
    arrays['stem_running_mean'] = model.state_dict()['backbone.stage1_conv1_1.bn.running_mean']
    print('   train C4:', {k: round(float(v.detach()), 4) for k, v in losses.items()})
    save('train_c4', **arrays)
    torch.set_num_threads(1)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'train_c4':
        gen_train_c4()
        sys.exit(0)
    gen_train()
    gen_losses()
    gen_nms()
    gen_decode()
    gen_layers()
    gen_net('net_r18vd_128', PPYOLO_r18vd_Config(), 128)
    gen_net('net_r50vd_128', PPYOLO_2x_Config(), 128)

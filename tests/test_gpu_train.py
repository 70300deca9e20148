"""GPU parity of the training step (BASELINE config C4 path at test size): losses, a head gradient and the backbone's
BatchNorm running statistics after one forward+backward against the unmodified reference (tests/golden/train.npz);
fused SGD against torch.optim.SGD; batch-statistic BN kernels against torch."""
import numpy as np
import pytest
import torch

import config as cfgs
from ppyolo_b200 import synth, targets as tg
from tests.helpers import CONFIGS

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def build_train_model(tag):
    from model.ppyolo import PPYOLO
    cfg = CONFIGS[tag]()
    iou_loss = cfgs.select_loss(cfg.iou_loss_type)(**cfg.iou_loss)
    iou_aware = cfgs.select_loss(cfg.iou_aware_loss_type)(**cfg.iou_aware_loss) if cfg.head['iou_aware'] else None
    yolo = cfgs.select_loss(cfg.yolo_loss_type)(iou_loss=iou_loss, iou_aware_loss=iou_aware, **cfg.yolo_loss)
    head_kw = dict(cfg.head)
    head_kw['drop_block'] = False
    backbone = cfgs.select_backbone(cfg.backbone_type)(**cfg.backbone)
    head = cfgs.select_head(cfg.head_type)(yolo_loss=yolo, is_train=True, nms_cfg=cfg.nms_cfg, **head_kw)
    model = PPYOLO(backbone, head)
    synth.randomize_(model, seed=0)
    model.train()
    backbone.freeze()
    return model.to(DEV), cfg


def train_inputs(cfg, size=128, batch=2):
    x = synth.images(batch, size, seed=1).to(DEV)
    gt_bbox, gt_class, gt_score = tg.synthetic_ground_truth(batch, seed=3)
    targets = tg.gt2yolo_target(gt_bbox, gt_class, gt_score, h=size, w=size, **cfg.gt2YoloTarget)
    to = lambda a: torch.from_numpy(a).to(DEV)
    return x, to(gt_bbox), to(gt_class), to(gt_score), [to(t) for t in targets]


@pytest.mark.parametrize('tag', ['r50vd', 'r18vd'])
def test_train_forward_backward_vs_reference(golden, tag):
    z = golden('train')
    model, cfg = build_train_model(tag)
    x, gb, gc, gs, targets = train_inputs(cfg)
    losses = model(x, None, False, gb, gc, gs, targets)
    sum(losses.values()).backward()
    for k, v in losses.items():
        np.testing.assert_allclose(float(v.detach()), float(z['%s_%s' % (tag, k)]), rtol=5e-3, err_msg=k)   # head convs: ATen TF32
    sd = model.state_dict()
    np.testing.assert_allclose(sd['backbone.stage1_conv1_1.bn.running_mean'].cpu().numpy(), z[tag + '_stem_running_mean'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd['backbone.stage1_conv1_1.bn.running_var'].cpu().numpy(), z[tag + '_stem_running_var'], rtol=1e-4, atol=1e-6)
    last = 'backbone.stage5_%d.conv%d.bn.running_var' % ((2, 3) if tag == 'r50vd' else (1, 2))
    np.testing.assert_allclose(sd[last].cpu().numpy(), z[tag + '_last_running_var'], rtol=2e-3, atol=1e-6)
    g = model.head.yolo_output_convs[0].conv.bias.grad.cpu().numpy()
    want = z[tag + '_out0_bias_grad']
    np.testing.assert_allclose(g, want, rtol=0, atol=1e-2 * np.abs(want).max())
    w = model.head.detection_blocks[0].layers[1].conv.weight.grad
    # (the signed sum cancels to ~0.1% of the absolute sum, i.e. below the TF32 noise of ATen's convs: compare magnitudes)
    np.testing.assert_allclose(float(w.double().abs().sum()), z[tag + '_blk0_w_gradsum'][1], rtol=3e-2)


def test_bn_batch_stats_kernels():
    from ppyolo_b200 import ops
    from ppyolo_b200._lib import lib, check, PPY_F32, PPY_BF16
    g = torch.Generator().manual_seed(0)
    for code, tol in ((PPY_F32, 1e-5), (PPY_BF16, 1e-2)):
        x = torch.randn((3, 24, 17, 19), generator=g) * 2 + 0.5
        res = torch.randn((3, 24, 17, 19), generator=g)
        bn = torch.nn.BatchNorm2d(24)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        xr = x if code == PPY_F32 else x.to(torch.bfloat16).float()
        rr = res if code == PPY_F32 else res.to(torch.bfloat16).float()
        want = torch.relu(bn.train()(xr) + rr)
        xh, rh = ops.to_nhwc(x.to(DEV), code), ops.to_nhwc(res.to(DEV), code)
        rm, rv = torch.zeros(24, device=DEV), torch.ones(24, device=DEV)
        with torch.no_grad():
            bn2 = torch.nn.BatchNorm2d(24)
            bn2.load_state_dict({k: v for k, v in bn.state_dict().items()})
        # start from the same running stats as the torch module had BEFORE its forward
        sc, sh = torch.empty(24, device=DEV), torch.empty(24, device=DEV)
        ws = torch.empty(48, dtype=torch.float64, device=DEV)
        rm0 = torch.zeros(24, device=DEV); rv0 = torch.ones(24, device=DEV)
        gam, bet = bn.weight.detach().to(DEV), bn.bias.detach().to(DEV)
        n, h, w, ld = xh.shape
        check(lib.ppy_bn_batch_stats(ops.ptr(xh), ld, n * h * w, 24, code, ops.ptr(gam), ops.ptr(bet), 1e-5, 0.1, ops.ptr(rm0), ops.ptr(rv0),
                                     ops.ptr(sc), ops.ptr(sh), ops.ptr(ws), ops.stream_ptr()), 'bn_stats')
        y = torch.empty_like(xh)
        check(lib.ppy_scale_shift_act(ops.ptr(xh), ld, ops.ptr(y), ld, n * h * w, 24, code, ops.ptr(sc), ops.ptr(sh), ops.ptr(rh), ld, 1,
                                      ops.stream_ptr()), 'scale_shift_act')
        got = ops.from_nhwc(y, 24).cpu()
        np.testing.assert_allclose(got.numpy(), want.detach().numpy(), rtol=0, atol=tol * float(want.abs().max()))
        ref_bn = torch.nn.BatchNorm2d(24).train()
        ref_bn(xr)
        np.testing.assert_allclose(rm0.cpu().numpy(), ref_bn.running_mean.numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(rv0.cpu().numpy(), ref_bn.running_var.numpy(), rtol=1e-4, atol=1e-6)


def test_trainer_step_matches_torch_sgd():
    """Two Trainer.step()s (fused SGD kernel on the flat gradient bucket) vs torch.optim.SGD fed with the same gradients."""
    from ppyolo_b200.trainer import Trainer, calc_lr
    model, cfg = build_train_model('r18vd')
    x, gb, gc, gs, targets = train_inputs(cfg)
    trainer = Trainer(model, cfg)
    assert len(trainer.groups) == 19                       # frozen-backbone param groups of r18vd (reference: 19)
    shadow = [p.detach().clone().requires_grad_(True) for p in trainer.params]
    opt = torch.optim.SGD([{'params': [s], 'lr': g['lr'], 'weight_decay': g['weight_decay']} for s, g in zip(shadow, trainer.groups)],
                          lr=trainer.base_lr, momentum=trainer.momentum)
    for it in range(2):
        losses = trainer.step(x, gb, gc, gs, targets)
        assert all(torch.isfinite(v) for v in losses.values())
        lr = calc_lr(it, cfg)
        for s, g, pg, i in zip(shadow, trainer.groups, opt.param_groups, range(len(shadow))):
            pg['lr'] = lr * g['base_lr'] / trainer.base_lr
            s.grad = trainer.bucket.view(i).reshape(s.shape).clone()
        opt.step()
        if it == 0:      # iteration 0 has lr 0 (linear warm-up from 0): parameters must not move
            pass
    for s, p in zip(shadow, trainer.params):
        np.testing.assert_allclose(p.detach().cpu().numpy(), s.detach().cpu().numpy(), rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize('n,c,o,k,hw,extra', [(2, 64, 128, 3, 16, 0), (2, 256, 258, 1, 12, 0), (1, 128, 64, 1, 19, 2), (2, 128, 256, 3, 19, 2),
                                              (3, 512, 96, 1, 7, 0)])
def test_conv2d_kernels_forward_dgrad_wgrad(n, c, o, k, hw, extra):
    """conv_autograd.conv2d_kernels: forward, input gradient and weight gradient on the tcgen05 conv kernel against torch's
    fp32 conv on the same bf16-rounded operands (``extra`` = CoordConv input channels of the weight the conv must skip)."""
    from ppyolo_b200.conv_autograd import conv2d_kernels
    g = torch.Generator().manual_seed(c + o + k)
    rb = lambda t: t.to(torch.bfloat16).float()
    x = rb(torch.randn((n, c, hw, hw), generator=g)).to(DEV).requires_grad_(True)
    w = torch.randn((o, c + extra, k, k), generator=g) * (1.0 / (c * k * k) ** 0.5)
    w[:, :c] = rb(w[:, :c])
    w = w.to(DEV).requires_grad_(True)
    b = (torch.randn(o, generator=g) * 0.1).to(DEV).requires_grad_(True)
    dy = rb(torch.randn((n, o, hw, hw), generator=g)).to(DEV)
    pad = (k - 1) // 2
    y = conv2d_kernels(x, w, b, padding=pad, c_main=c, out_f32=True)
    y.backward(dy)
    got = (y.detach().float(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    yr = torch.nn.functional.conv2d(x, w[:, :c], b, 1, pad)
    yr.backward(dy)
    scale = lambda t: float(t.abs().max())
    np.testing.assert_allclose(got[0].cpu().numpy(), yr.detach().cpu().numpy(), rtol=0, atol=2e-4 * scale(yr))
    np.testing.assert_allclose(got[1].cpu().numpy(), x.grad.cpu().numpy(), rtol=0, atol=8e-3 * scale(x.grad))       # dx leaves as bf16
    np.testing.assert_allclose(got[2][:, :c].cpu().numpy(), w.grad[:, :c].cpu().numpy(), rtol=0, atol=2e-4 * scale(w.grad))
    assert float(got[2][:, c:].abs().max()) == 0.0 if extra else True
    np.testing.assert_allclose(got[3].cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-4, atol=1e-4 * scale(b.grad))


def test_head_kernels_impl_matches_aten():
    """One training forward+backward of ppyolo_2x at 128x128 with the head's convs on the tcgen05 kernel ('kernels': bf16
    operands) against the ATen head on the same fp32 backbone features: losses and gradient norms agree to bf16 noise."""
    results = {}
    for impl in ('aten', 'kernels'):
        model, cfg = build_train_model('r50vd')
        model.train_head_impl = impl
        x, gb, gc, gs, targets = train_inputs(cfg)
        losses = model(x, None, False, gb, gc, gs, targets)
        sum(losses.values()).backward()
        grads = {n: p.grad.detach().float().clone() for n, p in model.head.named_parameters() if p.grad is not None}
        results[impl] = ({k: float(v.detach()) for k, v in losses.items()}, grads)
    la, ga = results['aten']
    lk, gk = results['kernels']
    assert set(ga) == set(gk)
    for k in la:
        np.testing.assert_allclose(lk[k], la[k], rtol=3e-2, err_msg=k)
    # gradients: same direction and size.  bf16 GEMM operands through eleven batch-stat BN layers on this random net put the
    # deepest layers at a cosine of ~0.91 against ATen/TF32 (torch's own bf16 convs in the same structure: 0.85-0.88; exact
    # fp32 convs: > 0.99, see autograd_head.ACT_FP32), the layers next to the outputs at > 0.99
    worst_cos, worst_ratio = 1.0, 0.0
    for n in ga:
        a, k = ga[n].flatten(), gk[n].flatten()
        if float(a.norm()) < 1e-3:
            continue
        cos = float((a * k).sum() / (a.norm() * k.norm()))
        worst_cos = min(worst_cos, cos)
        worst_ratio = max(worst_ratio, abs(float(k.norm() / a.norm()) - 1.0))
        if n.startswith('yolo_output_convs'):
            assert cos > 0.99, (n, cos)
    print('head kernels vs aten: worst gradient cosine %.4f, worst norm deviation %.4f' % (worst_cos, worst_ratio))
    assert worst_cos > 0.80 and worst_ratio < 0.15


def test_graphed_training_step_matches_eager():
    """model.train_graph: backbone engine + head forward/losses/backward replayed as CUDA graphs give the same losses and
    gradients as the eager step (DropBlock off: identical arithmetic), twice in a row, without advancing the BatchNorm running
    statistics during capture."""
    out = {}
    for graph in (False, True):
        model, cfg = build_train_model('r50vd')
        model.train_precision = 'bf16'
        model.train_graph = graph
        x, gb, gc, gs, targets = train_inputs(cfg)
        for rep in range(2):
            for p in model.parameters():
                p.grad = None
            losses = model(x, None, False, gb, gc, gs, targets)
            sum(losses.values()).backward()
        out[graph] = ({k: float(v.detach()) for k, v in losses.items()},
                      {n: p.grad.detach().clone() for n, p in model.head.named_parameters() if p.grad is not None},
                      {k: v.detach().clone() for k, v in model.state_dict().items() if 'running_' in k})
    le, ge, se = out[False]
    lg, gg, sg = out[True]
    for k in le:
        np.testing.assert_allclose(lg[k], le[k], rtol=2e-3, err_msg=k)          # wgrad partial sums: atomics order varies
    for n in ge:
        a, b = ge[n].flatten().float(), gg[n].flatten().float()
        assert float((a - b).norm()) <= 2e-2 * float(a.norm()) + 1e-6, n
    for k in se:
        np.testing.assert_allclose(sg[k].cpu().numpy(), se[k].cpu().numpy(), rtol=2e-3, atol=1e-5, err_msg=k)


def test_trainer_fused_ema_checkpoint_and_eval_refresh(tmp_path):
    """Trainer with the EMA fused into the optimizer kernel: (1) the shadow equals a stand-alone ExponentialMovingAverage
    (model/EMA.py semantics) updated after every step; (2) save_checkpoint / load_checkpoint restore weights, momentum, EMA and the
    iteration; (3) after training, eager module-level eval sees the NEW weights (packed-weight cache refreshed: ADVICE r1) and
    agrees with a freshly built engine."""
    from model.EMA import ExponentialMovingAverage
    from ppyolo_b200.trainer import Trainer
    model, cfg = build_train_model('r18vd')
    x, gb, gc, gs, targets = train_inputs(cfg)
    trainer = Trainer(model, cfg, ema=True)
    ref_ema = ExponentialMovingAverage(model, cfg.ema_decay)
    ref_ema.register()
    trainer.iter_id = 2000                                  # past lr = 0 of the first warm-up step
    for _ in range(3):
        trainer.step(x, gb, gc, gs, targets)
        ref_ema.update()
    assert torch.equal(trainer.ema._shadow_flat, ref_ema._shadow_flat)
    assert trainer.ema._update_step == 3
    path = trainer.save_checkpoint(str(tmp_path))
    w_before = [p.detach().clone() for p in trainer.params]
    mom_before = trainer.momentum_flat.clone()
    trainer.step(x, gb, gc, gs, targets)                    # move on, then restore
    assert not torch.equal(trainer.momentum_flat, mom_before)
    it = trainer.load_checkpoint(path)
    assert it == 2003 and torch.equal(trainer.momentum_flat, mom_before)
    for a, b in zip(w_before, trainer.params):
        assert torch.equal(a, b.detach())
    assert torch.equal(trainer.ema._shadow_flat, ref_ema._shadow_flat)
    # eval after training: module-level path (packs weights lazily) vs a fresh engine, both must see the trained weights
    model.eval()
    model.head.set_dropblock(True)
    model.precision = 'fp32'
    im = synth.im_sizes(x.shape[0]).to(DEV)
    model.use_engine = True
    model(x, im)
    outs_engine = [o.clone() for o in model.engine(x.shape[0], x.shape[2], x.shape[3]).head_outputs_nchw()]
    model.use_engine = False
    outs_eager = model.head._get_outputs(model.backbone(x))
    for a, b in zip(outs_engine, outs_eager):
        np.testing.assert_allclose(a.cpu().numpy(), b.detach().cpu().numpy(), rtol=0, atol=1e-4 * float(a.abs().max()))

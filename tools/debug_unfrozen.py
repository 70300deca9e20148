"""Debug aid: unfrozen-backbone training step in three arithmetic variants against the reference golden.
(a) fp32 ATen convs + oracle dcnv2 autograd (structure check), (b) kernels path."""
import os, sys
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
from tests.test_gpu_unfrozen import _unfrozen_model, train_inputs
from ppyolo_b200 import autograd_head, autograd_backbone
from oracle import ppyolo_ref as ref
import torch.nn.functional as F

z = np.load(os.path.join(REPO, 'tests/golden/train_unfrozen.npz'))
orig_conv_unit = autograd_head.conv_unit

def conv_unit_fp32(u, x, impl='aten', coord=False):
    if not isinstance(u.conv, torch.nn.Conv2d):
        d = u.conv
        y = ref.dcnv2(x, d.conv_offset.weight, d.conv_offset.bias, d.dcn_weight, d.stride, d.padding)
        bn = u.bn
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, 0.1, bn.eps)
        return F.relu(y) if u.act_name == 'relu' else y
    return orig_conv_unit(u, x, 'aten', coord)

def conv_unit_atendcn(u, x, impl='aten', coord=False):
    if not isinstance(u.conv, torch.nn.Conv2d):
        return orig_conv_unit(u, x, 'kernels', coord).float()
    return orig_conv_unit(u, x, 'aten', coord)

def conv_unit_torchbf16(u, x, impl='kernels', coord=False):
    """same rounding points as the kernels path, computed by torch's own bf16 convs / the oracle's DCN"""
    rb = lambda t: t.to(torch.bfloat16)
    if not isinstance(u.conv, torch.nn.Conv2d):
        d = u.conv
        y = ref.dcnv2(x.float(), rb(d.conv_offset.weight).float(), d.conv_offset.bias, rb(d.dcn_weight).float(), d.stride, d.padding).to(torch.bfloat16)
    else:
        w = u.conv.weight
        xx = autograd_head.coord_concat(x.float()).to(torch.bfloat16) if coord else x
        y = F.conv2d(xx.float(), rb(w).float(), u.conv.bias, u.stride, u.padding)
        if u.bn is not None:
            y = y.to(torch.bfloat16)
    if u.bn is not None:
        bn = u.bn
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, 0.1, bn.eps)
    if u.act_name == 'relu': y = F.relu(y)
    elif u.act_name == 'leaky': y = F.leaky_relu(y, 0.1)
    return y

BN_EVAL = False
SIZE = 128
def run(tag, fa, variant):
    model, cfg = _unfrozen_model(tag, fa)
    if BN_EVAL:
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
    x, gb, gc, gs, targets = train_inputs(cfg, size=SIZE)
    if variant in ('fp32', 'atendcn', 'torchbf16'):
        fn = {'fp32': conv_unit_fp32, 'atendcn': conv_unit_atendcn, 'torchbf16': conv_unit_torchbf16}[variant]
        autograd_head.conv_unit = fn
        autograd_backbone.conv_unit = fn
        impl = 'kernels' if variant == 'torchbf16' else 'aten'
        feats = autograd_backbone.backbone_features(model.backbone, x, impl)
        model.head.train_impl = impl
        losses = model.head.get_loss_autograd(feats, gb, gc, gs, targets)
        autograd_head.conv_unit = orig_conv_unit
        autograd_backbone.conv_unit = orig_conv_unit
    else:
        losses = model(x, None, False, gb, gc, gs, targets)
    sum(losses.values()).backward()
    print(tag, variant, {k: round(float(v.detach()), 3) for k, v in losses.items()})
    params = dict(model.named_parameters())
    out = {}
    for key in (z.files if (SIZE == 128 and not BN_EVAL) else []):
        if key.startswith(tag + '_grad:'):
            name = key.split(':', 1)[1]
            want = z[key].astype(np.float64)
            got = params[name].grad.detach().float().flatten()[:want.size].cpu().numpy().astype(np.float64)
            cos = float((got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
            out[name.replace('backbone.', '')] = round(cos, 4)
    print('   cos:', out)
    return {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}


def cmp(a, b, label):
    cs = {}
    for n in a:
        if float(a[n].norm()) > 1e-6:
            cs[n] = float((a[n] * b[n]).sum() / (a[n].norm() * b[n].norm()))
    bb = [v for n, v in cs.items() if n.startswith('backbone')]
    hd = [v for n, v in cs.items() if n.startswith('head')]
    print('%s: backbone cos min %.4f median %.4f | head cos min %.4f median %.4f' % (label, min(bb), float(np.median(bb)), min(hd), float(np.median(hd))))
    worst = sorted(cs.items(), key=lambda kv: kv[1])[:5]
    print('    worst:', [(n, round(v, 4)) for n, v in worst])

for bn_eval, size in ((True, 128), (False, 256), (True, 256)):
    BN_EVAL, SIZE = bn_eval, size
    print('==== BN eval mode %s, size %d' % (bn_eval, size))
    a = run('r50vd', 3, 'fp32')
    b = run('r50vd', 3, 'kernels')
    c = run('r50vd', 3, 'atendcn')
    cmp(a, b, 'kernels vs fp32')
    cmp(a, c, 'aten+dcn-kernels vs fp32')

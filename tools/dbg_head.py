import os, sys
sys.path.insert(0, '/root/repo/pytorch-ppyolo_b200'); sys.path.insert(0, '/root/repo')
import torch, numpy as np
from tests.test_gpu_train import build_train_model, train_inputs
from ppyolo_b200 import autograd_head
orig_ho = autograd_head.head_outputs
def run(impl, bnk):
    autograd_head.BN_KERNELS = bnk
    kept = {}
    def ho(head, feats, impl_='aten'):
        outs = orig_ho(head, feats, impl_)
        for o in outs: o.retain_grad()
        kept['outs'] = outs
        return outs
    autograd_head.head_outputs = ho
    model, cfg = build_train_model('r50vd')
    model.train_head_impl = impl
    x, gb, gc, gs, targets = train_inputs(cfg)
    losses = model(x, None, False, gb, gc, gs, targets)
    sum(losses.values()).backward()
    autograd_head.head_outputs = orig_ho
    return [o.detach().float() for o in kept['outs']], [o.grad.detach().float() for o in kept['outs']]
oa, ga = run('aten', False)
for bnk in (False, True):
    ok, gk = run('kernels', bnk)
    for i in range(3):
        cos = float((ga[i] * gk[i]).sum() / (ga[i].norm() * gk[i].norm()))
        d = (gk[i] - ga[i]).abs()
        idx = int(d.argmax())
        ch = (idx // (ga[i].shape[2] * ga[i].shape[3])) % ga[i].shape[1]
        print('fusedBN' if bnk else 'atenBN ', 'out%d' % i, 'dY cos %.5f' % cos, 'max |dY diff| %.3e at channel %d (|dY| max %.3e)' % (float(d.max()), ch, float(ga[i].abs().max())),
              'fwd err at that element %.3e' % float((ok[i] - oa[i]).flatten()[idx].abs()), 'fwd value %.3f' % float(oa[i].flatten()[idx]))

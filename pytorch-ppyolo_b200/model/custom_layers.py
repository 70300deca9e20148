"""Op library of the PP-YOLO hot path, B200-native.

Mirrors the public surface of the reference's ``model/custom_layers.py``
(``Conv2dUnit`` :65-253, ``CoordConv`` :256-272, ``SPP`` :275-290,
``DropBlock`` :293-342, ``DCNv2`` :486-677, ``get_norm`` :22-34) so that
reference checkpoints load with ``strict=True`` and the reference's entry
scripts keep working.  The torch modules held inside (``conv``, ``bn``) are
*parameter containers only*: every forward goes through the hand-written
sm_100a kernels behind the C ABI (``include/ppyolo_b200.h``) and raises when
the extension or a GPU is missing -- there is no torch/CPU fallback.
"""
import math

import torch

from ppyolo_b200 import ops

ACT_CODES = {None: 0, 'relu': 1, 'leaky': 2, 'mish': 3}


def get_norm(norm_type):
    """(bn, gn, af) flags for a norm name; 'sync_bn' is plain BN as in the reference (:22-34)."""
    table = {'bn': (1, 0, 0), 'sync_bn': (1, 0, 0), 'gn': (0, 1, 0), 'affine_channel': (0, 0, 1)}
    return table.get(norm_type, (0, 0, 0))


class AffineChannel(torch.nn.Module):
    """Per-channel scale+shift (reference :46-62). Parameter container; folded like an eval BN."""

    def __init__(self, num_features):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(num_features))
        self.bias = torch.nn.Parameter(torch.randn(num_features))


class DCNv2(torch.nn.Module):
    """Modulated deformable 3x3 conv (reference :486-677).

    Parameters keep the reference names: ``conv_offset.{weight,bias}`` (27 ch,
    zero-init) and ``dcn_weight`` [Cout,Cin,kH,kW]; no bias by default.
    Offsets are (dy,dx) interleaved per tap, mask = sigmoid of the last kH*kW
    channels; samples outside the image contribute 0.
    """

    def __init__(self, input_dim, filters, filter_size, stride=1, padding=0, bias_attr=False,
                 distribution='normal', gain=1):
        super().__init__()
        if distribution not in ('uniform', 'normal'):
            raise AssertionError(distribution)
        self.input_dim, self.filters, self.filter_size = input_dim, filters, filter_size
        self.stride, self.padding = stride, padding
        self.conv_offset = torch.nn.Conv2d(input_dim, filter_size * filter_size * 3, kernel_size=filter_size,
                                           stride=stride, padding=padding, bias=True)
        torch.nn.init.zeros_(self.conv_offset.weight)
        torch.nn.init.zeros_(self.conv_offset.bias)
        self.sigmoid = torch.nn.Sigmoid()
        self.dcn_weight = torch.nn.Parameter(torch.empty(filters, input_dim, filter_size, filter_size))
        self.dcn_bias = torch.nn.Parameter(torch.zeros(filters)) if bias_attr else None
        init = torch.nn.init.xavier_uniform_ if distribution == 'uniform' else torch.nn.init.xavier_normal_
        init(self.dcn_weight, gain=gain)

    def forward(self, x):
        return ops.dcnv2(x, self.conv_offset.weight, self.conv_offset.bias, self.dcn_weight, self.dcn_bias,
                         stride=self.stride, padding=self.padding)


class Conv2dUnit(torch.nn.Module):
    """conv (or DCNv2) -> norm -> activation, executed as ONE fused kernel.

    Same constructor as the reference (:65-139).  ``forward`` folds the norm
    into a per-channel scale/shift applied in the conv epilogue together with
    the activation.  BN uses running statistics in eval mode; in train mode the
    batch statistics are computed by a reduction kernel first (the reference
    leaves frozen-backbone BNs in train mode).
    """

    def __init__(self, input_dim, filters, filter_size, stride=1, bias_attr=False, bn=0, gn=0, af=0, groups=32,
                 act=None, freeze_norm=False, is_test=False, norm_decay=0., lr=1., bias_lr=None,
                 weight_init=None, bias_init=None, use_dcn=False, name=''):
        super().__init__()
        if act not in ACT_CODES:
            raise NotImplementedError("Activation '{}' is not implemented.".format(act))
        self.groups, self.filters, self.filter_size, self.stride = groups, filters, filter_size, stride
        self.padding = (filter_size - 1) // 2
        self.freeze_norm, self.is_test, self.norm_decay = freeze_norm, is_test, norm_decay
        self.use_dcn, self.name, self.lr = use_dcn, name, lr
        self.act_name = act
        if use_dcn:
            self.conv = DCNv2(input_dim, filters, filter_size=filter_size, stride=stride,
                              padding=self.padding, bias_attr=False)
        else:
            if bias_attr:
                self.blr = bias_lr if bias_lr else lr
            self.conv = torch.nn.Conv2d(input_dim, filters, kernel_size=filter_size, stride=stride,
                                        padding=self.padding, bias=bias_attr)
        self.bn = torch.nn.BatchNorm2d(filters) if bn else None
        self.gn = torch.nn.GroupNorm(num_groups=groups, num_channels=filters) if gn else None
        self.af = AffineChannel(filters) if af else None
        # `act` attribute kept for surface compatibility (reference stores a module or None)
        self.act = None if act is None else act

    # ---- parameter bookkeeping (reference :142-241) -------------------------------------------
    def _conv_params(self):
        if isinstance(self.conv, DCNv2):
            ps = [self.conv.conv_offset.weight, self.conv.conv_offset.bias, self.conv.dcn_weight]
            if self.conv.dcn_bias is not None:
                ps.append(self.conv.dcn_bias)
            return ps
        ps = [self.conv.weight]
        if self.conv.bias is not None:
            ps.append(self.conv.bias)
        return ps

    def _norm_params(self):
        out = []
        for m in (self.bn, self.gn, self.af):
            if m is not None:
                out += [m.weight, m.bias]
        return out

    def freeze(self):
        for p in self._conv_params() + self._norm_params():
            p.requires_grad = False

    def add_param_group(self, param_groups, base_lr, base_wd):
        def push(p, lr_mult, wd):
            if p.requires_grad:
                param_groups.append({'params': [p], 'lr': base_lr * lr_mult, 'base_lr': base_lr * lr_mult,
                                     'weight_decay': wd})
        if isinstance(self.conv, DCNv2):
            # the reference decays all three DCN tensors, offset bias included (:182-200)
            for p in (self.conv.conv_offset.weight, self.conv.conv_offset.bias, self.conv.dcn_weight):
                push(p, self.lr, base_wd)
        elif self.conv.weight.requires_grad:
            push(self.conv.weight, self.lr, base_wd)
            if self.conv.bias is not None:
                push(self.conv.bias, self.blr, 0.0)
        for p in self._norm_params():
            push(p, self.lr, 0.0)

    # ---- execution ---------------------------------------------------------------------------
    def folded_scale_shift(self):
        """Per-output-channel (scale, shift) equivalent to bias + eval-mode norm."""
        w = self.conv.dcn_weight if isinstance(self.conv, DCNv2) else self.conv.weight
        cout = w.shape[0]
        bias = None
        if isinstance(self.conv, DCNv2):
            bias = self.conv.dcn_bias
        elif self.conv.bias is not None:
            bias = self.conv.bias
        scale = torch.ones(cout, dtype=torch.float32, device=w.device)
        shift = torch.zeros(cout, dtype=torch.float32, device=w.device)
        if bias is not None:
            shift = shift + bias.detach().float()
        if self.bn is not None:
            inv = torch.rsqrt(self.bn.running_var.float() + self.bn.eps) * self.bn.weight.detach().float()
            shift = (shift - self.bn.running_mean.float()) * inv + self.bn.bias.detach().float()
            scale = scale * inv
        if self.af is not None:
            shift = shift * self.af.weight.detach().float() + self.af.bias.detach().float()
            scale = scale * self.af.weight.detach().float()
        if self.gn is not None:
            raise NotImplementedError('GroupNorm is not on the PP-YOLO hot path (no config uses it)')
        return scale.contiguous(), shift.contiguous()

    def forward(self, x, residual=None):
        if self.bn is not None and self.bn.training:
            raise NotImplementedError(
                'train-mode BatchNorm (batch statistics) is not built yet; call model.eval() first')
        scale, shift = self.folded_scale_shift()
        act = ACT_CODES[self.act_name]
        if isinstance(self.conv, DCNv2):
            return ops.dcnv2(x, self.conv.conv_offset.weight, self.conv.conv_offset.bias, self.conv.dcn_weight, None,
                             stride=self.stride, padding=self.padding, scale=scale, shift=shift, act=act,
                             residual=residual)
        return ops.conv_bn_act(x, self.conv.weight, scale, shift, stride=self.stride, padding=self.padding, act=act,
                               residual=residual)


class CoordConv(torch.nn.Module):
    """Appends x in [-1,1] along W then y along H as two channels (reference :256-272)."""

    def __init__(self, coord_conv=True):
        super().__init__()
        self.coord_conv = coord_conv

    def forward(self, x):
        return ops.coord_concat(x) if self.coord_conv else x


class SPP(torch.nn.Module):
    """cat([x, maxpool5, maxpool9, maxpool13]) in one kernel (reference :275-290)."""

    def __init__(self, seq='asc'):
        super().__init__()
        if seq not in ('desc', 'asc'):
            raise AssertionError(seq)
        self.seq = seq

    def forward(self, x):
        return ops.spp(x, descending=(self.seq == 'desc'))


class DropBlock(torch.nn.Module):
    """Train-time block dropout (reference :293-342); identity when ``is_test``."""

    def __init__(self, block_size=3, keep_prob=0.9, is_test=False):
        super().__init__()
        self.block_size, self.keep_prob, self.is_test = block_size, keep_prob, is_test

    def gamma(self, h):
        bs = self.block_size
        return (1.0 - self.keep_prob) * h * h / float(bs * bs * (h - bs + 1) ** 2)

    def forward(self, x):
        if self.is_test:
            return x
        return ops.drop_block(x, self.block_size, self.keep_prob)


class Mish(torch.nn.Module):
    """x * tanh(softplus(x)); defined by the reference (:37-43) but unused by both configs."""

    def forward(self, x):
        return ops.activation(x, ACT_CODES['mish'])


def mish_reference_value(v):
    return v * math.tanh(math.log1p(math.exp(v)))

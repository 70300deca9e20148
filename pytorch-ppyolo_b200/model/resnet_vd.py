"""ResNet-vd backbones of PP-YOLO on the B200 kernel set.

Same classes, constructor arguments, attribute names and state_dict keys as
the reference's ``model/resnet_vd.py`` (``ConvBlock`` :15-57,
``IdentityBlock`` :60-87, ``Resnet50Vd`` :89-220, ``BasicBlock`` :224-267,
``Resnet18Vd`` :270-366).  The residual add + ReLU of every block is fused
into the epilogue of the block's last conv kernel instead of running as
separate elementwise passes.
"""
import torch

from ppyolo_b200 import ops
from model.custom_layers import Conv2dUnit, get_norm


class _Block(torch.nn.Module):
    """Shared freeze / param-group plumbing for residual blocks."""

    _units = ()

    def _conv_units(self):
        return [getattr(self, n) for n in self._units if getattr(self, n, None) is not None]

    def freeze(self):
        for u in self._conv_units():
            u.freeze()

    def add_param_group(self, param_groups, base_lr, base_wd):
        for u in self._conv_units():
            u.add_param_group(param_groups, base_lr, base_wd)


class ConvBlock(_Block):
    """Bottleneck with projection shortcut; stride lives in the 3x3 when ``downsample_in3x3``.

    Non-first blocks use the "vd" shortcut: AvgPool2x2 then 1x1 stride 1 (reference :29-33).
    """
    _units = ('conv1', 'conv2', 'conv3', 'conv4')

    def __init__(self, in_c, filters, bn, gn, af, freeze_norm, norm_decay, lr, use_dcn=False, stride=2,
                 downsample_in3x3=True, is_first=False, block_name=''):
        super().__init__()
        f1, f2, f3 = filters
        s1, s2 = (1, stride) if downsample_in3x3 else (stride, 1)
        kw = dict(bn=bn, gn=gn, af=af, freeze_norm=freeze_norm, norm_decay=norm_decay, lr=lr)
        self.is_first = is_first
        self.conv1 = Conv2dUnit(in_c, f1, 1, stride=s1, act='relu', name=block_name + '_branch2a', **kw)
        self.conv2 = Conv2dUnit(f1, f2, 3, stride=s2, act='relu', use_dcn=use_dcn, name=block_name + '_branch2b', **kw)
        self.conv3 = Conv2dUnit(f2, f3, 1, stride=1, act=None, name=block_name + '_branch2c', **kw)
        if is_first:
            self.conv4 = Conv2dUnit(in_c, f3, 1, stride=stride, act=None, name=block_name + '_branch1', **kw)
        else:
            self.avg_pool = torch.nn.AvgPool2d(kernel_size=2, stride=2, padding=0)  # container; kernel is ops.avg_pool2
            self.conv4 = Conv2dUnit(in_c, f3, 1, stride=1, act=None, name=block_name + '_branch1', **kw)
        self.act = torch.nn.ReLU(inplace=True)

    def forward(self, x):
        sc_in = x if self.is_first else ops.avg_pool2(x)
        shortcut = self.conv4(sc_in)
        y = self.conv2(self.conv1(x))
        return ops.conv_unit_residual_relu(self.conv3, y, shortcut)


class IdentityBlock(_Block):
    _units = ('conv1', 'conv2', 'conv3')

    def __init__(self, in_c, filters, bn, gn, af, freeze_norm, norm_decay, lr, use_dcn=False, block_name=''):
        super().__init__()
        f1, f2, f3 = filters
        kw = dict(bn=bn, gn=gn, af=af, freeze_norm=freeze_norm, norm_decay=norm_decay, lr=lr)
        self.conv1 = Conv2dUnit(in_c, f1, 1, stride=1, act='relu', name=block_name + '_branch2a', **kw)
        self.conv2 = Conv2dUnit(f1, f2, 3, stride=1, act='relu', use_dcn=use_dcn, name=block_name + '_branch2b', **kw)
        self.conv3 = Conv2dUnit(f2, f3, 1, stride=1, act=None, name=block_name + '_branch2c', **kw)
        self.act = torch.nn.ReLU(inplace=True)

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return ops.conv_unit_residual_relu(self.conv3, y, x)


class BasicBlock(_Block):
    """Two 3x3 convs; stride in the first; 1x1 projection when strided or first (reference :224-267)."""
    _units = ('conv1', 'conv2', 'conv3')

    def __init__(self, in_c, filters, bn, gn, af, freeze_norm, norm_decay, lr, stride=1, is_first=False,
                 block_name=''):
        super().__init__()
        f1, f2 = filters
        kw = dict(bn=bn, gn=gn, af=af, freeze_norm=freeze_norm, norm_decay=norm_decay, lr=lr)
        self.is_first, self.stride = is_first, stride
        self.conv1 = Conv2dUnit(in_c, f1, 3, stride=stride, act='relu', name=block_name + '_branch2a', **kw)
        self.conv2 = Conv2dUnit(f1, f2, 3, stride=1, act=None, name=block_name + '_branch2b', **kw)
        self.conv3 = None
        if stride == 2 or is_first:
            if not is_first:
                self.avg_pool = torch.nn.AvgPool2d(kernel_size=2, stride=2, padding=0)
                self.conv3 = Conv2dUnit(in_c, f2, 1, stride=1, act=None, name=block_name + '_branch1', **kw)
            else:
                self.conv3 = Conv2dUnit(in_c, f2, 1, stride=stride, act=None, name=block_name + '_branch1', **kw)
        self.act = torch.nn.ReLU(inplace=True)

    def forward(self, x):
        if self.conv3 is not None:
            shortcut = self.conv3(x if self.is_first else ops.avg_pool2(x))
        else:
            shortcut = x
        return ops.conv_unit_residual_relu(self.conv2, self.conv1(x), shortcut)


class _ResNetVd(torch.nn.Module):
    """Deep-stem ResNet-vd trunk: 3x(3x3) stem, 3x3/2 max-pool, four stages; returns the C2..C5 picks."""

    def _init_common(self, norm_type, feature_maps, freeze_at, lr_mult_list):
        assert freeze_at in [0, 1, 2, 3, 4, 5]
        assert len(lr_mult_list) == 4, "lr_mult_list length must be 4 but got {}".format(len(lr_mult_list))
        assert norm_type in ['bn', 'sync_bn', 'gn', 'affine_channel']
        self.norm_type, self.feature_maps = norm_type, feature_maps
        self.lr_mult_list, self.freeze_at = lr_mult_list, freeze_at
        return get_norm(norm_type)

    def _make_stem(self, bn, gn, af, freeze_norm, norm_decay):
        kw = dict(bn=bn, gn=gn, af=af, freeze_norm=freeze_norm, norm_decay=norm_decay, act='relu')
        self.stage1_conv1_1 = Conv2dUnit(3, 32, 3, stride=2, name='conv1_1', **kw)
        self.stage1_conv1_2 = Conv2dUnit(32, 32, 3, stride=1, name='conv1_2', **kw)
        self.stage1_conv1_3 = Conv2dUnit(32, 64, 3, stride=1, name='conv1_3', **kw)
        self.pool = torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1)  # container; kernel is ops.max_pool3s2

    def stage_names(self, stage):
        return ['stage%d_%d' % (stage, i) for i in range(self.depths[stage - 2])]

    def get_block(self, name):
        return getattr(self, name)

    def _stem_units(self):
        return [self.stage1_conv1_1, self.stage1_conv1_2, self.stage1_conv1_3]

    def forward(self, x):
        for u in self._stem_units():
            x = u(x)
        x = ops.max_pool3s2(x)
        outs = []
        for stage in (2, 3, 4, 5):
            for name in self.stage_names(stage):
                x = getattr(self, name)(x)
            if stage in self.feature_maps:
                outs.append(x)
        return outs

    def freeze(self):
        if self.freeze_at >= 1:
            for u in self._stem_units():
                u.freeze()
        for stage in (2, 3, 4, 5):
            if self.freeze_at >= stage:
                for name in self.stage_names(stage):
                    getattr(self, name).freeze()

    def add_param_group(self, param_groups, base_lr, base_wd):
        for u in self._stem_units():
            u.add_param_group(param_groups, base_lr, base_wd)
        for stage in (2, 3, 4, 5):
            for name in self.stage_names(stage):
                getattr(self, name).add_param_group(param_groups, base_lr, base_wd)


class Resnet50Vd(_ResNetVd):
    depths = (3, 4, 6, 3)

    def __init__(self, norm_type='bn', feature_maps=[3, 4, 5], dcn_v2_stages=[5], downsample_in3x3=True, freeze_at=0,
                 freeze_norm=False, norm_decay=0., lr_mult_list=[1., 1., 1., 1.]):
        super().__init__()
        bn, gn, af = self._init_common(norm_type, feature_maps, freeze_at, lr_mult_list)
        self._make_stem(bn, gn, af, freeze_norm, norm_decay)
        in_c = 64
        for stage, depth in zip((2, 3, 4, 5), self.depths):
            width = 64 * 2 ** (stage - 2)
            filters = [width, width, width * 4]
            lr = lr_mult_list[stage - 2]
            use_dcn = stage in dcn_v2_stages and stage > 2
            res = 'res%d' % stage
            first = ConvBlock(in_c, filters, bn, gn, af, freeze_norm, norm_decay, lr, use_dcn=use_dcn,
                              stride=1 if stage == 2 else 2, downsample_in3x3=downsample_in3x3,
                              is_first=(stage == 2), block_name=res + 'a')
            setattr(self, 'stage%d_0' % stage, first)
            for i in range(1, depth):
                blk = IdentityBlock(width * 4, filters, bn, gn, af, freeze_norm, norm_decay, lr, use_dcn=use_dcn,
                                    block_name=res + chr(ord('a') + i))
                setattr(self, 'stage%d_%d' % (stage, i), blk)
            in_c = width * 4


class Resnet18Vd(_ResNetVd):
    depths = (2, 2, 2, 2)

    def __init__(self, norm_type='bn', feature_maps=[4, 5], dcn_v2_stages=[], freeze_at=0, freeze_norm=False,
                 norm_decay=0., lr_mult_list=[1., 1., 1., 1.]):
        super().__init__()
        bn, gn, af = self._init_common(norm_type, feature_maps, freeze_at, lr_mult_list)
        self._make_stem(bn, gn, af, freeze_norm, norm_decay)
        in_c = 64
        for stage in (2, 3, 4, 5):
            width = 64 * 2 ** (stage - 2)
            lr = lr_mult_list[stage - 2]
            res = 'res%d' % stage
            setattr(self, 'stage%d_0' % stage,
                    BasicBlock(in_c, [width, width], bn, gn, af, freeze_norm, norm_decay, lr,
                               stride=1 if stage == 2 else 2, is_first=(stage == 2), block_name=res + 'a'))
            setattr(self, 'stage%d_1' % stage,
                    BasicBlock(width, [width, width], bn, gn, af, freeze_norm, norm_decay, lr, stride=1,
                               block_name=res + 'b'))
            in_c = width

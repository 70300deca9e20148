"""Name -> class selectors (reference config/get_model.py:15-42)."""
import torch

from model.head import YOLOv3Head
from model.iou_losses import IouLoss, IouAwareLoss
from model.losses import YOLOv3Loss
from model.resnet_vd import Resnet50Vd, Resnet18Vd

__all__ = ['select_backbone', 'select_head', 'select_loss', 'select_optimizer']

_BACKBONES = {'Resnet50Vd': Resnet50Vd, 'Resnet18Vd': Resnet18Vd}
_HEADS = {'YOLOv3Head': YOLOv3Head}
_OPTIMIZERS = {'Momentum': torch.optim.SGD, 'SGD': torch.optim.SGD, 'Adam': torch.optim.Adam}


def select_backbone(name):
    return _BACKBONES.get(name)


def select_head(name):
    return _HEADS.get(name)


_LOSSES = {'YOLOv3Loss': YOLOv3Loss, 'IouLoss': IouLoss, 'IouAwareLoss': IouAwareLoss}


def select_loss(name):
    return _LOSSES.get(name)


def select_optimizer(name):
    return _OPTIMIZERS.get(name)

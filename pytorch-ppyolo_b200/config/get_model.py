"""Name -> class selectors (reference config/get_model.py:15-42)."""
import torch

from model.head import YOLOv3Head
from model.resnet_vd import Resnet50Vd, Resnet18Vd

__all__ = ['select_backbone', 'select_head', 'select_loss', 'select_optimizer']

_BACKBONES = {'Resnet50Vd': Resnet50Vd, 'Resnet18Vd': Resnet18Vd}
_HEADS = {'YOLOv3Head': YOLOv3Head}
_OPTIMIZERS = {'Momentum': torch.optim.SGD, 'SGD': torch.optim.SGD, 'Adam': torch.optim.Adam}


def select_backbone(name):
    return _BACKBONES.get(name)


def select_head(name):
    return _HEADS.get(name)


def select_loss(name):
    """Training losses are a later row of the scope table; selecting one fails loudly."""
    if name in ('YOLOv3Loss', 'IouLoss', 'IouAwareLoss'):
        def _missing(*args, **kwargs):
            raise NotImplementedError('{} (training path) is not built yet'.format(name))
        return _missing
    return None


def select_optimizer(name):
    return _OPTIMIZERS.get(name)

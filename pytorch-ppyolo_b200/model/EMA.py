"""ExponentialMovingAverage with the reference's interface (model/EMA.py:16-57: register / update / apply / restore), kept
entirely on the device.

The reference copies every trainable parameter to the host and back each step (model/EMA.py:35-41, README.md:67).  Here the
shadow lives in ONE flat fp32 device buffer and ``update()`` is a single kernel launch over all tensors
(``ppy_ema_update``) with numpy's float32 operation order, so the shadow values are bit-identical to the reference's.
``apply()`` / ``restore()`` swap ``param.data`` like the reference (device tensors instead of host arrays)."""
import ctypes

import numpy as np
import torch


class ExponentialMovingAverage(object):
    def __init__(self, model, decay, thres_steps=True):
        self._model = model
        self._decay = decay
        self._thres_steps = thres_steps
        self._names, self._params = [], []
        self._shadow_flat = None
        self._backup = {}
        self._update_step = 0

    # ------------------------------------------------------------------ helpers
    def _trainable(self):
        return [(n, p) for n, p in self._model.named_parameters() if p.requires_grad is True]

    @property
    def _shadow(self):
        """name -> shadow tensor view (the reference's ``_shadow`` dict of arrays)."""
        return {n: self._shadow_flat[self._offsets[i]:self._offsets[i + 1]].view(p.shape)
                for i, (n, p) in enumerate(zip(self._names, self._params))}

    # ------------------------------------------------------------------ reference interface
    def register(self):
        self._update_step = 0
        pairs = self._trainable()
        if not pairs:
            raise ValueError('ExponentialMovingAverage.register: the model has no trainable parameter')
        self._names, self._params = [n for n, _ in pairs], [p for _, p in pairs]
        dev = self._params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('ppyolo_b200: ExponentialMovingAverage keeps its shadow on the GPU -- move the model to CUDA first')
        if any(p.dtype != torch.float32 or not p.is_contiguous() for p in self._params):
            raise ValueError('ExponentialMovingAverage expects contiguous fp32 parameters')
        self._offsets = [0]
        for p in self._params:
            self._offsets.append(self._offsets[-1] + p.numel())
        self._shadow_flat = torch.cat([p.detach().reshape(-1) for p in self._params]).clone()
        self._offsets_dev = torch.tensor(self._offsets, dtype=torch.int64, device=dev)
        self._table = None

    def _pointer_table(self):
        ptrs = [p.data_ptr() for p in self._params]
        if self._table is None or self._table[0] != ptrs:          # apply()/restore() rebind param.data
            self._table = (ptrs, torch.tensor(ptrs, dtype=torch.int64, device=self._shadow_flat.device))
        return self._table[1]

    def update(self):
        from ppyolo_b200._lib import lib, check
        from ppyolo_b200 import ops
        if self._shadow_flat is None:
            raise RuntimeError('call register() first')
        step = self._update_step
        decay = min(self._decay, (1 + step) / (10 + step)) if self._thres_steps else self._decay
        table = self._pointer_table()
        check(lib.ppy_ema_update(ctypes.c_void_p(self._shadow_flat.data_ptr()), ctypes.c_void_p(table.data_ptr()),
                                 ctypes.c_void_p(self._offsets_dev.data_ptr()), len(self._params),
                                 float(np.float32(decay)), float(np.float32(1 - decay)), ops.stream_ptr()), 'ema_update')
        self._update_step += 1
        return decay

    def apply(self):
        shadow = self._shadow
        for n, p in zip(self._names, self._params):
            self._backup[n] = p.data
            p.data = shadow[n].clone()
        self._invalidate()

    def restore(self):
        for n, p in zip(self._names, self._params):
            assert n in self._backup
            p.data = self._backup[n]
        self._backup = {}
        self._invalidate()

    def _invalidate(self):
        inv = getattr(self._model, 'invalidate_engines', None)     # compiled inference plans hold folded weights
        if inv is not None:
            inv()

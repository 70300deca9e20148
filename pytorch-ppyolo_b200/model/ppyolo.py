"""Top-level detector module (reference ``model/ppyolo.py`` :13-29), same forward signature.

In eval mode on a CUDA device the whole network runs through ``ppyolo_b200.engine.InferenceEngine``:
a static plan of fused sm_100a kernels over pre-allocated NHWC buffers (built lazily per input shape,
replayed as a CUDA graph).  ``precision`` selects the arithmetic of the conv kernels:
``'f16x2'`` (default: tcgen05 tensor cores on fp16 hi/lo pairs, three MMAs per K block, fp32 accumulate -- the accuracy of
the reference's fp32 arithmetic), ``'bf16'`` (tcgen05, bf16 operands; fastest, ~1e-2 of scale drift) or ``'fp32'`` (SIMT
fp32; the slow exact path).

Compiled plans snapshot packed conv weights and folded BN parameters.  They are dropped by ``load_state_dict``, by the
EMA's ``apply``/``restore``, by ``Trainer.step`` and -- when any parameter or buffer changed (version counters / storage
pointers) -- by ``model.eval()`` / ``model.train()``; after any OTHER in-place weight update call
``model.invalidate_engines()`` before the next ``model(x, im_size)``.
"""
import torch


class PPYOLO(torch.nn.Module):
    def __init__(self, backbone, head):
        super().__init__()
        self.backbone = backbone
        self.head = head
        self.precision = 'f16x2'
        self.train_precision = 'fp32'     # arithmetic of the frozen-backbone forward inside a training step
        self.train_head_impl = None       # 'kernels' | 'aten' | None = kernels with a bf16 backbone, ATen (TF32) with fp32
        self.train_graph = False          # capture head forward + losses + backward as CUDA graphs (static shapes; see forward_train)
        self._graphed_heads = {}
        self.dcn_impl = None          # None = engine default; 'fused' | 'gather_gemm'
        self.postprocess_impl = None  # None = engine default ('dense'); 'sparse' | 'dense' (see engine.py)
        self.f16x2_act_scale = 8.0    # power of two the 'f16x2' engine stores its activations multiplied by (see engine.py)
        self._weights_seen = None
        self.normalize = None         # dict(mean, std, is_scale) of the uint8 input path; None = the configs' ImageNet values
        self.use_engine = True
        self._engines = {}
        self._mode_mods = {}
        self._mish = None             # cached: does any Conv2dUnit use Mish (module structure is static)

    def engine(self, batch, height, width, input_u8=False):
        from ppyolo_b200.engine import InferenceEngine
        key = (batch, height, width, self.precision, self.dcn_impl, self.postprocess_impl, bool(input_u8))
        eng = self._engines.get(key)
        if eng is None:
            eng = InferenceEngine(self, batch, height, width, precision=self.precision, dcn_impl=self.dcn_impl, input_u8=input_u8)
            self._engines[key] = eng
            self._weights_seen = self._weights_fingerprint()
        return eng

    def _has_mish(self):
        """Mish units (defined by the reference, used by neither config) run module by module: the engine's fused epilogues
        carry none / relu / leaky only."""
        if self._mish is None:
            from model.custom_layers import Conv2dUnit
            self._mish = any(isinstance(m, Conv2dUnit) and m.act_name == 'mish' for m in self.modules())
        return self._mish

    def invalidate_engines(self):
        """Drop compiled plans and packed-weight caches (call after mutating weights in place)."""
        from ppyolo_b200 import ops
        self._engines = {}
        ops._PACK_CACHE.clear()

    def _weights_fingerprint(self):
        acc = 0
        for t in list(self.parameters()) + list(self.buffers()):
            acc = (acc * 1000003 + t._version * 31 + t.data_ptr()) & 0xFFFFFFFFFFFFFFFF
        return acc

    def train(self, mode=True):
        """Mode switches are where the reference's loops go from optimising to evaluating (train.py:484-489): drop the
        compiled plans if any parameter or buffer changed since they were built."""
        fp = self._weights_fingerprint()
        if self._weights_seen is not None and fp != self._weights_seen and self._engines:
            self.invalidate_engines()
        self._weights_seen = fp
        return super().train(mode)

    def invalidate_engines_for_weights(self):
        """After an optimizer step only the (eval) inference plans hold stale folded head weights; the frozen-backbone
        training engine reads BN statistics and conv weights that did not change."""
        from ppyolo_b200 import ops
        self._engines = {k: v for k, v in self._engines.items() if k[-1] == 'train'}
        ops._PACK_CACHE.clear()

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_engines()
        return super().load_state_dict(*args, **kwargs)

    def forward(self, x, im_size, eval=True, gt_box=None, gt_label=None, gt_score=None, targets=None):
        if eval and self.use_engine and not self.training and x.is_cuda and not self._has_mish():
            if x.dtype == torch.uint8:       # resized uint8 RGB batch [N, H, W, 3]: normalisation + permute inside the stem kernel
                n, h, w, _ = x.shape
                return self.engine(n, h, w, input_u8=True).run(x, im_size)
            n, _, h, w = x.shape
            return self.engine(n, h, w).run(x, im_size)
        if not eval:
            return self.forward_train(x, gt_box, gt_label, gt_score, targets)
        body_feats = self.backbone(x)
        return self.head.get_prediction(body_feats, im_size)

    def forward_train(self, x, gt_box, gt_label, gt_score, targets):
        """Training forward (reference train.py:427 -> model/head.py:400-422): dict of scalar losses.

        The frozen backbone (``freeze_at=5`` in both configs) runs under no_grad on the kernel engine with BATCH-statistic
        BatchNorm, exactly like the reference's train-mode frozen BNs; the trainable head runs as differentiable tensor code
        (``ppyolo_b200.autograd_head``) so torch autograd provides its backward, and the losses are ``model/losses.py``.
        With ``freeze_at < 5`` the backbone's trainable stages join the autograd graph (``ppyolo_b200.autograd_backbone``)."""
        if not x.is_cuda:
            raise RuntimeError('ppyolo_b200: training needs CUDA tensors -- the backbone kernels have no CPU fallback')
        if any(p.requires_grad for p in self.backbone.parameters()):
            # freeze_at < 5 (reference model/resnet_vd.py:170-220): the trainable stages -- stage 5's DCNv2 units included -- run as
            # differentiable kernel calls (ppyolo_b200.autograd_backbone), bf16 operands, the frozen prefix under no_grad
            from ppyolo_b200 import autograd_backbone
            if self.train_head_impl == 'aten':
                raise NotImplementedError("an unfrozen backbone trains on the kernel path only (train_head_impl='kernels')")
            self.head.train_impl = 'kernels'
            if self.train_graph:
                return self._graphed_head_loss([x], gt_box, gt_label, gt_score, targets, full=True)
            feats = autograd_backbone.backbone_features(self.backbone, x, 'kernels')
            return self.head.get_loss_autograd(feats, gt_box, gt_label, gt_score, targets)
        n, _, h, w = x.shape
        self.head.train_impl = self.train_head_impl or ('kernels' if self.train_precision == 'bf16' else 'aten')
        with torch.no_grad():      # the kernel head reads the engine's bf16 NHWC feature buffers as they are
            feats = self.backbone_train_engine(n, h, w).run_backbone(x, native=self.head.train_impl == 'kernels')
        if self.train_graph:
            return self._graphed_head_loss(feats, gt_box, gt_label, gt_score, targets)
        return self.head.get_loss_autograd(feats, gt_box, gt_label, gt_score, targets)

    def _graphed_head_loss(self, feats, gt_box, gt_label, gt_score, targets, full=False):
        """Head forward + the six losses (and, through autograd, their backward) replayed as CUDA graphs: the step is
        launch-bound (hundreds of small loss / BatchNorm / activation kernels), the graphs remove the launches.  One pair of
        graphs per input shape (``torch.cuda.make_graphed_callables``); BatchNorm buffers are restored after the capture's
        warm-up passes so capturing does not advance the running statistics.  ``full``: ``feats`` is ``[x]`` and the graphs hold
        the differentiable backbone (``freeze_at < 5``) as well."""
        head = self.head
        from model.custom_layers import DropBlock
        # everything the captured kernels depend on besides the tensors: shapes, conv implementation, BatchNorm mode of every unit
        # (batch vs running statistics) and the DropBlock switches -- toggling one of them must not replay a stale graph
        cached = self._mode_mods.get(bool(full))
        if cached is None:                       # the module tree is static: walk it once
            mods = list(head.modules()) + (list(self.backbone.modules()) if full else [])
            cached = ([m for m in mods if isinstance(m, torch.nn.BatchNorm2d)], [m for m in mods if isinstance(m, DropBlock)])
            self._mode_mods[bool(full)] = cached
        modes = tuple(m.training for m in cached[0]) + tuple(bool(m.is_test) for m in cached[1])
        key = (bool(full), tuple(tuple(f.shape) for f in feats), tuple(gt_box.shape), tuple(tuple(t.shape) for t in targets), head.train_impl, modes)
        entry = self._graphed_heads.get(key)
        if entry is None:
            n_feats, n_targets = len(feats), len(targets)
            backbone = self.backbone

            class HeadLoss(torch.nn.Module):
                def __init__(self, head):
                    super().__init__()
                    self.head = head
                    if full:
                        self.backbone = backbone

                def forward(self, *args):
                    fs, rest = list(args[:n_feats]), args[n_feats:]
                    gb, gl, gs = rest[0], rest[1], rest[2]
                    tg = list(rest[3:3 + n_targets])
                    if full:
                        from ppyolo_b200 import autograd_backbone
                        fs = autograd_backbone.backbone_features(self.backbone, fs[0], 'kernels')
                    losses = self.head.get_loss_autograd(fs, gb, gl, gs, tg)
                    return tuple(losses[k] for k in sorted(losses))

            wrapper = HeadLoss(head)
            sample = tuple(f.detach().clone() for f in feats) + (gt_box.clone(), gt_label.clone(), gt_score.clone()) + \
                tuple(t.clone() for t in targets)
            saved = {k: v.clone() for k, v in wrapper.state_dict().items() if 'running_' in k or 'num_batches_tracked' in k}
            yl = head.yolo_loss
            names = sorted(['loss_xy', 'loss_wh', 'loss_obj', 'loss_cls'] + (['loss_iou'] if yl._iou_loss is not None else []) +
                           (['loss_iou_aware'] if yl._iou_aware_loss is not None else []))
            graphed = torch.cuda.make_graphed_callables(wrapper, sample, allow_unused_input=True)
            wrapper.load_state_dict(saved, strict=False)
            entry = (graphed, names)
            self._graphed_heads[key] = entry
        graphed, names = entry
        out = graphed(*[f.detach() for f in feats], gt_box, gt_label, gt_score, *targets)
        return dict(zip(names, out))

    def backbone_train_engine(self, batch, height, width):
        from ppyolo_b200.engine import InferenceEngine
        key = (batch, height, width, self.train_precision, bool(self.backbone.training), bool(self.train_graph), 'train')
        eng = self._engines.get(key)
        if eng is None:
            eng = InferenceEngine(self, batch, height, width, precision=self.train_precision, dcn_impl=self.dcn_impl,
                                  train_bn=self.backbone.stage1_conv1_1.bn is not None and self.backbone.training,
                                  backbone_only=True, use_graph=self.train_graph)
            self._engines[key] = eng
        return eng

    def add_param_group(self, param_groups, base_lr, base_wd):
        self.backbone.add_param_group(param_groups, base_lr, base_wd)
        self.head.add_param_group(param_groups, base_lr, base_wd)

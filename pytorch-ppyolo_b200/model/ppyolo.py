"""Top-level detector module (reference ``model/ppyolo.py`` :13-29), same forward signature.

In eval mode on a CUDA device the whole network runs through ``ppyolo_b200.engine.InferenceEngine``:
a static plan of fused sm_100a kernels over pre-allocated NHWC buffers (built lazily per input shape,
replayed as a CUDA graph).  ``precision`` selects the arithmetic of the conv kernels:
``'bf16'`` (tcgen05 tensor cores, fp32 accumulate; the throughput path) or ``'fp32'`` (SIMT fp32,
the 1e-4 parity path).
"""
import torch


class PPYOLO(torch.nn.Module):
    def __init__(self, backbone, head):
        super().__init__()
        self.backbone = backbone
        self.head = head
        self.precision = 'bf16'
        self.dcn_impl = None          # None = engine default; 'fused' | 'gather_gemm'
        self.use_engine = True
        self._engines = {}

    def engine(self, batch, height, width):
        from ppyolo_b200.engine import InferenceEngine
        key = (batch, height, width, self.precision, self.dcn_impl)
        eng = self._engines.get(key)
        if eng is None:
            eng = InferenceEngine(self, batch, height, width, precision=self.precision, dcn_impl=self.dcn_impl)
            self._engines[key] = eng
        return eng

    def invalidate_engines(self):
        """Drop compiled plans (call after mutating weights, e.g. load_state_dict)."""
        self._engines = {}

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_engines()
        return super().load_state_dict(*args, **kwargs)

    def forward(self, x, im_size, eval=True, gt_box=None, gt_label=None, gt_score=None, targets=None):
        if eval and self.use_engine and not self.training and x.is_cuda:
            n, _, h, w = x.shape
            return self.engine(n, h, w).run(x, im_size)
        body_feats = self.backbone(x)
        if eval:
            return self.head.get_prediction(body_feats, im_size)
        return self.head.get_loss(body_feats, gt_box, gt_label, gt_score, targets)

    def add_param_group(self, param_groups, base_lr, base_wd):
        self.backbone.add_param_group(param_groups, base_lr, base_wd)
        self.head.add_param_group(param_groups, base_lr, base_wd)

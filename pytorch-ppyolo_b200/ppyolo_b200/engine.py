"""Whole-network inference engine: PPYOLO.forward(eval=True) as a static plan of fused sm_100a kernels.

Built once per (batch, H, W, precision) from the module tree (so it sees the same parameters a loaded
reference checkpoint provides), then replayed -- as a CUDA graph -- for every batch:

  * activations live in pre-allocated NHWC buffers (bf16 on the tcgen05 path, fp32 on the parity path);
    180 GB of HBM means no buffer reuse games are needed even at bs=32 x 608^2
  * every Conv2dUnit is ONE kernel: implicit GEMM + folded BN/bias + activation, with the block's residual
    add + ReLU, the head's nearest-x2 upsample and the concat placement (channel-slice writes) fused into
    its epilogue; CoordConv is folded into a per-pixel bias map computed once at build time
    (its two channels are constants of the feature-map shape), so no K padding and no concat at run time
  * DCNv2 = offset/mask conv (fp32 out) + the fused deformable-gather GEMM
  * decode = one kernel per scale writing straight into the batch-wide boxes/scores buffers;
    Matrix-NMS = 3 launches for the whole batch; the only host sync is the final read of the counts.

Reference path being replaced: model/ppyolo.py:19-22 -> model/resnet_vd.py:132-168 (or :302-330) ->
model/head.py:424-469.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import lib, check, ConvParams, PPY_F32, PPY_BF16, PPY_F16X2


DEFAULT_NORMALIZE = dict(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], is_scale=True)      # config/ppyolo_2x.py:205-210


def normalize_lut(mean, std, is_scale=True):
    """[3][256] float32 table of the reference's NormalizeImage (tools/transform.py:906-917, is_channel_first=False) applied to
    every byte value: the same numpy expression on a 256 x 1 x 3 uint8 image (float32 array, float64 mean/std operands of the
    in-place ops), so a table lookup reproduces the reference's floats bit for bit."""
    im = np.tile(np.arange(256, dtype=np.uint8)[:, None, None], (1, 1, 3))
    im = im.astype(np.float32, copy=False)
    if is_scale:
        im = im / 255.0
    im -= np.array(mean)[np.newaxis, np.newaxis, :]
    im /= np.array(std)[np.newaxis, np.newaxis, :]
    return np.ascontiguousarray(im[:, 0, :].T.astype(np.float32))


class TensorRef(object):
    """A channel slice [c_off, c_off+c) of an NHWC buffer [N,H,W,ld] -- or of a PPY_F16X2 buffer [2,N,H,W,ld] (fp16 hi plane,
    lo plane; ``plane`` = element offset between them)."""

    def __init__(self, t, c=None, c_off=0):
        self.t, self.c_off = t, c_off
        self.c = t.shape[-1] - c_off if c is None else c

    pair = property(lambda self: self.t.dim() == 5)
    n = property(lambda self: self.t.shape[-4])
    h = property(lambda self: self.t.shape[-3])
    w = property(lambda self: self.t.shape[-2])
    ld = property(lambda self: self.t.shape[-1])
    plane = property(lambda self: self.t.stride(0) if self.t.dim() == 5 else 0)
    esz = property(lambda self: self.t.element_size() * (2 if self.t.dim() == 5 else 1))      # bytes per value
    code = property(lambda self: PPY_F16X2 if self.t.dim() == 5 else (PPY_BF16 if self.t.dtype == torch.bfloat16 else PPY_F32))

    def pixel_pairs(self):
        """[.., H, W, C] seen as [.., H, W/2, 2C] (free reinterpretation)."""
        shp = tuple(self.t.shape[:-2]) + (self.t.shape[-2] // 2, 2 * self.t.shape[-1])
        return TensorRef(self.t.view(shp))

    def pixel_unpairs(self, c):
        shp = tuple(self.t.shape[:-2]) + (self.t.shape[-2] * (self.t.shape[-1] // c), c)
        return TensorRef(self.t.view(shp))

    @property
    def ptr(self):
        return self.t.data_ptr() + self.c_off * self.t.element_size()

    def slice(self, c_off, c):
        return TensorRef(self.t, c, self.c_off + c_off)


class InferenceEngine(object):
    def __init__(self, model, batch, height, width, precision='bf16', use_graph=True, dcn_impl=None, train_bn=False,
                 backbone_only=False, input_u8=False):
        if not torch.cuda.is_available():
            raise RuntimeError('InferenceEngine needs a CUDA device (no CPU fallback)')
        if precision not in ops.PRECISIONS:
            raise ValueError(precision)
        if precision in ('bf16', 'f16x2') and not lib.ppy_conv_bf16_supported():
            raise RuntimeError('the tcgen05 conv path needs an sm_100 device')
        if height % 32 or width % 32 or height != width:
            raise ValueError('input must be square with a side that is a multiple of 32 (reference yolo_box '
                             'assumes square maps, model/head.py:25-27)')
        self.model, self.n, self.h, self.w = model, batch, height, width
        self.precision, self.code = precision, ops.dtype_code(precision)
        self.dev = next(model.parameters()).device
        if self.dev.type != 'cuda':
            raise RuntimeError('model must be on a CUDA device')
        self.act_dtype = ops.torch_dtype(self.code)
        # DCNv2: 'fused' (default) = the im2col-free whole-layer kernel (csrc/dcn_umma.cu: CTA pairs, full-N accumulators,
        # coalesced corner gather; the generic producer mode of the conv kernel for layers it does not fit);
        # 'gather_gemm' (bf16 / fp32 only) = sampling kernel -> [M x 9C] matrix -> TMA-fed 1x1 GEMM (round 1's default)
        self.dcn_impl = dcn_impl or getattr(model, 'dcn_impl', None) or 'fused'
        if precision == 'f16x2':
            self.dcn_impl = 'fused'            # (the sampling kernel has no pair variant)
            if train_bn:
                raise NotImplementedError('batch-statistic BatchNorm engines run in bf16 or fp32')
        # 'f16x2': every activation is stored as act_scale * value (a power of two, exact): lifts typical O(1) activations away
        # from the fp16 subnormal range of their lo parts; values beyond 65504 / act_scale raise the overflow flag
        self.act_scale = float(getattr(model, 'f16x2_act_scale', 8.0)) if precision == 'f16x2' else 1.0
        # train_bn: BatchNorm layers normalise with BATCH statistics and update their running stats, as the reference's
        # frozen backbone does during training; backbone_only: stop at the C3/C4/C5 feature maps
        self.train_bn, self.backbone_only = train_bn, backbone_only
        # input_u8: the static input is the RESIZED uint8 RGB batch [n, h, w, 3] (a quarter of the upload bytes); the reference's
        # NormalizeImage + Permute run inside the stem kernel through a 3 x 256 table built with the reference's numpy expression
        self.input_u8 = bool(input_u8)
        self.postprocess_impl = getattr(model, 'postprocess_impl', None) or 'dense'
        if self.postprocess_impl not in ('sparse', 'dense'):
            raise ValueError(self.postprocess_impl)
        self.steps = []          # (name, callable)
        self.keep = []           # tensors / ctypes structs referenced by raw pointer
        self.conv_flops = 0      # algorithmic 2*MAC of all convs in the plan (per batch)
        self.graph = None
        self.launches_per_run = 0
        self.bn_modules = []
        self.step_info = {}      # per conv step: algorithmic flops / minimum HBM bytes / GEMM shape
        with torch.no_grad():
            self._build()
        # one eager pass: first-use initialisation (func attributes, tensor maps) + launch count
        # (train_bn engines update BatchNorm running statistics: this dry run on a zero batch must leave them untouched)
        saved = [(bn, bn.running_mean.clone(), bn.running_var.clone()) for bn in self.bn_modules]
        before = _lib.launch_count()
        self._run_steps()
        torch.cuda.synchronize(self.dev)
        self.launches_per_run = _lib.launch_count() - before
        for bn, mean, var in saved:
            bn.running_mean.copy_(mean)
            bn.running_var.copy_(var)
        if use_graph:
            self.graph = self._capture()
            self.graphs = [self.graph]

    # ------------------------------------------------------------------ buffers
    def _new(self, n, h, w, c, dtype=None):
        if dtype is None and self.code == PPY_F16X2:       # activation of the pair path: hi plane | lo plane
            t = torch.zeros((2, n, h, w, c), dtype=torch.float16, device=self.dev)
        else:
            t = torch.zeros((n, h, w, c), dtype=dtype or self.act_dtype, device=self.dev)
        self.keep.append(t)
        return t

    def _keep(self, t):
        self.keep.append(t)
        return t

    # ------------------------------------------------------------------ op builders
    def _add(self, name, fn):
        self.steps.append((name, fn))

    def _simple(self, name, cfn, x, out, *extra):
        """pool-like op: cfn(x_ptr, x_ld, y_ptr, y_ld, n, h, w, c, dtype, stream)"""
        if x.pair:
            cfn = getattr(lib, cfn.__name__ + '_f16x2')
            args = (ctypes.c_void_p(x.ptr), x.ld, x.plane, ctypes.c_void_p(out.ptr), out.ld, out.plane, x.n, x.h, x.w, x.c)
        else:
            args = (ctypes.c_void_p(x.ptr), x.ld, ctypes.c_void_p(out.ptr), out.ld, x.n, x.h, x.w, x.c, x.code)

        def run():
            check(cfn(*args, ops.stream_ptr()), name)
        self._add(name, run)
        return out

    def _maxpool(self, x):
        out = TensorRef(self._new(x.n, (x.h + 1) // 2, (x.w + 1) // 2, x.c))
        return self._simple('maxpool3x3s2', lib.ppy_maxpool3x3s2, x, out)

    def _avgpool(self, x):
        out = TensorRef(self._new(x.n, x.h // 2, x.w // 2, x.c))
        return self._simple('avgpool2x2', lib.ppy_avgpool2x2, x, out)

    def _spp(self, x, seq='asc'):
        if seq != 'asc':
            raise NotImplementedError("SPP(seq='desc') is not on any config's path (model/custom_layers.py:286-289)")
        out = TensorRef(self._new(x.n, x.h, x.w, 4 * x.c))
        return self._simple('spp', lib.ppy_spp, x, out)

    def _stem(self, unit, raw=False):
        """``raw``: the conv output itself (scale 1, shift 0, no activation) -- the train-mode BatchNorm kernel follows."""
        from model.custom_layers import ACT_CODES
        scale, shift = unit.folded_scale_shift()
        if raw:
            scale, shift = torch.ones_like(scale), torch.zeros_like(shift)
        w_host = np.ascontiguousarray(unit.conv.weight.detach().float().cpu().numpy())
        sc_host = np.ascontiguousarray(scale.cpu().numpy())
        sh_host = np.ascontiguousarray(shift.cpu().numpy())
        if self.act_scale != 1.0:            # pair path: activations are stored scaled (the activation commutes with it)
            sc_host, sh_host = sc_host * np.float32(self.act_scale), sh_host * np.float32(self.act_scale)
        self.keep += [w_host, sc_host, sh_host]
        ho, wo = (self.h - 1) // 2 + 1, (self.w - 1) // 2 + 1
        out = TensorRef(self._new(self.n, ho, wo, 32))
        fp = ctypes.POINTER(ctypes.c_float)
        wargs = (w_host.ctypes.data_as(fp), sc_host.ctypes.data_as(fp), sh_host.ctypes.data_as(fp), 32, 0 if raw else ACT_CODES[unit.act_name],
                 ctypes.c_void_p(out.ptr), out.ld)
        if self.input_u8:
            lut = self._keep(torch.from_numpy(normalize_lut(**self.normalize)).to(self.dev))
            args = (self.n, self.h, self.w, ops.ptr(lut)) + wargs + (self.code, out.plane)
            stem_fn = lib.ppy_stem_conv3x3s2_u8
        else:
            args = (self.n, self.h, self.w) + wargs + ((out.plane,) if out.pair else (self.code,))
            stem_fn = lib.ppy_stem_conv3x3s2_f16x2 if out.pair else lib.ppy_stem_conv3x3s2

        def run():            # reads the CURRENT input slot (one captured graph per slot, see add_input_slot)
            check(stem_fn(ops.ptr(self._src), *args, ops.stream_ptr()), 'stem.conv1_1')
        self._add('stem.conv1_1', run)
        self.conv_flops += 2 * self.n * ho * wo * 32 * 27
        return out

    def _coord_bias_map(self, weight, c_main, h, w):
        """Contribution of the two CoordConv channels for one image: [h*w, cout] fp32 (pre-BN)."""
        cout, _, k, _ = weight.shape
        coords = self._new(1, h, w, 8, torch.float32)
        check(lib.ppy_coord_channels(ops.ptr(coords), 8, h, w, PPY_F32, ops.stream_ptr()), 'coord_channels')
        packed = ops.pack_weight(weight, PPY_F32, c_begin=c_main, c_count=2)
        one = self._keep(torch.ones(cout, dtype=torch.float32, device=self.dev))
        zero = self._keep(torch.zeros(cout, dtype=torch.float32, device=self.dev))
        out = self._new(1, h, w, cout, torch.float32)
        ops.conv_nhwc(coords, packed, 2, cout, k, 1, (k - 1) // 2, one, zero, 0, PPY_F32, out=out, out_code=PPY_F32)
        return out

    def _conv(self, name, x, weight, scale, shift, stride, act, residual=None, dst=None, coord=False, upsample=False,
              out_code=None, offset_mask=None, gemm_taps=False, split_k=False):
        """``gemm_taps``: ``x`` already is the sampled [n,ho,wo,k*k*cin] matrix of a DCN gather; run the k x k
        weight as a 1x1 GEMM over it (same packed K order (tap, c))."""
        cout, cin_total, k, _ = weight.shape
        if gemm_taps:
            return self._conv_over_taps(name, x, weight, scale, shift, act, residual, dst)
        c_main = cin_total - (2 if coord else 0)
        if c_main != x.c and not (c_main < x.c and x.c == ops.round_up(c_main, 8)):
            raise ValueError('%s: weight expects %d input channels, buffer has %d' % (name, c_main, x.c))
        pad = (k - 1) // 2
        pair = self.code == PPY_F16X2
        chan_scale = None
        # pair path: CoordConv's two channels enter as ONE extra K block from a batch-invariant second source (ppy_conv_params.x2)
        # instead of a per-pixel fp32 bias map in the epilogue (1x1 convs, and the 3x3 convs that run the patch / im2col loaders)
        coord_x2 = (pair and coord and stride == 1 and c_main % 64 == 0 and x.c == c_main and offset_mask is None and residual is None and
                    (k == 1 or (k == 3 and cout > 128)) and x.h > 1 and x.w > 1 and getattr(self.model, 'coord_as_k', True))
        x2 = None
        if pair:
            wc = weight.detach().float()[:, c_main:c_main + 2] if coord_x2 else None
            packed, cin_pad, k_pad, cout_pad, chan_scale = ops.pack_weight_pair(
                weight, 0, c_main, amax_with=wc.abs().amax(dim=(1, 2, 3)) if coord_x2 else None)
            if coord_x2:
                extra = torch.zeros((cout, 64), dtype=torch.float32, device=self.dev)
                extra[:, :2 * k * k] = wc.permute(0, 2, 3, 1).reshape(cout, 2 * k * k)        # column 2*tap + {0: x, 1: y}
                packed = ops.append_weight_block(packed, cout, extra * chan_scale.view(-1, 1))
                k_pad += 64
                x2 = self._keep(ops.coord_source(x.h, x.w, k, self.act_scale, self.dev))
        else:
            packed, cin_pad, k_pad, cout_pad = ops.pack_weight(weight, self.code, c_begin=0, c_count=c_main)
        self._keep(packed)
        ho = (x.h + 2 * pad - k) // stride + 1
        wo = (x.w + 2 * pad - k) // stride + 1
        out_code = self.code if out_code is None else out_code
        if dst is None:
            oh, ow = (2 * ho, 2 * wo) if upsample else (ho, wo)
            dst = TensorRef(self._new(x.n, oh, ow, ops.round_up(cout, 8), None if out_code == self.code else ops.torch_dtype(out_code)),
                            c=cout)
        # CoordConv fold: a 1x1 conv sees the two coordinate channels as the rank-2 term wx*xc + wy*yc (two per-channel vectors
        # for the TMA epilogue); a 3x3 conv needs the per-pixel map (zero padding breaks the rank-2 structure at the borders)
        coord_vec = (coord and k == 1 and stride == 1 and self.code == PPY_BF16 and out_code == PPY_BF16 and not upsample and
                     cout >= 64 and cout % 8 == 0 and k_pad <= (512 if cout % 256 == 0 else 1152) and x.h > 1 and x.w > 1)
        bias_map = self._coord_bias_map(weight, c_main, x.h, x.w) if (coord and not coord_vec and not coord_x2) else None
        if pair:
            # the accumulator holds act_scale * chan_scale * (true pre-norm value): x is stored scaled by act_scale, every weight
            # row by chan_scale (both powers of two, so all of this is exact).  Pair outputs are stored scaled by act_scale again
            # (relu / leaky commute with a positive factor); fp32 outputs (head outputs, DCN offsets) are true values.
            a = self.act_scale
            to_acc = chan_scale * a
            if bias_map is not None:
                bias_map = self._keep((bias_map * to_acc).contiguous())
            if out_code == PPY_F16X2:
                scale, shift = scale / chan_scale, shift * a
            else:
                scale = scale / to_acc
            scale, shift = scale.contiguous(), shift.contiguous()
        coord_w = None
        if coord_vec:
            coord_w = self._keep(weight.detach().float()[:, c_main:c_main + 2, 0, 0].t().contiguous())    # [2][cout]: wx | wy
        p = ConvParams()
        p.x, p.x_ld = x.ptr, x.ld
        p.n, p.h, p.w, p.cin = x.n, x.h, x.w, cin_pad
        p.weight = packed.data_ptr()
        p.cout, p.kh, p.kw, p.stride, p.pad = cout, k, k, stride, pad
        p.k_pad, p.cout_pad = k_pad, cout_pad
        p.scale, p.shift = self._keep(scale).data_ptr(), self._keep(shift).data_ptr()
        p.bias_map = bias_map.data_ptr() if bias_map is not None else None
        p.coord_w = coord_w.data_ptr() if coord_w is not None else None
        p.residual = residual.ptr if residual is not None else None
        p.res_ld = residual.ld if residual is not None else 0
        p.act = act
        p.y, p.y_ld, p.out_dtype = dst.ptr, dst.ld, out_code
        p.upsample2x = 1 if upsample else 0
        p.offset_mask = offset_mask.ptr if offset_mask is not None else None
        p.om_ld = offset_mask.ld if offset_mask is not None else 0
        if pair:
            p.x_plane, p.y_plane = x.plane, dst.plane
            p.res_plane = residual.plane if residual is not None else 0
            p.overflow = self.overflow.data_ptr()
            if x2 is not None:
                p.x2, p.x2_ld, p.x2_plane = x2.data_ptr(), 64, x2.stride(0)
                p.x2_kb, p.x2_tiled, p.x2_row_mod, p.x2_rows = 1, 0, x.h * x.w, x2.shape[1]
        # split_k: few output tiles and a long K (the DCN offset conv): K splits added atomically into the zeroed fp32 output
        split_k = split_k and self.code == PPY_BF16 and out_code == PPY_F32 and act == 0 and residual is None and not coord
        p.accumulate = 1 if split_k else 0
        self._keep(p)
        fn = {PPY_BF16: lib.ppy_conv_bf16, PPY_F16X2: lib.ppy_conv_f16x2}.get(self.code, lib.ppy_conv_f32)
        ref = ctypes.byref(p)
        zero_t = dst.t

        def run():
            if split_k:
                zero_t.zero_()
            check(fn(ref, ops.stream_ptr()), name)
        self._add(name, run)
        flops = 2 * x.n * ho * wo * cout * cin_total * k * k
        self.conv_flops += flops
        esz_in, esz_out = x.esz, dst.esz
        nbytes = (x.n * x.h * x.w * c_main * esz_in + packed.numel() * packed.element_size() +
                  x.n * ho * wo * cout * esz_out * (4 if upsample else 1) + (x.n * ho * wo * cout * esz_out if residual is not None else 0))
        self.step_info[name] = {'flops': flops, 'bytes': nbytes, 'm': x.n * ho * wo, 'n': cout, 'k': cin_total * k * k}
        return dst

    def _conv_over_taps(self, name, xcol, weight, scale, shift, act, residual, dst):
        cout, cin, k, _ = weight.shape
        packed, cin_pad, k_pad, cout_pad = ops.pack_weight(weight, self.code)
        if cin_pad != cin or k_pad != k * k * cin or xcol.c != k * k * cin:
            raise ValueError('%s: gather_gemm needs cin %% 64 == 0' % name)
        self._keep(packed)
        if dst is None:
            dst = TensorRef(self._new(xcol.n, xcol.h, xcol.w, ops.round_up(cout, 8)), c=cout)
        p = ConvParams()
        p.x, p.x_ld = xcol.ptr, xcol.ld
        p.n, p.h, p.w, p.cin = xcol.n, xcol.h, xcol.w, k * k * cin
        p.weight = packed.data_ptr()
        p.cout, p.kh, p.kw, p.stride, p.pad = cout, 1, 1, 1, 0
        p.k_pad, p.cout_pad = k_pad, cout_pad
        p.scale, p.shift = self._keep(scale).data_ptr(), self._keep(shift).data_ptr()
        p.bias_map = None
        p.residual = residual.ptr if residual is not None else None
        p.res_ld = residual.ld if residual is not None else 0
        p.act = act
        p.y, p.y_ld, p.out_dtype = dst.ptr, dst.ld, self.code
        p.upsample2x, p.offset_mask, p.om_ld = 0, None, 0
        self._keep(p)
        fn = lib.ppy_conv_bf16 if self.code == PPY_BF16 else lib.ppy_conv_f32
        ref = ctypes.byref(p)

        def run():
            check(fn(ref, ops.stream_ptr()), name)
        self._add(name, run)
        flops = 2 * xcol.n * xcol.h * xcol.w * cout * cin * k * k
        self.conv_flops += flops
        m = xcol.n * xcol.h * xcol.w
        self.step_info[name] = {'flops': flops, 'bytes': (m * k * k * cin + packed.numel() + m * cout * (2 if residual is not None else 1)) * 2,
                                'm': m, 'n': cout, 'k': cin * k * k}
        return dst

    def _dcn_gather(self, name, x, om, k, stride):
        pad = (k - 1) // 2
        ho = (x.h + 2 * pad - k) // stride + 1
        wo = (x.w + 2 * pad - k) // stride + 1
        out = TensorRef(self._new(x.n, ho, wo, k * k * x.c))
        args = (ctypes.c_void_p(x.ptr), x.ld, x.n, x.h, x.w, x.c, ctypes.c_void_p(om.ptr), om.ld, k, stride, pad,
                ctypes.c_void_p(out.ptr), x.code)

        def run():
            check(lib.ppy_dcn_gather(*args, ops.stream_ptr()), name)
        self._add(name, run)
        return out

    def _head_output_conv(self, name, unit, x):
        """1x1 output conv to A*(5|6+C) channels, fp32 out.  258 channels would need a third, almost empty 128-wide N
        tile (re-reading the whole input for 2 columns); instead the first 256 channels run as full 256-wide tiles and
        the 2 leftover channels as a 32-wide launch into the same buffer."""
        cout = unit.filters
        if self.code == PPY_F32 or cout <= 256 or cout % 256 > 32 or unit.bn is not None:
            return self._unit(name, unit, x, out_code=PPY_F32)
        from model.custom_layers import ACT_CODES
        scale, shift = unit.folded_scale_shift()
        w = unit.conv.weight.detach()
        out = TensorRef(self._new(x.n, x.h, x.w, ops.round_up(cout, 8), torch.float32), c=cout)
        act = ACT_CODES[unit.act_name]
        main = (cout // 256) * 256
        self._conv(name, x, w[:main].contiguous(), scale[:main].contiguous(), shift[:main].contiguous(), 1, act,
                   dst=out.slice(0, main), out_code=PPY_F32)
        # the tail runs 8 channels wide (zero weights past cout, landing in the buffer's padding columns) so its epilogue
        # takes the aligned 16-byte vector path; the extra columns are never read and not counted as algorithmic FLOPs
        tail = cout - main
        wide = min(ops.round_up(tail, 8), out.ld - main)
        wt = torch.zeros((wide,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
        wt[:tail] = w[main:]
        sc = torch.ones(wide, dtype=scale.dtype, device=scale.device)
        sh = torch.zeros(wide, dtype=shift.dtype, device=shift.device)
        sc[:tail], sh[:tail] = scale[main:], shift[main:]
        before = self.conv_flops
        self._conv(name + '.tail', x, wt, sc, sh, 1, act, dst=out.slice(main, wide), out_code=PPY_F32)
        true_flops = (self.conv_flops - before) * tail // wide
        self.conv_flops = before + true_flops
        self.step_info[name + '.tail']['flops'] = true_flops
        self.step_info[name + '.tail']['n'] = tail
        return out

    def _unit_pixel_pairs(self, name, unit, x):
        """3x3 stride-1 conv with 32 input channels run on the free reinterpretation [N,H,W,32] == [N,H,W/2,64]:
        two horizontally adjacent pixels form one 64-channel "pixel", so every K block is a full 128-byte row and
        the 4-D TMA patch loader applies.  The 3x3 weight becomes a 3x3 weight over pixel pairs
        (W2[(px,co),(qx,ci),ky,d+1] = W[co,ci,ky,2d+qx-px+1], zero where that tap does not exist); the output
        [N,H,W/2,2*cout] is bit-for-bit the [N,H,W,cout] tensor.  Half of the issued MACs multiply structural
        zeros -- the layer is HBM/L2-bound, the tensor pipe has the headroom -- and conv_flops counts the original."""
        from model.custom_layers import ACT_CODES
        cout = unit.conv.weight.shape[0]
        w2 = self._pixel_pair_weight(unit.conv.weight)
        scale, shift = unit.folded_scale_shift()
        scale2, shift2 = torch.cat([scale, scale]).contiguous(), torch.cat([shift, shift]).contiguous()
        xp = x.pixel_pairs()
        flops_before = self.conv_flops
        out = self._conv(name, xp, self._keep(w2), scale2, shift2, 1, ACT_CODES[unit.act_name])
        self.conv_flops = flops_before + 2 * x.n * x.h * x.w * cout * 32 * 9
        return out.pixel_unpairs(cout)

    @staticmethod
    def _pixel_pair_weight(weight):
        """[cout, cin, 3, 3] -> the 3x3 weight over pixel pairs [2 cout, 2 cin, 3, 3] of ``_unit_pixel_pairs``."""
        w = weight.detach().float()
        cout, cin, k, _ = w.shape
        w2 = torch.zeros((2 * cout, 2 * cin, 3, 3), dtype=torch.float32, device=w.device)
        for px in range(2):
            for qx in range(2):
                for d in (-1, 0, 1):
                    kx = 2 * d + qx - px + 1
                    if 0 <= kx <= 2:
                        w2[px * cout:(px + 1) * cout, qx * cin:(qx + 1) * cin, :, d + 1] = w[:, :, :, kx]
        return w2

    def _unit_batch_stats(self, name, unit, x, residual, act, dst, coord, raw=None):
        """conv (raw) -> per-channel batch statistics (+ running-stat update) -> normalise + residual + act."""
        from model.custom_layers import DCNv2
        bn = unit.bn
        cout = bn.num_features
        one = torch.ones(cout, dtype=torch.float32, device=self.dev)
        zero = torch.zeros(cout, dtype=torch.float32, device=self.dev)
        if raw is not None:
            pass                               # (the caller ran the conv: the fused stem kernel)
        elif isinstance(unit.conv, DCNv2):
            d = unit.conv
            n_om = d.conv_offset.weight.shape[0]
            om = self._conv(name + '.offset', x, d.conv_offset.weight.detach(), torch.ones(n_om, dtype=torch.float32, device=self.dev),
                            d.conv_offset.bias.detach().float().contiguous(), unit.stride, 0, out_code=PPY_F32)
            om = TensorRef(om.t)
            bias = d.dcn_bias.detach().float() if d.dcn_bias is not None else zero
            if self.dcn_impl == 'gather_gemm' and x.c % 64 == 0:
                xcol = self._dcn_gather(name + '.gather', x, om, d.dcn_weight.shape[-1], unit.stride)
                raw = self._conv(name + '.raw', xcol, d.dcn_weight.detach(), one, bias, 1, 0, gemm_taps=True)
            else:
                raw = self._conv(name + '.raw', x, d.dcn_weight.detach(), one, bias, unit.stride, 0, offset_mask=om)
        else:
            bias = unit.conv.bias.detach().float() if unit.conv.bias is not None else zero
            pairable = (self.code == PPY_BF16 and unit.stride == 1 and tuple(unit.conv.weight.shape[1:]) == (32, 3, 3) and x.c == 32 and
                        x.ld == 32 and x.c_off == 0 and x.w % 2 == 0 and cout % 8 == 0 and unit.conv.bias is None and not coord and
                        not os.environ.get('PPY_NO_TRAIN_PIXEL_PAIRS'))
            if pairable:
                # the stem's 32-channel 3x3 convs on the pixel-pair view (see _unit_pixel_pairs): full 128-byte K rows through the
                # 4-D TMA patch loader instead of the cp.async gather mode (2 launches of a bs-8 step: 0.39 -> ~0.1 ms); the raw
                # output [N,H,W/2,2 cout] IS the [N,H,W,cout] tensor the statistics kernel reads
                flops_before = self.conv_flops
                raw = self._conv(name + '.raw', x.pixel_pairs(), self._keep(self._pixel_pair_weight(unit.conv.weight)),
                                 torch.cat([one, one]).contiguous(), torch.cat([zero, zero]).contiguous(), 1, 0).pixel_unpairs(cout)
                self.conv_flops = flops_before + 2 * x.n * x.h * x.w * cout * 32 * 9
            else:
                raw = self._conv(name + '.raw', x, unit.conv.weight.detach(), one, bias, unit.stride, 0, coord=coord)
        if dst is None:
            dst = TensorRef(self._new(raw.n, raw.h, raw.w, ops.round_up(cout, 8)), c=cout)
        scale = self._keep(torch.empty(cout, dtype=torch.float32, device=self.dev))
        shift = self._keep(torch.empty(cout, dtype=torch.float32, device=self.dev))
        ws = self._keep(torch.zeros(32 * cout + 1, dtype=torch.float64, device=self.dev))      # PPY_BN_WORKSPACE_DOUBLES(cout)
        rows = raw.n * raw.h * raw.w
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        a1 = (ctypes.c_void_p(raw.ptr), raw.ld, rows, cout, raw.code, ops.ptr(bn.weight.data), ops.ptr(bn.bias.data), float(bn.eps),
              momentum, ops.ptr(bn.running_mean), ops.ptr(bn.running_var), ops.ptr(scale), ops.ptr(shift), ops.ptr(ws))
        a2 = (ctypes.c_void_p(raw.ptr), raw.ld, ctypes.c_void_p(dst.ptr), dst.ld, rows, cout, raw.code, ops.ptr(scale), ops.ptr(shift),
              ctypes.c_void_p(residual.ptr) if residual is not None else ctypes.c_void_p(0),
              residual.ld if residual is not None else 0, act)
        if cout <= 2048 and not os.environ.get('PPY_NO_BN_FUSED'):
            # statistics + normalise + activation (+ residual) as ONE cooperative launch (workspace self-cleaning: no memset node)
            a3 = (ctypes.c_void_p(raw.ptr), raw.ld, ctypes.c_void_p(dst.ptr), dst.ld, rows, cout, raw.code, ops.ptr(bn.weight.data),
                  ops.ptr(bn.bias.data), float(bn.eps), momentum, ops.ptr(bn.running_mean), ops.ptr(bn.running_var), ops.ptr(scale),
                  ops.ptr(shift), ctypes.c_void_p(residual.ptr) if residual is not None else ctypes.c_void_p(0),
                  residual.ld if residual is not None else 0, act, ops.ptr(ws), ctypes.c_void_p(0), ctypes.c_void_p(0))
            self._add(name + '.bn_fused', lambda: check(lib.ppy_bn_train_fused(*a3, ops.stream_ptr()), name + '.bn_fused'))
        else:
            self._add(name + '.bn_stats', lambda: check(lib.ppy_bn_batch_stats(*a1, ops.stream_ptr()), name + '.bn_stats'))
            self._add(name + '.bn_apply', lambda: check(lib.ppy_scale_shift_act(*a2, ops.stream_ptr()), name + '.bn_apply'))
        self.bn_modules.append(bn)
        return dst

    def _unit(self, name, unit, x, residual=None, act=None, dst=None, coord=False, upsample=False, out_code=None):
        """One Conv2dUnit (conv|DCNv2 -> folded norm -> act) as one (DCN: two) kernels."""
        from model.custom_layers import DCNv2, ACT_CODES
        act = ACT_CODES[unit.act_name] if act is None else act
        if self.train_bn and unit.bn is not None:
            return self._unit_batch_stats(name, unit, x, residual, act, dst, coord)
        scale, shift = unit.folded_scale_shift()
        if isinstance(unit.conv, DCNv2):
            d = unit.conv
            n_om = d.conv_offset.weight.shape[0]
            one = torch.ones(n_om, dtype=torch.float32, device=self.dev)
            om = self._conv(name + '.offset', x, d.conv_offset.weight.detach(), one,
                            d.conv_offset.bias.detach().float().contiguous(), unit.stride, 0, out_code=PPY_F32)
            # (no split-K here: atomically added partial sums make the offsets, hence the detections, run-to-run non-deterministic)
            om = TensorRef(om.t)     # the sampler reads the padded row (ld) directly
            # (folded_scale_shift() already carries a DCN bias through the norm: nothing to add here)
            if self.dcn_impl == 'gather_gemm' and x.c % 64 == 0:
                xcol = self._dcn_gather(name + '.gather', x, om, d.dcn_weight.shape[-1], unit.stride)
                return self._conv(name, xcol, d.dcn_weight.detach(), scale, shift, 1, act, residual=residual, dst=dst,
                                  gemm_taps=True)
            return self._conv(name, x, d.dcn_weight.detach(), scale, shift, unit.stride, act, residual=residual,
                              dst=dst, offset_mask=om)
        return self._conv(name, x, unit.conv.weight.detach(), scale, shift, unit.stride, act, residual=residual,
                          dst=dst, coord=coord, upsample=upsample, out_code=out_code)

    def _dual_ok(self, conv3, y, sc_unit, sc_in):
        """`conv3(y) + shortcut` of a bottleneck as ONE pair GEMM with a second K source (ppy_conv_params.x2)?"""
        from model.custom_layers import DCNv2
        if self.code != PPY_F16X2 or self.train_bn or not getattr(self.model, 'fuse_shortcut', True):
            return False
        units = [conv3] + ([sc_unit] if sc_unit is not None else [])
        for u in units:
            if (isinstance(u.conv, DCNv2) or u.conv.weight.shape[-1] != 1 or u.stride != 1 or u.conv.weight.shape[1] % 64 or
                    u.act_name is not None):
                return False
        if y.c != conv3.conv.weight.shape[1] or sc_in.n != y.n or sc_in.h != y.h or sc_in.w != y.w:
            return False
        if sc_unit is None:
            return sc_in.c == conv3.filters and ops.ident_tile(conv3.filters, y.c) is not None
        return sc_in.c == sc_unit.conv.weight.shape[1]

    def _conv_dual(self, name, conv3, y, sc_in, sc_unit=None, act=0, dst=None):
        """act(norm3(conv3(y)) + shortcut) as one launch of the pair kernel: the shortcut is a second K source -- the projection
        conv4 K-concatenated (both folded norm scales multiplied into the weight rows), or the identity shortcut as identity
        weight blocks so the tensor core does the residual add and the residual travels through the TMA operand pipeline
        (reference model/resnet_vd.py:44-56, :75-86).  Returns None when the packer declines (tiny weights): caller falls back."""
        cout, cin = conv3.filters, y.c
        s3, b3 = conv3.folded_scale_shift()
        w3 = conv3.conv.weight.detach()
        if sc_unit is not None:
            s4, b4 = sc_unit.folded_scale_shift()
            res = ops.pack_weight_pair_dual(w3, s3, sc_unit.conv.weight.detach(), s4)
            shift = b3 + b4
            k_alg = cin + sc_in.c
        else:
            res = ops.pack_weight_pair_dual(w3, s3, ident_bn=ops.ident_tile(cout, cin))
            shift = b3
            k_alg = cin
        if res is None:
            return None
        packed, k_pad, cout_pad, chan_scale = res
        self._keep(packed)
        if dst is None:
            dst = TensorRef(self._new(y.n, y.h, y.w, ops.round_up(cout, 8)), c=cout)
        # accumulator = act_scale * chan_scale * (true value - shift); pair outputs are stored times act_scale
        scale = self._keep((1.0 / chan_scale).contiguous())
        shift = self._keep((shift.float() * self.act_scale).contiguous())
        p = ConvParams()
        p.x, p.x_ld, p.x_plane = y.ptr, y.ld, y.plane
        p.n, p.h, p.w, p.cin = y.n, y.h, y.w, cin
        p.weight = packed.data_ptr()
        p.cout, p.kh, p.kw, p.stride, p.pad = cout, 1, 1, 1, 0
        p.k_pad, p.cout_pad = k_pad, cout_pad
        p.scale, p.shift = scale.data_ptr(), shift.data_ptr()
        p.act = act
        p.y, p.y_ld, p.out_dtype, p.y_plane = dst.ptr, dst.ld, PPY_F16X2, dst.plane
        p.x2, p.x2_ld, p.x2_plane = sc_in.ptr, sc_in.ld, sc_in.plane
        p.x2_kb, p.x2_tiled = (k_pad - cin) // 64, 0 if sc_unit is not None else 1
        p.overflow = self.overflow.data_ptr()
        self._keep(p)
        ref = ctypes.byref(p)
        self._add(name, lambda: check(lib.ppy_conv_f16x2(ref, ops.stream_ptr()), name))
        m = y.n * y.h * y.w
        flops = 2 * m * cout * k_alg
        self.conv_flops += flops
        nbytes = m * (cin + sc_in.c) * y.esz + packed.numel() * packed.element_size() + m * cout * dst.esz
        self.step_info[name] = {'flops': flops, 'bytes': nbytes, 'm': m, 'n': cout, 'k': k_alg}
        return dst

    def _pool_into_conv_ok(self, unit, x):
        """AvgPool2d(2,2) + 1x1 conv of the vd shortcut (model/resnet_vd.py:29-33) as ONE 2x2 / stride-2 conv with the weight
        replicated over the window and divided by 4 (exact in bf16): the input is read once instead of pooled, written and
        read again.  The GEMM does 4x the MACs, so only where the shortcut is clearly HBM-bound (large maps)."""
        from model.custom_layers import DCNv2
        return (self.code == PPY_BF16 and not self.train_bn and not isinstance(unit.conv, DCNv2) and unit.conv.weight.shape[-1] == 1 and
                x.c % 64 == 0 and x.h % 2 == 0 and x.w % 2 == 0 and (x.h // 2) * (x.w // 2) >= 5000 and      # per image: batch-independent results
                getattr(self.model, 'fuse_avgpool', True))

    def _unit_avgpooled(self, name, unit, x):
        from model.custom_layers import ACT_CODES
        scale, shift = unit.folded_scale_shift()
        w = unit.conv.weight.detach().float()
        w2 = (w * 0.25).expand(-1, -1, 2, 2).contiguous()
        before = self.conv_flops
        out = self._conv(name, x, w2, scale, shift, 2, ACT_CODES[unit.act_name])
        true_flops = (self.conv_flops - before) // 4          # the algorithmic conv is the 1x1 on the pooled map
        self.conv_flops = before + true_flops
        info = self.step_info[name]
        info['flops'], info['k'] = true_flops, info['k'] // 4
        return out

    # ------------------------------------------------------------------ network walk
    def _block(self, name, blk, x, dst=None):
        from model.resnet_vd import ConvBlock, IdentityBlock, BasicBlock
        relu = _lib.ACT_RELU
        if isinstance(blk, ConvBlock):
            y = self._unit(name + '.conv1', blk.conv1, x)
            y = self._unit(name + '.conv2', blk.conv2, y)
            if self.code == PPY_F16X2 and not self.train_bn:
                # pair path: conv3 and the projection shortcut conv4 as one GEMM over [y ; shortcut input] (no shortcut tensor)
                sc_in = x if blk.is_first else self._avgpool(x)
                if self._dual_ok(blk.conv3, y, blk.conv4, sc_in):
                    out = self._conv_dual(name + '.conv3', blk.conv3, y, sc_in, blk.conv4, relu, dst)
                    if out is not None:
                        return out
                sc = self._unit(name + '.conv4', blk.conv4, sc_in)
            elif blk.is_first:
                sc = self._unit(name + '.conv4', blk.conv4, x)
            elif self._pool_into_conv_ok(blk.conv4, x):
                sc = self._unit_avgpooled(name + '.conv4', blk.conv4, x)
            else:
                sc = self._unit(name + '.conv4', blk.conv4, self._avgpool(x))
            return self._unit(name + '.conv3', blk.conv3, y, residual=sc, act=relu, dst=dst)
        if isinstance(blk, IdentityBlock):
            y = self._unit(name + '.conv1', blk.conv1, x)
            y = self._unit(name + '.conv2', blk.conv2, y)
            if self._dual_ok(blk.conv3, y, None, x):
                out = self._conv_dual(name + '.conv3', blk.conv3, y, x, None, relu, dst)
                if out is not None:
                    return out
            return self._unit(name + '.conv3', blk.conv3, y, residual=x, act=relu, dst=dst)
        if isinstance(blk, BasicBlock):
            sc = x
            if blk.conv3 is not None:
                sc = self._unit(name + '.conv3', blk.conv3, x if blk.is_first else self._avgpool(x))
            y = self._unit(name + '.conv1', blk.conv1, x)
            return self._unit(name + '.conv2', blk.conv2, y, residual=sc, act=relu, dst=dst)
        raise TypeError(type(blk))

    def _build(self):
        m, n = self.model, self.n
        bb, head = m.backbone, m.head
        n_out = len(head.anchor_masks)
        # static input: NCHW fp32 exactly as Decode.predict uploads it (model/decode_np.py:142-147)
        if self.input_u8:
            self.x_in = torch.zeros((n, self.h, self.w, 3), dtype=torch.uint8, device=self.dev)
            self.normalize = dict(getattr(m, 'normalize', None) or DEFAULT_NORMALIZE)
        else:
            self.x_in = torch.zeros((n, 3, self.h, self.w), dtype=torch.float32, device=self.dev)
        # per-image detection counts + one overflow flag of the pair path, read back together (the one host sync of a run)
        self._flags = torch.zeros((n + 1,), dtype=torch.int32, device=self.dev)
        self.overflow = self._flags[n:]
        if self.code == PPY_F16X2:
            self._add('reset_flags', lambda: self.overflow.zero_())
        self._src = self.x_in                 # input slot the first kernel reads
        self.input_slots = [self.x_in]
        self.graphs = []
        self.im_size = torch.zeros((n, 2), dtype=torch.float32, device=self.dev)
        stem_units = list(zip(bb._stem_units(), ('conv1_1', 'conv1_2', 'conv1_3')))
        u0 = stem_units[0][0]
        stem_ok = tuple(u0.conv.weight.shape) == (32, 3, 3, 3) and u0.stride == 2 and u0.act_name in (None, 'relu', 'leaky')
        if stem_ok and not self.train_bn:
            # conv1_1 fused with the NCHW->NHWC change (K = 27: HBM-bound, fp32 SIMT, weights in the constant bank)
            x0 = self._stem(u0)
            stem_units = stem_units[1:]
        elif (stem_ok and self.code == PPY_BF16 and u0.bn is not None and u0.conv.bias is None and not self.input_u8 and
              not os.environ.get('PPY_NO_TRAIN_STEM')):
            # train-mode BatchNorm: the same kernel writes the RAW conv output (scale 1, shift 0, no activation) for the statistics
            # kernel -- instead of a layout pass + the generic gather-mode conv over 3 channels padded to 8
            from model.custom_layers import ACT_CODES
            x0 = self._unit_batch_stats('stem.conv1_1', u0, None, None, ACT_CODES[u0.act_name], None, False, raw=self._stem(u0, raw=True))
            stem_units = stem_units[1:]
        elif self.input_u8:
            raise NotImplementedError('uint8 input needs the fused 3 -> 32 stride-2 stem conv (both PP-YOLO backbones have it)')
        else:
            x0 = TensorRef(self._new(n, self.h, self.w, 8))
            args = (ctypes.c_void_p(x0.ptr), n, 3, self.h, self.w, 8, self.code)
            self._add('nchw_to_nhwc', lambda: check(lib.ppy_nchw_to_nhwc(ops.ptr(self._src), *args, ops.stream_ptr()), 'nchw_to_nhwc'))

        # head level i > 0 consumes cat([upsampled route, backbone feature]); give the backbone stage that
        # produces the feature a destination inside that concat buffer (no copy at run time)
        fmaps = list(bb.feature_maps)
        head_feats = fmaps[-1:-n_out - 1:-1]              # stages in head order (deepest first)
        route_c = [0] + [head.upsample_layers[2 * (i - 1)].filters for i in range(1, n_out)]
        stage_dst = {}
        self.concat = {}
        for i, stage in enumerate(head_feats):
            if i == 0 or self.backbone_only:
                continue
            side = self.h // (2 ** stage)
            feat_c = self._stage_channels(bb, stage)
            buf = self._new(n, side, side, route_c[i] + feat_c)
            self.concat[i] = TensorRef(buf)
            stage_dst[stage] = TensorRef(buf, feat_c, route_c[i])

        x = x0
        for u, nm in stem_units:
            # (pair path: these layers run the 32-element-K-block build of the kernel instead -- no structural zeros)
            pairs_ok = self.code == PPY_BF16 or (self.code == PPY_F16X2 and u.conv.weight.shape[0] < getattr(self.model, 'k32_min_cout', 64))
            pairable = (pairs_ok and not self.train_bn and not hasattr(u.conv, 'dcn_weight') and u.stride == 1 and
                        tuple(u.conv.weight.shape[1:]) == (32, 3, 3) and x.c == 32 and x.ld == 32 and x.c_off == 0 and
                        x.w % 2 == 0 and u.conv.weight.shape[0] % 8 == 0 and u.conv.bias is None)
            x = self._unit_pixel_pairs('stem.' + nm, u, x) if pairable else self._unit('stem.' + nm, u, x)
        x = self._maxpool(x)
        feats = {}
        for stage in (2, 3, 4, 5):
            names = bb.stage_names(stage)
            for j, nm in enumerate(names):
                dst = stage_dst.get(stage) if j == len(names) - 1 else None
                x = self._block(nm, getattr(bb, nm), x, dst=dst)
            feats[stage] = x

        self.feats = [feats[stage] for stage in fmaps]
        if self.backbone_only:
            return
        # ---- head
        from model.custom_layers import Conv2dUnit, CoordConv, SPP, DropBlock
        self.head_outs = []
        route = None
        for i, stage in enumerate(head_feats):
            x = feats[stage] if i == 0 else self.concat[i]
            blk = head.detection_blocks[i]
            coord = False

            def walk(layers, x, prefix):
                nonlocal coord
                for j, ly in enumerate(layers):
                    if isinstance(ly, CoordConv):
                        coord = coord or ly.coord_conv
                    elif isinstance(ly, Conv2dUnit):
                        x = self._unit('%s.%d' % (prefix, j), ly, x, coord=coord)
                        coord = False
                    elif isinstance(ly, SPP):
                        x = self._spp(x, ly.seq)
                    elif isinstance(ly, DropBlock):
                        if not ly.is_test:
                            raise RuntimeError('DropBlock must be in test mode for inference (head.set_dropblock(True))')
                    else:
                        raise TypeError(type(ly))
                return x
            route = walk(blk.layers, x, 'head.block%d.layers' % i)
            tip = walk(blk.tip_layers, route, 'head.block%d.tip' % i)
            out = self._head_output_conv('head.out%d' % i, head.yolo_output_convs[i], tip)
            self.head_outs.append(out)
            if i < n_out - 1:
                nxt = self.concat[i + 1]
                self._unit('head.transition%d' % i, head.upsample_layers[2 * i], route, upsample=True,
                           dst=nxt.slice(0, route_c[i + 1]))

        # ---- decode + NMS
        an_per = [len(mk) for mk in head.anchor_masks]
        sizes = [o.h for o in self.head_outs]
        self.total_boxes = sum(s * s * a for s, a in zip(sizes, an_per))
        nc = head.num_classes
        self.boxes = torch.zeros((n, self.total_boxes, 4), dtype=torch.float32, device=self.dev)
        cfg = dict(head.nms_cfg)
        if cfg.pop('nms_type') != 'matrix_nms':
            raise NotImplementedError('only matrix_nms is on the PP-YOLO path')
        self.keep_top_k = cfg['keep_top_k']
        self.nms_out = torch.zeros((n, self.keep_top_k, 6), dtype=torch.float32, device=self.dev)
        self.nms_counts = self._flags[:n]
        nms_args = (cfg['score_threshold'], cfg['post_threshold'], cfg['nms_top_k'], cfg['keep_top_k'],
                    cfg.get('use_gaussian', False), cfg.get('gaussian_sigma', 2.0))
        # 'dense' (default): yolo_box's dense scores; the decode kernels also fill the per-image score histogram, so the
        # NMS front end needs one pass over the scores (collect) instead of two.  Cost independent of the data.
        # 'sparse': the decode kernels emit Matrix-NMS candidates (score > score_threshold) directly and skip the class
        # scores of anchors whose objectness is already below the threshold; no dense [boxes x C] score tensor.  Wins when
        # few scores pass the threshold (a trained detector), loses when most do (atomics per candidate).
        sparse = self.postprocess_impl == 'sparse' and cfg['score_threshold'] > 0
        if sparse:
            self.scores = None
            self.cand_cap = self.total_boxes * nc      # every score could pass: the list can never overflow (8 B/key of HBM)
            ws = self._keep(ops.nms_candidate_workspace(n, self.cand_cap, self.dev))
            self.cand_ws = ws
        else:
            self.scores = torch.zeros((n, self.total_boxes, nc), dtype=torch.float32, device=self.dev)
            ws = self._keep(torch.empty_like(ops.nms_workspace(n, self.total_boxes, nc, self.dev)))   # private: holds the histogram
        dense_hist = not sparse and cfg['score_threshold'] > 0
        off = 0
        for i, o in enumerate(self.head_outs):
            anchors = head._anchors[head.anchor_masks[i]].reshape(-1)

            def run(o=o, anchors=anchors, off=off, i=i):
                if sparse:
                    if i == 0:
                        ops.nms_candidates_reset(ws, n, self.cand_cap)
                    ops.yolo_decode_candidates_nhwc(o.t, o.ld, n, o.h, anchors, head.downsample[i], nc, head.scale_x_y,
                                                    self.im_size, head.clip_bbox, head.iou_aware, head.iou_aware_factor,
                                                    self.boxes, off, self.total_boxes, cfg['score_threshold'], ws,
                                                    self.cand_cap)
                else:
                    if i == 0 and dense_hist:
                        ops.nms_candidates_reset(ws, n, 1)
                    ops.yolo_decode_nhwc(o.t, o.ld, n, o.h, anchors, head.downsample[i], nc, head.scale_x_y, self.im_size,
                                         head.clip_bbox, head.iou_aware, head.iou_aware_factor, self.boxes, self.scores,
                                         off, self.total_boxes, hist_threshold=cfg['score_threshold'] if dense_hist else None,
                                         nms_workspace=ws if dense_hist else None)
            self._add('decode%d' % i, run)
            off += o.h * o.h * an_per[i]
        if sparse:
            self._add('matrix_nms', lambda: ops.matrix_nms_candidates_launch(self.boxes, nc, self.nms_out, self.nms_counts, ws,
                                                                             self.cand_cap, *nms_args))
        else:
            self._add('matrix_nms', lambda: ops.matrix_nms_launch(self.boxes, self.scores, self.nms_out, self.nms_counts, ws,
                                                                  *nms_args, have_hist=dense_hist))

    @staticmethod
    def _stage_channels(bb, stage):
        last = getattr(bb, bb.stage_names(stage)[-1])
        unit = last.conv3 if hasattr(last, 'conv4') or type(last).__name__ == 'IdentityBlock' else last.conv2
        return unit.filters

    # ------------------------------------------------------------------ execution
    def _run_steps(self):
        for _, fn in self.steps:
            fn()

    def _capture(self):
        # (train_bn engines: the warm-up pass below must not advance the BatchNorm running statistics)
        saved = [(bn, bn.running_mean.clone(), bn.running_var.clone()) for bn in self.bn_modules]
        try:
            return self._capture_graph()
        finally:
            for bn, mean, var in saved:
                bn.running_mean.copy_(mean)
                bn.running_var.copy_(var)

    def _capture_graph(self):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self._run_steps()                     # warm-up on the capture stream
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        with torch.cuda.graph(g, stream=s):
            self._run_steps()
        return g

    def capture_steps(self, keep):
        """CUDA graph of the plan steps whose name satisfies ``keep`` (same launch parameters and buffers as the full plan;
        used by bench.py to time one kernel family under the conditions of the real step: graph replay, programmatic
        dependent launch).  The buffers hold whatever the last full run left in them."""
        all_steps = self.steps
        self.steps = [(n, f) for n, f in all_steps if keep(n)]
        try:
            return self._capture()
        finally:
            self.steps = all_steps

    def add_input_slot(self):
        """A second (third, ...) static input buffer with its own captured graph, so a caller can upload batch i+1 straight
        into the engine while batch i is being computed -- no staging copy.  All other buffers are shared: launches of
        different slots must be ordered on one stream.  Returns the slot index."""
        t = torch.zeros_like(self.x_in)
        self.input_slots.append(t)
        if self.graph is not None:
            self._src = t
            self.graphs.append(self._capture())
            self._src = self.x_in
        return len(self.input_slots) - 1

    def launch(self, slot=0):
        """Enqueue one forward over the static input buffer `slot` on the current stream (no sync)."""
        if self.graph is not None:
            self.graphs[slot].replay()
        else:
            self._src = self.input_slots[slot]
            self._run_steps()
            self._src = self.x_in

    def run(self, x, im_size):
        """PPYOLO.forward(x, im_size) -> list of [M,6] tensors (reference model/ppyolo.py:19-22)."""
        if tuple(x.shape) != tuple(self.x_in.shape):
            raise ValueError('engine built for input %s, got %s' % (tuple(self.x_in.shape), tuple(x.shape)))
        self.x_in.copy_(x, non_blocking=True)
        self.im_size.copy_(im_size.reshape(self.n, 2), non_blocking=True)
        self.launch()
        flags = self._flags.cpu().tolist()               # the one host sync
        self.check_overflow(flags[self.n])
        out = self.nms_out.clone()
        return ops.split_predictions(out, flags[:self.n])

    def check_overflow(self, flag=None):
        flag = int(self.overflow.item()) if flag is None else flag
        if flag:
            raise _lib.KernelError("precision 'f16x2': an activation left the range of an fp16 pair (|v| * act_scale > 65504, "
                                   "or NaN); the results of this batch are invalid -- lower model.f16x2_act_scale or use "
                                   "model.precision = 'fp32'")

    def candidate_counts(self):
        """Scores above score_threshold per image in the last run (sparse post-processing only): int32 [n] tensor."""
        if self.scores is not None:
            raise RuntimeError('dense post-processing keeps no candidate list')
        off = 4 * 4096 * self.n                   # workspace layout: hist[n][4096] u32 | count[n] u32 | keys (nms_common.cuh)
        return self.cand_ws[off:off + 4 * self.n].view(torch.int32).clone()

    def head_outputs_nchw(self):
        """Raw head outputs of the last run as NCHW fp32 tensors (parity tests)."""
        return [ops.from_nhwc(o.t, o.c) for o in self.head_outs]

    def run_backbone(self, x, native=False):
        """Backbone forward only (``backbone_only`` engines): list of NCHW fp32 feature maps (fresh tensors).  ``native`` (bf16 /
        fp32 engines): logical NCHW VIEWS of the engine's own NHWC buffers instead (channels_last, the engine's dtype; valid until
        the next run) -- what the kernel head of the training step consumes without a layout or dtype round trip."""
        self.x_in.copy_(x, non_blocking=True)
        self.launch()
        if self.train_bn and self.bn_modules:
            torch._foreach_add_([bn.num_batches_tracked for bn in self.bn_modules], 1)      # one launch for all counters
        if native and not any(f.pair for f in self.feats):
            return [f.t[..., f.c_off:f.c_off + f.c].permute(0, 3, 1, 2) for f in self.feats]
        return self.feature_maps_nchw()

    def feature_maps_nchw(self):
        """Backbone feature maps of the last run as NCHW fp32 tensors (true values)."""
        outs = []
        for f in self.feats:
            if f.pair:
                t = ops.join_pair(f.t)[..., f.c_off:f.c_off + f.c] / self.act_scale
                outs.append(t.permute(0, 3, 1, 2).contiguous())
            else:
                outs.append(ops.from_nhwc(f.t[..., f.c_off:f.c_off + f.c].contiguous() if f.c_off else f.t, f.c))
        return outs

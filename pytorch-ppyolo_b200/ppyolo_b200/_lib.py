"""ctypes binding of libppyolo_b200.so (the C ABI declared in include/ppyolo_b200.h).

The library must exist: there is no Python/torch fallback for any entry point.  If the shared object is
missing and nvcc is available it is built in-tree once; otherwise import fails loudly.
"""
import ctypes
import os

from . import build as _build

c_int, c_ll, c_float, c_double, c_void_p, c_size_t = (ctypes.c_int, ctypes.c_longlong, ctypes.c_float,
                                                      ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t)

PPY_F32, PPY_BF16, PPY_F16X2 = 0, 1, 2
ABI_VERSION = 10
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_MISH = 0, 1, 2, 3


class ConvParams(ctypes.Structure):
    """Mirror of ``struct ppy_conv_params``."""
    _fields_ = [
        ('x', c_void_p), ('x_ld', c_int),
        ('n', c_int), ('h', c_int), ('w', c_int), ('cin', c_int),
        ('weight', c_void_p),
        ('cout', c_int), ('kh', c_int), ('kw', c_int), ('stride', c_int), ('pad', c_int),
        ('k_pad', c_int), ('cout_pad', c_int),
        ('scale', c_void_p), ('shift', c_void_p), ('bias_map', c_void_p),
        ('residual', c_void_p), ('res_ld', c_int),
        ('act', c_int),
        ('y', c_void_p), ('y_ld', c_int), ('out_dtype', c_int), ('upsample2x', c_int),
        ('offset_mask', c_void_p), ('om_ld', c_int),
        ('accumulate', c_int), ('split_k', c_int), ('wgrad_taps', c_int), ('wgrad_pitch', c_int), ('wgrad_tap_stride', c_int),
        ('coord_w', c_void_p),
        ('x_plane', c_ll), ('y_plane', c_ll), ('res_plane', c_ll), ('overflow', c_void_p),
        ('x2', c_void_p), ('x2_ld', c_int), ('x2_plane', c_ll), ('x2_kb', c_int), ('x2_tiled', c_int), ('x2_row_mod', c_int), ('x2_rows', c_int),
    ]


SIGNATURES = {
    'ppy_abi_version': (c_int, []),
    'ppy_status_string': (ctypes.c_char_p, [c_int]),
    'ppy_last_cuda_error': (c_int, []),
    'ppy_kernel_launch_count': (c_ll, []),
    'ppy_nchw_to_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_nhwc_to_nchw': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_maxpool3x3s2': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_avgpool2x2': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_spp': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_spp_backward': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_upsample2x': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_copy_channels': (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p]),
    'ppy_coord_channels': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_activation': (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p]),
    'ppy_pack_conv_weight_dgrad': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_pack_conv_weight': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                     c_int, c_void_p]),
    'ppy_stem_conv3x3s2': (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float),
                                   ctypes.POINTER(c_float), c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    'ppy_dcn_gather': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                               c_int, c_void_p]),
    'ppy_conv_f32': (c_int, [ctypes.POINTER(ConvParams), c_void_p]),
    'ppy_conv_bf16': (c_int, [ctypes.POINTER(ConvParams), c_void_p]),
    'ppy_conv_bf16_supported': (c_int, []),
    'ppy_conv_f16x2': (c_int, [ctypes.POINTER(ConvParams), c_void_p]),
    'ppy_stem_conv3x3s2_f16x2': (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float),
                                         ctypes.POINTER(c_float), c_int, c_int, c_void_p, c_int, c_ll, c_void_p]),
    'ppy_stem_conv3x3s2_u8': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(c_float),
                                      ctypes.POINTER(c_float), c_int, c_int, c_void_p, c_int, c_int, c_ll, c_void_p]),
    'ppy_maxpool3x3s2_f16x2': (c_int, [c_void_p, c_int, c_ll, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_avgpool2x2_f16x2': (c_int, [c_void_p, c_int, c_ll, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_spp_f16x2': (c_int, [c_void_p, c_int, c_ll, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    'ppy_split_f16x2': (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_ll, c_int, c_void_p]),
    'ppy_join_f16x2': (c_int, [c_void_p, c_int, c_ll, c_void_p, c_int, c_ll, c_int, c_void_p]),
    'ppy_bn_batch_stats': (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ppy_bn_train_fused': (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ppy_bn_act_backward': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ppy_scale_shift_act': (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                    c_int, c_void_p]),
    'ppy_dropblock_mask': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_int, c_float, c_void_p,
                                   c_void_p, c_void_p]),
    'ppy_dropblock_mask_from_seeds': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_void_p, c_void_p,
                                              c_void_p]),
    'ppy_dropblock_apply': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    'ppy_sgd_momentum': (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_int, c_void_p]),
    'ppy_sgd_ema_multi': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                                  c_float, c_int, c_float, c_float, c_void_p]),
    'ppy_allreduce_sgd_ema': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_uint, c_void_p, c_ll,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                                      c_float, c_int, c_float, c_float, c_void_p]),
    'ppy_ema_update': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p]),
    'ppy_im2col_kmajor': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    'ppy_im2col_kmajor_strided': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    'ppy_dcn_backward_sample': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_int, c_void_p, c_int, c_void_p]),
    'ppy_iou_aware_score': (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_double, c_void_p]),
    'ppy_yolo_decode': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, c_double,
                                c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'ppy_yolo_decode_hist': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, c_double,
                                     c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p,
                                     c_void_p]),
    'ppy_matrix_nms_batched_hist': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                                            c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ppy_pairwise_iou': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'ppy_resize_cubic_u8_batch': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    'ppy_yolo_loss_workspace_bytes': (c_int, [c_int, c_int, c_int]),
    'ppy_yolo_loss_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int,
                                      ctypes.c_double, c_float, c_int, c_int, c_float, c_int, c_float, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    'ppy_yolo_loss_backward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, ctypes.c_double,
                                       c_int, c_int, c_float, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ppy_gt2yolo_target': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_int),
                                   c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    'ppy_matrix_nms_workspace_bytes': (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    'ppy_nms_candidate_workspace_bytes': (c_int, [c_int, c_int, ctypes.POINTER(c_size_t)]),
    'ppy_nms_candidates_reset': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'ppy_yolo_decode_candidates': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, c_double,
                                           c_void_p, c_int, c_int, c_double, c_void_p, c_int, c_int, c_float, c_void_p, c_int,
                                           c_void_p]),
    'ppy_matrix_nms_candidates': (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_float,
                                          c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'ppy_matrix_nms_batched': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                                       c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}


def _load():
    path = _build.LIB_PATH
    # rebuild when a source is newer than the library (a stale .so would read past an older struct layout); a box without nvcc
    # (sources untouched since the build) never gets here
    if _build.needs_build() and (not os.path.exists(path) or _build.have_nvcc()):
        try:
            _build.build()
        except Exception as exc:  # no nvcc / compile error: fail loudly, never fall back
            raise ImportError('libppyolo_b200.so is missing and could not be built (%s). Run '
                              '`python -c "import __graft_entry__ as g; g.build()"` at the repo root.' % exc)
    lib = ctypes.CDLL(path)
    lib.ppy_abi_version.restype = c_int
    if lib.ppy_abi_version() != ABI_VERSION:
        raise ImportError('%s has ABI version %d, this package needs %d: rebuild it (python -c "import __graft_entry__ as g; '
                          'g.build()")' % (path, lib.ppy_abi_version(), ABI_VERSION))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


class _LazyLibrary(object):
    """Loads libppyolo_b200.so on the first use of an entry point (so building a model's parameter containers -- e.g. for the
    CPU reference arm of bench.py -- maps no native code); every later attribute access goes straight to the CDLL."""
    _real = None

    def _get(self):
        real = object.__getattribute__(self, '_real')
        if real is None:
            real = _load()
            _LazyLibrary._real = real
        return real

    def __getattr__(self, name):
        fn = getattr(self._get(), name)
        object.__setattr__(self, name, fn)          # next access bypasses __getattr__
        return fn


lib = _LazyLibrary()
LIB_PATH = _build.LIB_PATH


def is_loaded():
    return _LazyLibrary._real is not None


class KernelError(RuntimeError):
    pass


def check(status, what=''):
    if status != 0:
        msg = lib.ppy_status_string(status).decode()
        raise KernelError('%s failed: %s (status %d, cudaError %d)' % (what or 'ppyolo_b200 call', msg, status,
                                                                      lib.ppy_last_cuda_error()))


def launch_count():
    return int(lib.ppy_kernel_launch_count())

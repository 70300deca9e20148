"""CPU checks of the C-ABI boundary: the library loads and exports every symbol include/ppyolo_b200.h declares."""
import ctypes
import os
import re

from tests.conftest import REPO


def declared_symbols():
    text = open(os.path.join(REPO, 'include', 'ppyolo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ppy_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported():
    from ppyolo_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), 'missing export ' + name
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_status_strings_and_version():
    from ppyolo_b200 import _lib
    assert _lib.lib.ppy_abi_version() == _lib.ABI_VERSION
    assert _lib.lib.ppy_status_string(0) == b'ok'
    assert b'workspace' in _lib.lib.ppy_status_string(-2)
    assert _lib.launch_count() >= 0


def test_conv_params_struct_matches_header():
    """Field order of the ctypes mirror follows the header's struct declaration."""
    from ppyolo_b200 import _lib
    text = open(os.path.join(REPO, 'include', 'ppyolo_b200.h')).read()
    body = text[text.index('typedef struct ppy_conv_params {'):text.index('} ppy_conv_params;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split('{', 1)[1].split(';'):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace('*', ' ').split()
        tail = decl.split(',')
        first = tail[0].replace('*', ' ').split()[-1]
        fields.append(first)
        fields += [t.strip().lstrip('*') for t in tail[1:]]
    assert fields == [f[0] for f in _lib.ConvParams._fields_]


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from ppyolo_b200 import ops
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.matrix_nms_batched(torch.zeros(1, 4, 4), torch.zeros(1, 4, 80), 0.01, 0.01, 500, 100)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.conv_bn_act(torch.zeros(1, 8, 4, 4), torch.zeros(8, 8, 1, 1), torch.ones(8), torch.zeros(8))

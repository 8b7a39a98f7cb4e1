"""The C-ABI library loads and exports every symbol include/mate_b200.h declares (CPU: no
compute calls)."""

import ctypes
import os
import re

import pytest

from mate_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mate_b200.h'), encoding='UTF-8').read()
    return sorted(set(re.findall(r'\b(mate_b200_[a-z_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__ as entry

        entry.build()
    return _abi.load_library()


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.mate_b200_abi_version() == 2


def test_struct_layouts_match_the_header(lib):
    # sizes computed from the header's field lists (LP64)
    assert ctypes.sizeof(_abi.MateConfig) == 10 * 4 + 11 * 8 + 3 * 8
    assert ctypes.sizeof(_abi.MateStateView) == 16 * 8
    assert ctypes.sizeof(_abi.MateStepAux) == 14 * 8
    assert ctypes.sizeof(_abi.MateReplay) == 2 * 8


def test_create_without_gpu_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mate_b200.config import flatten_config, read_config

    cfg = _abi.make_config_struct(flatten_config(read_config('MATE-4v8-9.yaml')))
    handle = ctypes.c_void_p()
    rc = lib.mate_b200_create(ctypes.byref(cfg), 16, 0, 0, ctypes.byref(handle))
    assert rc != 0 and lib.mate_b200_last_error()
    from mate_b200.sim import BatchedSim

    with pytest.raises(RuntimeError):
        BatchedSim(flatten_config(read_config('MATE-4v8-9.yaml')), 16)

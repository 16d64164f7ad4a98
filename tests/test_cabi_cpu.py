"""CPU checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, validates configs, and refuses to run without a B200 (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from tf_face_toolbox_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_every_header_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "asoftmax_b200.h")).read()
    declared = set(re.findall(r"\b(asm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)


def test_version_and_lambda_schedule(lib):
    assert b"sm_100a" in lib.asm_version()
    assert lib.asm_lambda(1, 1000.0, 0.12, 1.0, 5.0) == pytest.approx(1000 / 1.12, rel=1e-6)
    assert lib.asm_lambda(10 ** 7, 1000.0, 0.12, 1.0, 5.0) == 5.0


def test_workspace_bytes_and_config_validation(lib):
    ok = _lib.AsmConfig(512, 85742, 85742, 0, 512, 4, _lib.MODE_BF16, 0, 1, None)
    n = lib.asm_workspace_bytes(C.byref(ok))
    # bf16 W copy + bf16 G'' dominate: 2 * 512 * 85760 * 2 bytes
    assert 2 * 512 * 85760 * 2 < n < 300e6
    for bad in [
        _lib.AsmConfig(500, 100, 100, 0, 32, 4, 0, 0, 1, None),       # D % 16
        _lib.AsmConfig(512, 100, 100, 0, 32, 5, 0, 0, 1, None),       # m
        _lib.AsmConfig(512, 100, 60, 50, 32, 4, 0, 0, 1, None),       # shard beyond C_total
        _lib.AsmConfig(48, 100, 100, 0, 32, 4, 1, 0, 1, None),        # bf16 needs D % 64
        _lib.AsmConfig(512, 100, 100, 0, 32, 4, 2, 0, 1, None),       # mode
    ]:
        assert lib.asm_workspace_bytes(C.byref(bad)) == 0
        h = C.c_void_p()
        assert lib.asm_create(C.byref(h), C.byref(bad)) == _lib.ASM_ERR_INVALID_ARG


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    cfg = _lib.AsmConfig(64, 100, 100, 0, 32, 4, 0, 0, 1, None)
    h = C.c_void_p()
    assert lib.asm_create(C.byref(h), C.byref(cfg)) == _lib.ASM_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.asm_last_error(None)
    from tf_face_toolbox_b200 import asoftmax_head
    X, W, y = torch.randn(4, 64), torch.randn(64, 100), torch.zeros(4, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        asoftmax_head(X, y, 100, 4, 5.0, weights=W, mode="fp32")


def test_new_entry_points_validate_arguments_without_a_gpu(lib):
    """Host-side argument checks of the optimizer / center-loss / NVLink-transport entry points."""
    # symmetric-memory size: only meaningful for 2..8 ranks, grows with the global batch
    c2 = _lib.AsmConfig(512, 85742, 42871, 0, 512, 4, _lib.MODE_BF16, 0, 2, None)
    c8 = _lib.AsmConfig(512, 85742, 10718, 0, 512, 4, _lib.MODE_BF16, 0, 8, None)
    c1 = _lib.AsmConfig(512, 85742, 85742, 0, 512, 4, _lib.MODE_BF16, 0, 1, None)
    n2, n8 = lib.asm_p2p_bytes(C.byref(c2)), lib.asm_p2p_bytes(C.byref(c8))
    assert lib.asm_p2p_bytes(C.byref(c1)) == 0
    # 2 x (dX [B, D] fp32 + stats) dominate: 2 * 512 * 512 * 4 bytes
    assert n2 > 2 * 512 * 512 * 4 and n8 > 2 * 512 * 512 * 4 and n2 > n8   # fewer local rows per rank at 8
    assert lib.asm_p2p_attach(None, None) == _lib.ASM_ERR_INVALID_ARG
    assert lib.asm_step_p2p(None, None, 0, None, 4, None, 0.0, None, None, None, None) == _lib.ASM_ERR_INVALID_ARG
    assert lib.asm_set_optimizer(None, None, None, None) == _lib.ASM_ERR_INVALID_ARG
    assert lib.asm_set_lambda_device(None, None) == _lib.ASM_ERR_INVALID_ARG
    assert lib.asm_center_loss(None, 4, 8, None, 4, None, 10, 0, 0.9, 1.0, None, None, None, None) == _lib.ASM_ERR_INVALID_ARG


def test_python_wrappers_refuse_cpu_tensors():
    from tf_face_toolbox_b200 import FusedOptimizer
    from tf_face_toolbox_b200.center import center_loss
    X, y, cen = torch.randn(4, 8), torch.zeros(4, dtype=torch.int32), torch.zeros(3, 8)
    with pytest.raises(RuntimeError):
        center_loss(X, y, cen)
    opt = FusedOptimizer("Adam", lr=1e-3)
    assert opt.kind == _lib.OPT_ADAM and opt.beta1 == 0.5 and opt.beta2 == 0.999   # data_parallel.py:193
    assert FusedOptimizer("Momentum").momentum == 0.9                               # data_parallel.py:191

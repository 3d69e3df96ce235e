"""The C-ABI library loads and exports every symbol include/vtts_b200.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "vtts_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(vtts_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("vtts_lr_rowsum", "vtts_lr_gather", "vtts_gen_create", "vtts_gen_load_layer",
                 "vtts_gen_workspace_bytes", "vtts_gen_forward", "vtts_gen_destroy", "vtts_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from vtts_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "run python viet-transformer-tts_b200/build.py"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_python_binding_covers_the_header():
    from vtts_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.vtts_version() == 100


def test_config_struct_size_matches_header():
    from vtts_b200 import _lib

    # 6 + 4*8 + 1 + 8 + 8 + 64 + 1 ints + 2 floats
    assert ctypes.sizeof(_lib.VttsGenConfig) == 4 * (6 + 32 + 1 + 8 + 8 + 64 + 1 + 2)
    assert ctypes.sizeof(_lib.VttsLayerInfo) == 36


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    from vtts_b200 import _lib

    lib = _lib.load()
    cfg = _lib.VttsGenConfig()
    out = ctypes.c_void_p()
    assert lib.vtts_gen_create(ctypes.byref(cfg), ctypes.byref(out)) == -1
    assert b"channels" in lib.vtts_last_error()
    cfg.in_channels, cfg.out_channels, cfg.channels, cfg.kernel_size = 80, 1, 512, 6
    assert lib.vtts_gen_create(ctypes.byref(cfg), ctypes.byref(out)) == -1
    assert b"odd" in lib.vtts_last_error()
    with pytest.raises(_lib.VttsError):
        _lib.check(lib.vtts_lr_gather(0, 0, 0, 1, 1, 1, 1, 3, 0, 0))

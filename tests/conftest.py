"""Shared test plumbing.

Markers: ``gpu`` = needs a B200 (run with ``-m gpu`` on the GPU box); everything else runs on
the CPU-only build container (``-m "not gpu"``).  The oracle (``oracle/``) is imported here and
only here / in tests: it is the checker, never the product path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "viet-transformer-tts_b200")
for p in (PKG_ROOT, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def split_cases(flat: dict) -> dict:
    cases: dict = {}
    for k, v in flat.items():
        case, field = k.split(".", 1)
        cases.setdefault(case, {})[field] = v
    return cases


def state_dict_from(flat: dict, prefix: str = "sd."):
    import torch

    return {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in flat.items() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def lr_c_oracle():
    """ctypes handle of the plain-C LengthRegulator restatement (oracle/lr_oracle.c)."""
    so = os.path.join(ROOT, "oracle", "_build", "liblr_oracle.so")
    src = os.path.join(ROOT, "oracle", "lr_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    lib.lr_oracle_scale.argtypes = [vp, i64, ctypes.c_float, vp]
    lib.lr_oracle_scale.restype = None
    lib.lr_oracle_fix_all_zero.argtypes = [vp, i64, i64]
    lib.lr_oracle_fix_all_zero.restype = ctypes.c_int
    lib.lr_oracle_rowsum.argtypes = [vp, i64, i64, vp]
    lib.lr_oracle_rowsum.restype = i64
    lib.lr_oracle_expand.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, vp]
    lib.lr_oracle_expand.restype = None
    return lib


def c_oracle_lr(lib, xs: np.ndarray, ds: np.ndarray, alpha: float = 1.0, pad: float = 0.0):
    """Run the C restatement end to end; returns (out, ds_after_fixup, mel_len)."""
    xs = np.ascontiguousarray(xs)
    ds = np.ascontiguousarray(ds.astype(np.int64)).copy()
    B, Tmax = ds.shape
    D = xs.shape[2]
    if alpha != 1.0:
        out = np.empty_like(ds)
        lib.lr_oracle_scale(ds.ctypes.data, ds.size, ctypes.c_float(alpha), out.ctypes.data)
        ds = out
    lib.lr_oracle_fix_all_zero(ds.ctypes.data, B, Tmax)
    mel_len = np.zeros(B, dtype=np.int64)
    t_out = lib.lr_oracle_rowsum(ds.ctypes.data, B, Tmax, mel_len.ctypes.data)
    assert t_out >= 0
    out = np.empty((B, t_out, D), dtype=xs.dtype)
    padv = np.array([pad], dtype=xs.dtype)
    lib.lr_oracle_expand(xs.ctypes.data, ds.ctypes.data, out.ctypes.data, B, Tmax, D, t_out, xs.dtype.itemsize,
                         padv.ctypes.data)
    return out, ds, mel_len


def rel_l2(a, b) -> float:
    import torch

    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())

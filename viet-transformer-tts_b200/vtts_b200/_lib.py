"""ctypes binding of libvtts_b200.so (see include/vtts_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets a RuntimeError.  Build it with ``python viet-transformer-tts_b200/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VTTS_B200_LIB", os.path.join(_HERE, "libvtts_b200.so"))

MAX_STAGES, MAX_BLOCKS, MAX_DILATIONS = 8, 8, 8
PRECISION = {"fp32": 0, "bf16": 1, "fp16": 2}


class VttsGenConfig(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("channels", C.c_int32),
        ("global_channels", C.c_int32), ("kernel_size", C.c_int32), ("num_upsamples", C.c_int32),
        ("upsample_scales", C.c_int32 * MAX_STAGES), ("upsample_kernel_sizes", C.c_int32 * MAX_STAGES),
        ("upsample_paddings", C.c_int32 * MAX_STAGES), ("upsample_output_paddings", C.c_int32 * MAX_STAGES),
        ("num_blocks", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * MAX_BLOCKS),
        ("num_dilations", C.c_int32 * MAX_BLOCKS),
        ("resblock_dilations", (C.c_int32 * MAX_DILATIONS) * MAX_BLOCKS),
        ("use_additional_convs", C.c_int32), ("lrelu_slope", C.c_float), ("final_lrelu_slope", C.c_float),
    ]


class VttsLayerInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "cin", "cout", "ksize", "dilation", "stage", "block", "unit", "which")]


# name -> (restype, argtypes); every symbol include/vtts_b200.h declares
SIGNATURES = {
    "vtts_version": (C.c_int, []),
    "vtts_last_error": (C.c_char_p, []),
    "vtts_device_arch": (C.c_int, []),
    "vtts_lr_scale_durations": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "vtts_lr_rowsum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vtts_lr_fix_zero_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vtts_lr_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64,
                                 C.c_int, C.c_void_p, C.c_void_p]),
    "vtts_gauss_upsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "vtts_path_generate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vtts_path_expand": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p]),
    "vtts_gen_create": (C.c_int, [C.POINTER(VttsGenConfig), C.POINTER(C.c_void_p)]),
    "vtts_gen_destroy": (C.c_int, [C.c_void_p]),
    "vtts_gen_num_layers": (C.c_int, [C.c_void_p]),
    "vtts_gen_layer_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(VttsLayerInfo)]),
    "vtts_gen_load_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vtts_gen_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "vtts_gen_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vtts_gen_set_valid_lengths": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vtts_gen_last_launch_count": (C.c_int, [C.c_void_p]),
    "vtts_gen_set_range_probe": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vtts_dwconv_glu_swish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]),
    "vtts_conv_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vtts_conv_destroy": (None, [C.c_void_p]),
    "vtts_conv_padded_channels": (C.c_int, [C.c_void_p]),
    "vtts_conv_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vtts_conv_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_float, C.c_int, C.c_void_p]),
    "vtts_dbg_conv1d_fp32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "vtts_dbg_conv1d_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "vtts_dbg_resblock_chain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vtts_dbg_trace": (C.c_int, [C.c_int, C.c_void_p, C.c_int]),
    "vtts_dbg_umma_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vtts_dbg_umma_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library (once).  Raises RuntimeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"vtts_b200: CUDA library not found at {LIB_PATH}. There is no CPU or PyTorch fallback for "
                    "the synthesis path; build it with `python viet-transformer-tts_b200/build.py`."
                )
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


class VttsError(RuntimeError):
    pass


def check(rc: int) -> int:
    if rc < 0:
        msg = load().vtts_last_error()
        raise VttsError(f"vtts_b200 error {rc}: {msg.decode() if msg else '?'}")
    return rc


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, 0 for None."""
    return 0 if t is None else t.data_ptr()


def current_stream(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream

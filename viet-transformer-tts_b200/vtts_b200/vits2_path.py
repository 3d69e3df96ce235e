"""vits2 duration expansion backed by csrc/path.cu.

``generate_path(duration, mask)`` mirrors ``models/gan_tts/vits2/utils.py:111-126`` (same argument shapes and
result); ``expand_by_path(x, duration, mask)`` is the fused form of the two matmuls that follow it in
``VITS2.inference`` (``models/gan_tts/vits2/generator.py:251-259``):

    attn = generate_path(w_ceil, attn_mask)
    m_p  = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)   ==  expand_by_path(m_p, w_ceil, attn_mask)

without writing the (B, t_y, t_x) attention tensor.  Synthesis only (no backward); CUDA tensors only.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def _check(duration: torch.Tensor, what: str):
    if not duration.is_cuda:
        raise RuntimeError(f"vtts_b200.{what}: inputs must be CUDA tensors (no CPU fallback)")
    if duration.dim() != 3 or duration.size(1) != 1:
        raise ValueError(f"vtts_b200.{what}: duration must be [b, 1, t_x]")


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """duration [b, 1, t_x], mask [b, 1, t_y, t_x] -> path [b, 1, t_y, t_x] (dtype of ``mask``)."""
    lib = _lib.load()
    _check(duration, "generate_path")
    if mask.dim() != 4 or mask.size(1) != 1 or mask.size(0) != duration.size(0) or mask.size(3) != duration.size(2):
        raise ValueError("vtts_b200.generate_path: mask must be [b, 1, t_y, t_x]")
    b, _, t_y, t_x = mask.shape
    dev = duration.device
    with torch.cuda.device(dev):
        d = duration.detach().to(torch.float32).contiguous()
        m = mask.detach().to(dev, torch.float32).contiguous()
        path = torch.empty((b, 1, t_y, t_x), dtype=torch.float32, device=dev)
        _lib.check(lib.vtts_path_generate(d.data_ptr(), m.data_ptr(), path.data_ptr(), b, t_y, t_x, _lib.current_stream(dev)))
    return path if mask.dtype == torch.float32 else path.to(mask.dtype)


def expand_by_path(x: torch.Tensor, duration: torch.Tensor, mask: Optional[torch.Tensor] = None,
                   t_y: Optional[int] = None) -> torch.Tensor:
    """x [b, d, t_x], duration [b, 1, t_x], mask [b, 1, t_y, t_x] (or ``t_y``) -> [b, d, t_y].

    Equals ``torch.matmul(generate_path(duration, mask).squeeze(1), x.transpose(1, 2)).transpose(1, 2)``.
    """
    lib = _lib.load()
    _check(duration, "expand_by_path")
    if not x.is_cuda:
        raise RuntimeError("vtts_b200.expand_by_path: inputs must be CUDA tensors (no CPU fallback)")
    if x.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("vtts_b200.expand_by_path: synthesis only (backward is not implemented)")
    b, d, t_x = x.shape
    if duration.size(0) != b or duration.size(2) != t_x:
        raise ValueError("vtts_b200.expand_by_path: duration must be [b, 1, t_x]")
    if mask is not None:
        if mask.dim() != 4 or mask.size(0) != b or mask.size(3) != t_x:
            raise ValueError("vtts_b200.expand_by_path: mask must be [b, 1, t_y, t_x]")
        t_y = mask.size(2)
    if t_y is None:
        raise ValueError("vtts_b200.expand_by_path: give mask or t_y")
    dev = x.device
    with torch.cuda.device(dev):
        xk = x.detach().to(torch.float32).contiguous()
        dk = duration.detach().to(torch.float32).contiguous()
        mk = None if mask is None else mask.detach().to(dev, torch.float32).contiguous()
        out = torch.empty((b, d, int(t_y)), dtype=torch.float32, device=dev)
        _lib.check(lib.vtts_path_expand(xk.data_ptr(), dk.data_ptr(), _lib.ptr(mk), out.data_ptr(), b, d, int(t_y), t_x,
                                        _lib.current_stream(dev)))
    return out if x.dtype == torch.float32 else out.to(x.dtype)

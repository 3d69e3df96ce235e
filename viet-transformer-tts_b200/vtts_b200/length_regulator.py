"""Drop-in ``LengthRegulator`` backed by the sm_100a kernels in csrc/lr.cu.

Mirrors ``models/tts/fastspeech2/layers.py:410-462`` of the reference: same constructor,
same ``forward(xs, ds, alpha=1.0)`` contract, same quirks (the all-zero-batch fix-up mutates
the caller's ``ds`` in place and logs the same warning; torch.round half-to-even for alpha).
"""
from __future__ import annotations

import logging
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"vtts_b200.LengthRegulator: {what} is on {t.device}; this module only runs its CUDA kernels "
            "(no CPU fallback). Move the inputs to a B200."
        )


class _ExpandFn(torch.autograd.Function):
    """Frame expansion with a gradient w.r.t. ``xs`` (training, ``train.py`` with ``use_gaussian: false``).

    The reference's repeat_interleave + pad_list (layers.py:460-462) is differentiable w.r.t. ``xs``: frame t of row b is
    a copy of token ``j(b, t)``, so ``grad_xs[b, j] = sum of grad_out[b, t] over the frames of token j``.  Forward is the
    CUDA gather; backward is a scatter-add over the same frame->token map, rebuilt from the durations with torch ops
    (the training backward is not part of the synthesis hot path).
    """

    @staticmethod
    def forward(ctx, xs, ds_k, t_out, pad_value):
        lib = _lib.load()
        B, Tmax, D = xs.shape
        dev = xs.device
        xs_k = xs.detach()
        xs_k = xs_k if xs_k.is_contiguous() else xs_k.contiguous()
        out = torch.empty((B, t_out, D), dtype=xs.dtype, device=dev)
        if out.numel():
            pad = torch.full((1,), pad_value, dtype=xs.dtype)  # host element bytes
            _lib.check(lib.vtts_lr_gather(xs_k.data_ptr(), ds_k.data_ptr(), out.data_ptr(), B, Tmax, D, t_out,
                                          xs.element_size(), pad.data_ptr(), _lib.current_stream(dev)))
        ctx.save_for_backward(ds_k)
        ctx.shape = (B, Tmax, D, t_out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (ds_k,) = ctx.saved_tensors
        B, Tmax, D, t_out = ctx.shape
        cum = ds_k.cumsum(1)                                               # (B, Tmax) exclusive end frame of every token
        t = torch.arange(t_out, device=ds_k.device).expand(B, t_out).contiguous()
        tok = torch.searchsorted(cum, t, right=True)                       # frame -> token (== Tmax on the padding)
        valid = tok < Tmax
        g = grad_out * valid.unsqueeze(-1).to(grad_out.dtype)
        grad_xs = torch.zeros((B, Tmax, D), dtype=grad_out.dtype, device=grad_out.device)
        grad_xs.scatter_add_(1, tok.clamp(max=Tmax - 1).unsqueeze(-1).expand(B, t_out, D), g)
        return grad_xs, None, None, None


class LengthRegulator(nn.Module):
    """Expand token-level features to frame level by repeating each row ``ds[b, i]`` times."""

    def __init__(self, pad_value: float = 0.0):
        super().__init__()
        self.pad_value = pad_value

    # -- reference API -----------------------------------------------------------------------
    def forward(self, xs: torch.Tensor, ds: torch.LongTensor, alpha: float = 1.0) -> torch.Tensor:
        """(B, Tmax, D), (B, Tmax) int64 -> (B, max_b sum(ds[b]), D)   [layers.py:434-462]."""
        out, _ = self.forward_with_lengths(xs, ds, alpha)
        return out

    # -- extension: also hand back mel_len (the caller computes it at layers.py:209) ----------
    def forward_with_lengths(
        self, xs: torch.Tensor, ds: torch.LongTensor, alpha: float = 1.0, max_len: Optional[int] = None
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Like :meth:`forward` but also returns ``mel_len = ds.sum(1)`` (after alpha / fix-up).

        ``max_len`` (extension, default None = reference behaviour): a caller-supplied output
        length >= the true maximum skips the one 24-byte device->host read this call otherwise
        needs to size the output; frames beyond each row's length hold ``pad_value``.
        """
        lib = _lib.load()
        _require_cuda(xs, "xs")
        _require_cuda(ds, "ds")
        if xs.dim() != 3 or ds.dim() != 2 or xs.shape[:2] != ds.shape:
            raise ValueError(f"LengthRegulator: expected xs (B,Tmax,D) and ds (B,Tmax); got {tuple(xs.shape)}, {tuple(ds.shape)}")
        if ds.dtype != torch.int64:
            raise TypeError(f"LengthRegulator: ds must be int64 (LongTensor), got {ds.dtype}")
        B, Tmax, D = xs.shape
        if B == 0:
            raise ValueError("max() arg is an empty sequence")  # pad_list, function.py:118
        dev = xs.device
        with torch.cuda.device(dev):
            stream = _lib.current_stream(dev)
            if alpha != 1.0:
                assert alpha > 0
                src = ds if ds.is_contiguous() else ds.contiguous()
                ds = torch.empty_like(src)
                _lib.check(lib.vtts_lr_scale_durations(src.data_ptr(), src.numel(), float(alpha), ds.data_ptr(), stream))
            ds_k = ds if ds.is_contiguous() else ds.contiguous()
            mel_len = torch.empty(B, dtype=torch.int64, device=dev)
            stats = torch.empty(3, dtype=torch.int64, device=dev)
            _lib.check(lib.vtts_lr_rowsum(ds_k.data_ptr(), B, Tmax, mel_len.data_ptr(), stats.data_ptr(), stream))
            if max_len is None:
                t_max, total, n_neg = stats.tolist()  # the single host sync of this call
                if n_neg:
                    raise RuntimeError("repeats can not be negative")
                if total == 0:
                    logging.warning(
                        "predicted durations includes all 0 sequences. fill the first element with 1."
                    )
                    # layers.py:458 -- in place on the caller's tensor
                    _lib.check(lib.vtts_lr_fix_zero_rows(ds_k.data_ptr(), B, Tmax, stream))
                    if ds_k is not ds:
                        ds.copy_(ds_k)
                    _lib.check(lib.vtts_lr_rowsum(ds_k.data_ptr(), B, Tmax, mel_len.data_ptr(), stats.data_ptr(), stream))
                    t_max = Tmax
                t_out = int(t_max)
            else:
                t_out = int(max_len)
            if torch.is_grad_enabled() and xs.requires_grad:
                # training (the reference's expansion is differentiable w.r.t. xs): same kernel forward, scatter-add backward
                return _ExpandFn.apply(xs, ds_k, t_out, float(self.pad_value)), mel_len
            xs_k = xs if xs.is_contiguous() else xs.contiguous()
            out = torch.empty((B, t_out, D), dtype=xs.dtype, device=dev)
            if out.numel():
                pad = torch.full((1,), self.pad_value, dtype=xs.dtype)  # host element bytes
                _lib.check(lib.vtts_lr_gather(xs_k.data_ptr(), ds_k.data_ptr(), out.data_ptr(), B, Tmax, D, t_out,
                                              xs.element_size(), pad.data_ptr(), stream))
        return out, mel_len

"""Drop-in ``GaussianUpsampling`` backed by csrc/gauss.cu.

Mirrors ``models/tts/fastspeech2/layers.py:465-520`` (and the identical class in
``models/gan_tts/jets/alignments.py:168-222``): same constructor and ``forward(hs, ds, h_masks, d_masks)``,
including the in-place all-zero-batch fix-up of ``ds`` and the reference quirk that without ``h_masks`` the
number of output frames is the duration sum over the WHOLE batch.
"""
from __future__ import annotations

import logging
from typing import Optional

import torch
import torch.nn as nn

from . import _lib


class GaussianUpsampling(nn.Module):
    """Gaussian upsampling with fixed temperature (https://arxiv.org/abs/2010.04301)."""

    def __init__(self, delta: float = 0.1):
        super().__init__()
        self.delta = delta

    def forward(self, hs: torch.Tensor, ds: torch.Tensor, h_masks: Optional[torch.Tensor] = None,
                d_masks: Optional[torch.Tensor] = None) -> torch.Tensor:
        """hs (B,T_text,adim), ds (B,T_text), h_masks (B,T_feats) bool, d_masks (B,T_text) bool -> (B,T_feats,adim)."""
        lib = _lib.load()
        if not hs.is_cuda or not ds.is_cuda:
            raise RuntimeError("vtts_b200.GaussianUpsampling: inputs must be CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and (hs.requires_grad or (ds.is_floating_point() and ds.requires_grad)):
            # training (train.py with use_gaussian: true differentiates through hs AND the durations): the same policy as
            # the generator shells - run the reference's formula through PyTorch autograd; the kernel is the synthesis path
            return self._forward_eager(hs, ds, h_masks, d_masks)
        B, T_text, D = hs.shape
        dev = hs.device
        with torch.cuda.device(dev):
            stream = _lib.current_stream(dev)
            ds_k = ds if (ds.dtype == torch.int64 and ds.is_contiguous()) else ds.to(torch.int64).contiguous()
            need_total = h_masks is None
            stats = torch.empty(3, dtype=torch.int64, device=dev)
            _lib.check(lib.vtts_lr_rowsum(ds_k.data_ptr(), B, T_text, 0, stats.data_ptr(), stream))
            # `if ds.sum() == 0` (layers.py:492) needs the value on the host, exactly like the reference
            _, total, _ = stats.tolist()
            if total == 0:
                logging.warning(
                    "predicted durations includes all 0 sequences. fill the first element with 1."
                )
                _lib.check(lib.vtts_lr_fix_zero_rows(ds_k.data_ptr(), B, T_text, stream))
                if ds_k is not ds:
                    ds.copy_(ds_k.to(ds.dtype))
                total = B * T_text
            T_feats = int(total) if need_total else int(h_masks.size(-1))
            hs_k = hs.detach().to(torch.float32).contiguous()
            hm = None if h_masks is None else h_masks.to(dev, torch.bool).contiguous()
            dm = None if d_masks is None else d_masks.to(dev, torch.bool).contiguous()
            out = torch.empty((B, T_feats, D), dtype=torch.float32, device=dev)
            _lib.check(lib.vtts_gauss_upsample(hs_k.data_ptr(), ds_k.data_ptr(), _lib.ptr(hm), _lib.ptr(dm), out.data_ptr(),
                                               B, T_text, D, T_feats, float(self.delta), stream))
        return out if hs.dtype == torch.float32 else out.to(hs.dtype)

    def _forward_eager(self, hs, ds, h_masks=None, d_masks=None):
        """Autograd form of layers.py:476-520 (softmax over token centres, then a matmul)."""
        B = ds.size(0)
        if ds.sum() == 0:
            logging.warning(
                "predicted durations includes all 0 sequences. fill the first element with 1."
            )
            ds[ds.sum(dim=1).eq(0)] = 1
        T_feats = int(ds.sum()) if h_masks is None else h_masks.size(-1)
        t = torch.arange(0, T_feats, device=ds.device).unsqueeze(0).repeat(B, 1).float()
        if h_masks is not None:
            t = t * h_masks.float()
        c = ds.cumsum(dim=-1) - ds / 2
        energy = -1 * self.delta * (t.unsqueeze(-1) - c.unsqueeze(1)) ** 2
        if d_masks is not None:
            energy = energy.masked_fill(~(d_masks.unsqueeze(1).repeat(1, T_feats, 1)), -float("inf"))
        return torch.matmul(torch.softmax(energy, dim=2), hs)

"""vits2 / jik876-style skin over the same kernels: ``Generator``, ``ResBlock1``, ``ResBlock2``.

API mirror of models/gan_tts/vits2/layers.py:107-186 and sublayers.py:215-354.  Differences to
the ESPnet skin that matter for parity: ``conv_pre`` / ``conv_post``
carry no weight norm, ``conv_post`` has no bias, ``ups`` padding is ``(k - u) // 2``, the
conditioning conv is called ``cond`` and ``ResBlock2`` has no second conv.  The two skins are the
same function, so both map onto one ``VttsGen`` handle type.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import remove_weight_norm, weight_norm

from . import _lib
from .hifigan import _GeneratorBase

LRELU_SLOPE = 0.1


def init_weights(m, mean: float = 0.0, std: float = 0.01):
    """vits2/utils.py:8-11 (a no-op in effect on weight-normed modules, kept for RNG parity)."""
    if m.__class__.__name__.find("Conv") != -1:
        m.weight.data.normal_(mean, std)


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """vits2/utils.py:14-15."""
    return int((kernel_size * dilation - dilation) / 2)


def _wn_conv(channels: int, kernel_size: int, dilation: int) -> nn.Module:
    return weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=dilation,
                                 padding=get_padding(kernel_size, dilation)))


class ResBlock1(nn.Module):
    """sublayers.py:215-309 -- dilated conv + plain conv per unit, optional x_mask."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([_wn_conv(channels, kernel_size, d) for d in dilation])
        self.convs1.apply(init_weights)
        self.convs2 = nn.ModuleList([_wn_conv(channels, kernel_size, 1) for _ in dilation])
        self.convs2.apply(init_weights)

    def forward(self, x, x_mask=None):
        for c1, c2 in zip(self.convs1, self.convs2):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            if x_mask is not None:
                xt = xt * x_mask
            xt = F.leaky_relu(c1(xt), LRELU_SLOPE)
            if x_mask is not None:
                xt = xt * x_mask
            x = c2(xt) + x
        return x if x_mask is None else x * x_mask

    def remove_weight_norm(self):
        for layer in list(self.convs1) + list(self.convs2):
            remove_weight_norm(layer)


class ResBlock2(nn.Module):
    """sublayers.py:312-354 -- one dilated conv per unit."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.convs = nn.ModuleList([_wn_conv(channels, kernel_size, d) for d in dilation])
        self.convs.apply(init_weights)

    def forward(self, x, x_mask=None):
        for c in self.convs:
            xt = F.leaky_relu(x, LRELU_SLOPE)
            if x_mask is not None:
                xt = xt * x_mask
            x = c(xt) + x
        return x if x_mask is None else x * x_mask

    def remove_weight_norm(self):
        for layer in self.convs:
            remove_weight_norm(layer)


class Generator(_GeneratorBase):
    """vits2 decoder (layers.py:107-186): conv_pre, ups, resblocks, conv_post, cond."""

    def __init__(
        self,
        initial_channel: int,
        resblock: str = "1",
        resblock_kernel_sizes: List[int] = [3, 7, 11],
        resblock_dilation_sizes: List[List[int]] = [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        upsample_rates: List[int] = [8, 8, 2, 2],
        upsample_initial_channel: int = 512,
        upsample_kernel_sizes: List[int] = [16, 16, 4, 4],
        gin_channels: int = 0,
    ):
        super().__init__()
        self.upsample_factor = int(np.prod(upsample_rates))
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        self._args = dict(
            initial_channel=initial_channel, resblock=str(resblock), rks=list(resblock_kernel_sizes),
            rds=[list(d) for d in resblock_dilation_sizes], rates=list(upsample_rates),
            width=upsample_initial_channel, uks=list(upsample_kernel_sizes), gin=gin_channels,
        )
        self.conv_pre = nn.Conv1d(initial_channel, upsample_initial_channel, 7, 1, padding=3)
        block_cls = ResBlock1 if str(resblock) == "1" else ResBlock2

        self.ups = nn.ModuleList()
        width = upsample_initial_channel
        for u, k in zip(upsample_rates, upsample_kernel_sizes):
            self.ups.append(weight_norm(nn.ConvTranspose1d(width, width // 2, k, u, padding=(k - u) // 2)))
            width //= 2
        self.resblocks = nn.ModuleList()
        width = upsample_initial_channel
        for _ in range(len(self.ups)):
            width //= 2
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(block_cls(width, k, d))
        self.conv_post = nn.Conv1d(width, 1, 7, 1, padding=3, bias=False)
        self.ups.apply(init_weights)
        if gin_channels != 0:
            self.cond = nn.Conv1d(gin_channels, upsample_initial_channel, 1)
        self._init_runtime()

    def forward(self, x: torch.Tensor, g: Optional[torch.Tensor] = None):
        if self._needs_autograd(x, g):
            return self._forward_eager(x, g)
        return self._run_kernels(x, g)

    def remove_weight_norm(self):
        print("Removing weight norm...")
        for layer in self.ups:
            remove_weight_norm(layer)
        for blk in self.resblocks:
            blk.remove_weight_norm()
        self.__dict__.pop("_layer_cache", None)
        self.invalidate()

    # -- plumbing ----------------------------------------------------------------------------
    def _forward_eager(self, x, g=None):
        x = self.conv_pre(x)
        if g is not None:
            x = x + self.cond(g)
        for i in range(self.num_upsamples):
            x = self.ups[i](F.leaky_relu(x, LRELU_SLOPE))
            xs = None
            for j in range(self.num_kernels):
                y = self.resblocks[i * self.num_kernels + j](x)
                xs = y if xs is None else xs + y
            x = xs / self.num_kernels
        return torch.tanh(self.conv_post(F.leaky_relu(x)))

    def _layer_modules(self):
        mods = [self.conv_pre]
        for i in range(self.num_upsamples):
            mods.append(self.ups[i])
            for j in range(self.num_kernels):
                blk = self.resblocks[i * self.num_kernels + j]
                if isinstance(blk, ResBlock1):
                    for c1, c2 in zip(blk.convs1, blk.convs2):
                        mods += [c1, c2]
                else:
                    mods += list(blk.convs)
        mods.append(self.conv_post)
        if self._args["gin"] != 0:
            mods.append(self.cond)
        return mods

    def _gen_config(self) -> _lib.VttsGenConfig:
        a = self._args
        cfg = _lib.VttsGenConfig()
        cfg.in_channels, cfg.out_channels, cfg.channels = a["initial_channel"], 1, a["width"]
        cfg.global_channels = a["gin"] if a["gin"] else 0
        cfg.kernel_size = 7
        cfg.num_upsamples = len(a["rates"])
        for i, (u, k) in enumerate(zip(a["rates"], a["uks"])):
            cfg.upsample_scales[i], cfg.upsample_kernel_sizes[i] = u, k
            cfg.upsample_paddings[i], cfg.upsample_output_paddings[i] = (k - u) // 2, 0
        cfg.num_blocks = len(a["rks"])
        for j, (k, dil) in enumerate(zip(a["rks"], a["rds"])):
            cfg.resblock_kernel_sizes[j] = k
            cfg.num_dilations[j] = len(dil)
            for m, d in enumerate(dil):
                cfg.resblock_dilations[j][m] = d
        cfg.use_additional_convs = 1 if a["resblock"] == "1" else 0
        cfg.lrelu_slope = LRELU_SLOPE
        cfg.final_lrelu_slope = 0.01
        return cfg

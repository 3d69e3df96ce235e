"""Generator training on the kernels (the generator backward of ``hifigan_trainer.py:143-167``).

``conv1d_tc`` is a differentiable stride-1 "same" Conv1d whose three GEMM-shaped pieces run on tensor cores without
cuDNN:

  forward   y  = conv(a, W) + b          tcgen05 implicit GEMM (``vtts_conv_forward``), 16-bit operands, fp32 accumulate
  dgrad     dx = conv(dy, flip(W)^T)     the same kernel on the transposed, tap-reversed weights
  wgrad     dW[:, :, j] = dy^T a_shift_j one cuBLAS GEMM per tap over the (batch x time) axis, 16-bit in / fp32 out
  dbias     dy summed over batch and time

``conv_transpose1d_tc`` lowers ConvTranspose1d(kernel 2s, stride s) to its polyphase form -- a 3-tap ``conv1d_tc`` with
``s * Cout`` output rows on the input extended by one zero frame, then an interleaving reshape and the padding crop -- so
its backward is the autograd of those pieces.  ``hifigan_forward_tc`` runs a ``HiFiGAN`` / ``ResidualBlock`` module tree
through them with the module's own parameters (weight norm re-applied differentiably), activations and residual adds left
to PyTorch autograd.  Select it with ``HiFiGAN.train_backend = "tc"``; the default ``"eager"`` runs the module tree through
PyTorch exactly like the reference (fp32, cuDNN).  Operand rounding is the synthesis path's (fp16 by default: the incoming
gradient of every layer is scaled per tensor by a power of two before it is rounded, so small gradients do not fall into
the fp16 subnormals; ``precision = "bf16"`` needs no scaling).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

_HANDLES: Dict[Tuple[int, int, int, int, int], int] = {}     # (device, cin, cout, k, dilation) -> VttsConv*


def _handle(dev: torch.device, cin: int, cout: int, k: int, dil: int) -> int:
    lib = _lib.load()
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, cin, cout, k, dil)
    h = _HANDLES.get(key)
    if h is None:
        out = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.vtts_conv_create(cin, cout, k, dil, ctypes.byref(out)))
        h = out.value
        _HANDLES[key] = h
    return h


def _dtype(precision: str) -> torch.dtype:
    if precision not in ("fp16", "bf16"):
        raise ValueError("the tensor-core training path takes precision 'fp16' or 'bf16'")
    return torch.float16 if precision == "fp16" else torch.bfloat16


def _run_conv(a16: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], k: int, dil: int, precision: str) -> torch.Tensor:
    """a16 (B, L, cin) 16-bit channels-last, w (cout, cin, k) fp32 -> (B, L, cout) fp32 (weights packed on the fly: they
    change every optimiser step)."""
    lib = _lib.load()
    dev = a16.device
    B, L, cin = a16.shape
    cout = w.shape[0]
    with torch.cuda.device(dev):
        h = _handle(dev, cin, cout, k, dil)
        width = _lib.check(lib.vtts_conv_padded_channels(h))
        if width != cin:
            a16 = F.pad(a16, (0, width - cin))
        a16 = a16.contiguous()
        w = w.detach().to(torch.float32).contiguous()
        bb = None if b is None else b.detach().to(torch.float32).contiguous()
        st = _lib.current_stream(dev)
        _lib.check(lib.vtts_conv_load(h, w.data_ptr(), _lib.ptr(bb), st))
        out = torch.empty((B, L, cout), dtype=torch.float32, device=dev)
        _lib.check(lib.vtts_conv_forward(h, a16.data_ptr(), _lib.PRECISION[precision], B, L, None, out.data_ptr(), None, 1.0, 0, st))
    return out


class _Conv1dTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, dilation, precision):
        if not x.is_cuda:
            raise RuntimeError("vtts_b200.training: the tensor-core training path needs CUDA tensors (no CPU fallback)")
        cout, cin, k = weight.shape
        if k % 2 != 1:
            raise ValueError("conv1d_tc: odd kernel sizes only ('same' padding)")
        a16 = x.detach().transpose(1, 2).to(_dtype(precision)).contiguous()           # (B, L, cin) operand, kept for wgrad
        y = _run_conv(a16, weight, bias, k, dilation, precision)
        ctx.save_for_backward(a16, weight)
        ctx.dilation, ctx.precision, ctx.has_bias = dilation, precision, bias is not None
        return y.transpose(1, 2)                                                      # (B, cout, L), channels-last memory

    @staticmethod
    def backward(ctx, dy):
        a16, weight = ctx.saved_tensors
        cout, cin, k = weight.shape
        d, prec = ctx.dilation, ctx.precision
        dt = _dtype(prec)
        # fp16 operands: gradients are small and fp16 has 5 exponent bits - scale dy per tensor by a power of two so that
        # its largest element sits near 2^10 (exact, undone on the fp32 results; dx and dW are linear in dy)
        inv = None
        if dt == torch.float16:
            _, e = torch.frexp(dy.detach().abs().amax().clamp_min(1e-30))             # amax = m * 2^e, 0.5 <= m < 1
            sc = torch.ldexp(torch.ones_like(e, dtype=torch.float32), 10 - e)
            inv = torch.ldexp(torch.ones_like(e, dtype=torch.float32), e - 10)
            d16 = (dy * sc).transpose(1, 2).to(dt).contiguous()                       # (B, L, cout)
        else:
            d16 = dy.transpose(1, 2).to(dt).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            w_t = weight.detach().flip(2).transpose(0, 1).contiguous()                # (cin, cout, k): dgrad weights
            dx = _run_conv(d16, w_t, None, k, d, prec).transpose(1, 2)
            if inv is not None:
                dx = dx * inv
        if ctx.needs_input_grad[1]:
            half = (k - 1) // 2
            B, L, _ = a16.shape
            a_pad = F.pad(a16, (0, 0, half * d, half * d))                            # zero "same" padding along time
            dyt = d16.transpose(1, 2)                                                 # (B, cout, L)
            # one GEMM per layer: the k shifted views side by side (B, L, k * cin) against dy^T, fp32 accumulate over time
            # (one strided copy: window (t, j) of a_pad starts at row t + j * d; torch.cat of the k slices was 1/3 of the step)
            Lp, cw = a_pad.shape[1], a_pad.shape[2]
            a_unf = a_pad[:, : L, :] if k == 1 else a_pad.as_strided((B, L, k, cw), (Lp * cw, cw, d * cw, 1)).reshape(B, L, k * cw)
            dwf = torch.bmm(dyt, a_unf, out_dtype=torch.float32).sum(0)               # (cout, k * cin)
            dw = dwf.view(cout, k, -1)[:, :, :cin].permute(0, 2, 1)
            if inv is not None:
                dw = dw * inv
            dw = dw.to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum((0, 2))
        return dx, dw, db, None, None


def conv1d_tc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, dilation: int = 1,
              precision: str = "fp16") -> torch.Tensor:
    """Differentiable F.conv1d(x, weight, bias, padding=(k-1)//2*dilation, dilation=dilation) on the tensor-core kernels."""
    return _Conv1dTC.apply(x, weight, bias, int(dilation), precision)


def conv_transpose1d_tc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int, padding: int,
                        output_padding: int = 0, precision: str = "fp16") -> torch.Tensor:
    """Differentiable F.conv_transpose1d for kernel size 2 * stride (the HiFi-GAN upsamples, generator.py:79-95)."""
    cin, cout, K = weight.shape
    s = int(stride)
    if K != 2 * s or output_padding > padding:
        raise ValueError("conv_transpose1d_tc: kernel size must be 2 * stride")
    B, _, L = x.shape
    L_out = (L - 1) * s - 2 * padding + K + output_padding
    # y_full[u] with u = i0 * s + r:  W[:, :, r]^T x[i0] + W[:, :, r + s]^T x[i0 - 1]   -> 3 taps at offsets (-1, 0, +1)
    w0 = weight[:, :, :s].permute(2, 1, 0).reshape(s * cout, cin)                     # rows (r, co), tap offset 0
    w1 = weight[:, :, s:].permute(2, 1, 0).reshape(s * cout, cin)                     # tap offset -1
    w3 = torch.stack([w1, w0, torch.zeros_like(w0)], dim=2)                           # (s * cout, cin, 3)
    x_ext = F.pad(x, (0, 1))                                                          # one zero frame: i0 runs to L
    p = conv1d_tc(x_ext, w3, None, 1, precision)                                      # (B, s * cout, L + 1)
    y_full = p.transpose(1, 2).reshape(B, (L + 1) * s, cout)
    y = y_full[:, padding: padding + L_out, :].transpose(1, 2)
    if bias is not None:
        y = y + bias.view(1, -1, 1)
    return y


def _eff_weight(m: nn.Module) -> torch.Tensor:
    if hasattr(m, "weight_g"):
        return torch._weight_norm(m.weight_v, m.weight_g, 0)                          # differentiable w = g * v / ||v||
    return m.weight


def _conv(m: nn.Conv1d, x: torch.Tensor, precision: str) -> torch.Tensor:
    return conv1d_tc(x, _eff_weight(m), m.bias, m.dilation[0], precision)


def residual_block_forward_tc(blk: nn.Module, x: torch.Tensor, precision: str) -> torch.Tensor:
    """ResidualBlock.forward (hifigan/layers.py:83-98) with its convs on the kernels."""
    for i in range(len(blk.convs1)):
        xt = _conv(blk.convs1[i][1], blk.convs1[i][0](x), precision)
        if blk.use_additional_convs:
            xt = _conv(blk.convs2[i][1], blk.convs2[i][0](xt), precision)
        x = xt + x
    return x


def hifigan_forward_tc(gen: nn.Module, c: torch.Tensor, g: Optional[torch.Tensor] = None, precision: Optional[str] = None) -> torch.Tensor:
    """HiFiGAN.forward (generator.py:132-156) under autograd with every conv on the tensor-core kernels."""
    precision = precision or (gen.precision if gen.precision in ("fp16", "bf16") else "fp16")
    c = _conv(gen.input_conv, c, precision)
    if g is not None:
        gw = _eff_weight(gen.global_conv)                                             # (channels, gc, 1): a matmul
        c = c + (torch.einsum("oc,bcl->bol", gw[:, :, 0], g) + gen.global_conv.bias.view(1, -1, 1))
    for i in range(gen.num_upsamples):
        act, up = gen.upsamples[i][0], gen.upsamples[i][1]
        c = conv_transpose1d_tc(act(c), _eff_weight(up), up.bias, up.stride[0], up.padding[0], up.output_padding[0], precision)
        cs = 0.0
        for j in range(gen.num_blocks):
            cs = cs + residual_block_forward_tc(gen.blocks[i * gen.num_blocks + j], c, precision)
        c = cs / gen.num_blocks
    act, conv = gen.output_conv[0], gen.output_conv[1]
    c = _conv(conv, act(c), precision)
    for m in list(gen.output_conv)[2:]:
        c = m(c)
    return c

"""Acoustic decoder + Postnet between the LengthRegulator and the generator.

Module shells with the reference's constructor signatures, attribute names and ``state_dict()`` keys:

  ``PositionwiseFeedForward`` / ``MultiHeadAttention`` / ``FFTBlock`` / ``Decoder``
      models/tts/fastspeech2/blocks/transformer.py:90-298
  ``ConvNorm`` / ``Postnet``
      models/tts/fastspeech2/sublayers.py:70-101, models/tts/fastspeech2/layers.py:571-625
  ``AcousticTail``  decoder -> feats_linear -> postnet(outs) + outs, the tail of ``FastSpeech2.inference``
      (models/tts/fastspeech2/model.py:250-257)

Synthesis (no autograd, CUDA tensors): every Conv1d -- the FFT blocks' position-wise convolutions (k = 9 and 1,
256 -> 1024 -> 256) and the Postnet's five Conv1d(k = 5) with their eval-mode BatchNorm folded in -- runs on the tcgen05
implicit-GEMM kernel through ``vtts_conv_*`` on channels-last activations (the layout the transformer already uses, so
the reference's two transposes per block disappear; ReLU / tanh / the FFN residual are fused into the conv epilogues).
Attention, LayerNorm and the 256 -> 80 projection stay PyTorch (cuBLAS) ops this round.  With autograd (training) the
modules run the reference's formula through PyTorch like the generator shells do; there is no CPU fallback on the
synthesis path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

DEFAULT_PRECISION = "fp16"


def get_sinusoid_encoding_table(n_position: int, d_hid: int, padding_idx: Optional[int] = None) -> torch.Tensor:
    """blocks/utils.py:14-35 (float64 table, cast to fp32)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    idx = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (idx // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    if padding_idx is not None:
        table[padding_idx] = 0.0
    return torch.FloatTensor(table)


class _TcConv:
    """One ``vtts_conv`` handle per device for a Conv1d (+ optional BatchNorm1d to fold); re-packs when a parameter changes."""

    def __init__(self, conv: nn.Conv1d, bn: Optional[nn.BatchNorm1d] = None):
        if conv.stride != (1,) or conv.groups != 1 or conv.kernel_size[0] % 2 != 1:
            raise RuntimeError("vtts_b200: only stride-1, odd-kernel, ungrouped Conv1d runs on the conv kernel")
        if conv.padding != ((conv.kernel_size[0] - 1) // 2 * conv.dilation[0],):
            raise RuntimeError("vtts_b200: Conv1d must use 'same' zero padding")
        self.conv, self.bn = conv, bn
        self._handles: Dict[int, int] = {}
        self._sig: Dict[int, tuple] = {}

    def _signature(self):
        ts = [self.conv.weight, self.conv.bias]
        if self.bn is not None:
            ts += [self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        return tuple((t.data_ptr(), t._version) for t in ts if t is not None)

    def handle(self, dev: torch.device) -> int:
        lib = _lib.load()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        c = self.conv
        if h is None:
            out = ctypes.c_void_p()
            _lib.check(lib.vtts_conv_create(c.in_channels, c.out_channels, c.kernel_size[0], c.dilation[0], ctypes.byref(out)))
            h = out.value
            self._handles[idx] = h
        sig = self._signature()
        if self._sig.get(idx) != sig:
            w = c.weight.detach().to(dev, torch.float32)
            b = None if c.bias is None else c.bias.detach().to(dev, torch.float32)
            if self.bn is not None:       # eval-mode BatchNorm1d: y = (conv + b - mean) * gamma / sqrt(var + eps) + beta
                bn = self.bn
                s = (bn.weight.detach() if bn.weight is not None else 1.0) / torch.sqrt(bn.running_var.detach() + bn.eps)
                s = s.to(dev, torch.float32)
                w = w * s.view(-1, 1, 1)
                b0 = b if b is not None else torch.zeros(c.out_channels, device=dev)
                b = (b0 - bn.running_mean.detach().to(dev, torch.float32)) * s
                if bn.bias is not None:
                    b = b + bn.bias.detach().to(dev, torch.float32)
            w = w.contiguous()
            b = None if b is None else b.contiguous()
            _lib.check(lib.vtts_conv_load(h, w.data_ptr(), _lib.ptr(b), _lib.current_stream(dev)))
            torch.cuda.current_stream(dev).synchronize()      # w / b are temporaries
            self._sig[idx] = sig
        return h

    def padded_channels(self, dev) -> int:
        return _lib.check(_lib.load().vtts_conv_padded_channels(self.handle(dev)))

    def run(self, act16: torch.Tensor, precision: str, want_x: bool, want_a: bool, res: Optional[torch.Tensor] = None,
            slope_out: float = 1.0, act_tanh: bool = False):
        """act16 (B, L, padded_channels) 16-bit -> (out_x fp32 (B, L, cout) | None, out_a 16-bit (B, L, cout) | None)."""
        lib = _lib.load()
        dev = act16.device
        B, L, _ = act16.shape
        with torch.cuda.device(dev):
            h = self.handle(dev)
            dt = torch.float16 if precision == "fp16" else torch.bfloat16
            assert act16.dtype == dt and act16.is_contiguous()
            ox = torch.empty((B, L, self.conv.out_channels), dtype=torch.float32, device=dev) if want_x else None
            oa = torch.empty((B, L, self.conv.out_channels), dtype=dt, device=dev) if want_a else None
            _lib.check(lib.vtts_conv_forward(h, act16.data_ptr(), _lib.PRECISION[precision], B, L, _lib.ptr(res), _lib.ptr(ox),
                                             _lib.ptr(oa), float(slope_out), 1 if act_tanh else 0, _lib.current_stream(dev)))
        return ox, oa

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self._handles.values():
                lib.vtts_conv_destroy(h)
        except Exception:
            pass


def _kernel_path(x: torch.Tensor, module: nn.Module) -> bool:
    """True: run the CUDA kernels.  False: autograd is needed (training) -> PyTorch formula.  CPU inference raises."""
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        return False                         # same policy as the generator shells (hifigan.py:_needs_autograd)
    if not x.is_cuda:
        raise RuntimeError(f"vtts_b200.{type(module).__name__}: input is on {x.device}; the synthesis path only runs its "
                           "CUDA kernels (no CPU fallback)")
    return True


def _operand(x: torch.Tensor, precision: str, width: int) -> torch.Tensor:
    """(B, L, C) float -> contiguous 16-bit operand (B, L, width), zero in the padding channels."""
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    a = x.to(dt)
    if a.shape[-1] != width:
        a = F.pad(a, (0, width - a.shape[-1]))
    return a.contiguous()


class ScaledDotProductAttention(nn.Module):
    def __init__(self, temperature):
        super().__init__()
        self.temperature = temperature
        self.softmax = nn.Softmax(dim=2)

    def forward(self, q, k, v, mask=None):
        attn = torch.bmm(q, k.transpose(1, 2)) / self.temperature
        if mask is not None:
            attn = attn.masked_fill(mask, -np.inf)
        attn = self.softmax(attn)
        return torch.bmm(attn, v), attn


class MultiHeadAttention(nn.Module):
    """blocks/transformer.py:192-243 (PyTorch ops: three projections, batched attention, output projection, LayerNorm)."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        self.attention = ScaledDotProductAttention(temperature=np.power(d_k, 0.5))
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        self.dropout = nn.Dropout(dropout)
        self.precision = DEFAULT_PRECISION      # operand format of the fused synthesis-path attention ("fp32": exact)

    def forward(self, q, k, v, mask=None, need_weights: bool = True):
        d_k, d_v, n_head = self.d_k, self.d_v, self.n_head
        sz_b, len_q, _ = q.size()
        len_k, len_v = k.size(1), v.size(1)
        residual = q
        if (not need_weights and q is k and k is v and q.is_cuda and mask is not None
                and not (torch.is_grad_enabled() and (q.requires_grad or self.w_qs.weight.requires_grad))):
            # synthesis path of the drop-in Decoder: one projection GEMM for q, k, v and one fused attention kernel
            # (same arithmetic in fp32: softmax(q k^T / sqrt(d_k) with -inf on the masked keys) v; the attention matrix
            # itself is not materialised, so callers that want it use need_weights=True)
            w = torch.cat([self.w_qs.weight, self.w_ks.weight, self.w_vs.weight], 0)
            b = torch.cat([self.w_qs.bias, self.w_ks.bias, self.w_vs.bias], 0)
            qkv = F.linear(q, w, b)
            qh = qkv[..., : n_head * d_k].view(sz_b, len_q, n_head, d_k).transpose(1, 2)
            kh = qkv[..., n_head * d_k: 2 * n_head * d_k].view(sz_b, len_k, n_head, d_k).transpose(1, 2)
            vh = qkv[..., 2 * n_head * d_k:].view(sz_b, len_v, n_head, d_v).transpose(1, 2)
            keep = ~mask[:, :1, :].unsqueeze(1)                       # (B, 1, 1, len_k): the mask rows are identical (key padding)
            if self.precision != "fp32":
                # 16-bit operands, fp32 softmax / accumulation inside the fused kernel (tensor cores): the same operand
                # rounding the conv kernels apply; fp32 SDPA runs on the CUDA cores and took 0.49 ms per block at C2
                dt = torch.float16 if self.precision == "fp16" else torch.bfloat16
                output = F.scaled_dot_product_attention(qh.to(dt), kh.to(dt), vh.to(dt), attn_mask=keep).float()
            else:
                output = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=keep)
            output = output.transpose(1, 2).reshape(sz_b, len_q, n_head * d_v)
            output = self.dropout(self.fc(output))
            return self.layer_norm(output + residual), None
        q = self.w_qs(q).view(sz_b, len_q, n_head, d_k).permute(2, 0, 1, 3).contiguous().view(-1, len_q, d_k)
        k = self.w_ks(k).view(sz_b, len_k, n_head, d_k).permute(2, 0, 1, 3).contiguous().view(-1, len_k, d_k)
        v = self.w_vs(v).view(sz_b, len_v, n_head, d_v).permute(2, 0, 1, 3).contiguous().view(-1, len_v, d_v)
        mask = mask.repeat(n_head, 1, 1)
        output, attn = self.attention(q, k, v, mask=mask)
        output = output.view(n_head, sz_b, len_q, d_v).permute(1, 2, 0, 3).contiguous().view(sz_b, len_q, -1)
        output = self.dropout(self.fc(output))
        return self.layer_norm(output + residual), attn


class PositionwiseFeedForward(nn.Module):
    """blocks/transformer.py:265-298: ``layer_norm(w_2(relu(w_1(x))) + x)`` with Conv1d w_1 (k[0]) and w_2 (k[1])."""

    def __init__(self, d_in, d_hid, kernel_size, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, kernel_size=kernel_size[0], padding=(kernel_size[0] - 1) // 2)
        self.w_2 = nn.Conv1d(d_hid, d_in, kernel_size=kernel_size[1], padding=(kernel_size[1] - 1) // 2)
        self.layer_norm = nn.LayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)
        self.precision = DEFAULT_PRECISION
        self.__dict__["_tc"] = None

    def _convs(self):
        tc = self.__dict__.get("_tc")
        if tc is None or tc[0].conv is not self.w_1 or tc[1].conv is not self.w_2:
            tc = (_TcConv(self.w_1), _TcConv(self.w_2))
            self.__dict__["_tc"] = tc
        return tc

    def _forward_eager(self, x):
        output = self.w_2(F.relu(self.w_1(x.transpose(1, 2)))).transpose(1, 2)
        return self.layer_norm(self.dropout(output) + x)

    def forward(self, x):
        if not _kernel_path(x, self):
            return self._forward_eager(x)
        c1, c2 = self._convs()
        xf = x.detach().to(torch.float32).contiguous()
        a = _operand(xf, self.precision, c1.padded_channels(x.device))
        _, hid = c1.run(a, self.precision, want_x=False, want_a=True, slope_out=0.0)          # relu(w_1(x)), 16-bit
        if hid.shape[-1] != c2.padded_channels(x.device):
            hid = F.pad(hid, (0, c2.padded_channels(x.device) - hid.shape[-1])).contiguous()
        y, _ = c2.run(hid, self.precision, want_x=True, want_a=False, res=xf)                 # w_2(.) + x, fp32
        return self.layer_norm(y).to(x.dtype)


class FFTBlock(nn.Module):
    """blocks/transformer.py:169-189."""

    def __init__(self, d_model, n_head, d_k, d_v, d_inner, kernel_size, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, kernel_size, dropout=dropout)

    def forward(self, enc_input, mask=None, slf_attn_mask=None, need_weights: bool = True):
        enc_output, enc_slf_attn = self.slf_attn(enc_input, enc_input, enc_input, mask=slf_attn_mask, need_weights=need_weights)
        if mask is not None:
            enc_output = enc_output.masked_fill(mask.unsqueeze(-1), 0)
        enc_output = self.pos_ffn(enc_output)
        if mask is not None:
            enc_output = enc_output.masked_fill(mask.unsqueeze(-1), 0)
        return enc_output, enc_slf_attn


class Decoder(nn.Module):
    """blocks/transformer.py:90-166: positional encoding + ``layers`` FFT blocks over the expanded frames."""

    def __init__(self, layers: int, hidden_dim: int, max_seq_len: int, config: Dict) -> None:
        super().__init__()
        self.config = config
        n_head = config["decoder_head"]
        d_k = d_v = hidden_dim // n_head
        self.max_seq_len = max_seq_len
        self.d_model = hidden_dim
        self.position_enc = nn.Parameter(get_sinusoid_encoding_table(max_seq_len + 1, hidden_dim).unsqueeze(0), requires_grad=False)
        self.layer_stack = nn.ModuleList([
            FFTBlock(hidden_dim, n_head, d_k, d_v, config["conv_filter_size"], config["conv_kernel_size"],
                     dropout=config["decoder_dropout"]) for _ in range(layers)])

    def forward(self, enc_seq, mask, return_attns=False):
        batch_size, max_len = enc_seq.shape[0], enc_seq.shape[1]
        if not self.training and enc_seq.shape[1] > self.max_seq_len:
            slf_attn_mask = mask.unsqueeze(1).expand(-1, max_len, -1)
            dec_output = enc_seq + get_sinusoid_encoding_table(enc_seq.shape[1], self.d_model)[: enc_seq.shape[1], :] \
                .unsqueeze(0).expand(batch_size, -1, -1).to(enc_seq.device)
        else:
            max_len = min(max_len, self.max_seq_len)
            slf_attn_mask = mask.unsqueeze(1).expand(-1, max_len, -1)
            dec_output = enc_seq[:, :max_len, :] + self.position_enc[:, :max_len, :].expand(batch_size, -1, -1)
            mask = mask[:, :max_len]
            slf_attn_mask = slf_attn_mask[:, :, :max_len]
        for dec_layer in self.layer_stack:
            dec_output, _ = dec_layer(dec_output, mask=mask, slf_attn_mask=slf_attn_mask, need_weights=return_attns)
        return dec_output, mask


class ConvNorm(nn.Module):
    """sublayers.py:70-101."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(w_init_gain))

    def forward(self, signal):
        return self.conv(signal)


class Postnet(nn.Module):
    """layers.py:571-625: five Conv1d(k=5) + BatchNorm1d, tanh after the first four; (B, T, n_mel) in and out."""

    def __init__(self, n_channels: int, config: Dict) -> None:
        super().__init__()
        self.conf = config
        emb, k, n = config["embedding_dim"], config["kernel_size"], config["conv_layers"]
        pad = int((k - 1) / 2)
        self.convolutions = nn.ModuleList()
        self.convolutions.append(nn.Sequential(ConvNorm(n_channels, emb, kernel_size=k, stride=1, padding=pad, dilation=1,
                                                        w_init_gain="tanh"), nn.BatchNorm1d(emb)))
        for _ in range(1, n - 1):
            self.convolutions.append(nn.Sequential(ConvNorm(emb, emb, kernel_size=k, stride=1, padding=pad, dilation=1,
                                                            w_init_gain="tanh"), nn.BatchNorm1d(emb)))
        self.convolutions.append(nn.Sequential(ConvNorm(emb, n_channels, kernel_size=k, stride=1, padding=pad, dilation=1,
                                                        w_init_gain="linear"), nn.BatchNorm1d(n_channels)))
        self.precision = DEFAULT_PRECISION
        self.__dict__["_tc"] = None

    def _convs(self):
        tc = self.__dict__.get("_tc")
        if tc is None or len(tc) != len(self.convolutions) or any(t.conv is not s[0].conv for t, s in zip(tc, self.convolutions)):
            tc = [_TcConv(s[0].conv, s[1]) for s in self.convolutions]
            self.__dict__["_tc"] = tc
        return tc

    def _forward_eager(self, x):
        x = x.contiguous().transpose(1, 2)
        for i in range(len(self.convolutions) - 1):
            x = F.dropout(torch.tanh(self.convolutions[i](x)), 0.5, self.training)
        x = F.dropout(self.convolutions[-1](x), 0.5, self.training)
        return x.contiguous().transpose(1, 2)

    def forward(self, x: torch.Tensor):
        if self.training or not _kernel_path(x, self):      # training-mode BatchNorm / dropout: the reference's formula
            return self._forward_eager(x)
        tc = self._convs()
        a = _operand(x.detach().to(torch.float32), self.precision, tc[0].padded_channels(x.device))
        for i, c in enumerate(tc[:-1]):
            _, a = c.run(a, self.precision, want_x=False, want_a=True, act_tanh=True)          # tanh(BN(conv)), 16-bit
            w = tc[i + 1].padded_channels(x.device)
            if a.shape[-1] != w:
                a = F.pad(a, (0, w - a.shape[-1])).contiguous()
        y, _ = tc[-1].run(a, self.precision, want_x=True, want_a=False)
        return y.to(x.dtype)


class AcousticTail(nn.Module):
    """decoder -> feats_linear -> postnet(outs) + outs -> (B, n_mel, T): FastSpeech2.inference, model.py:250-257.

    ``forward(frames (B, T, hidden), mel_len (B,))`` builds the padding mask the way the model does
    (``get_mask_from_lengths``, function.py:18-26: True on the padding) and returns the mel the vocoder consumes.
    """

    def __init__(self, decoder: nn.Module, feats_linear: nn.Module, postnet: Optional[nn.Module] = None):
        super().__init__()
        self.decoder, self.feats_linear, self.postnet = decoder, feats_linear, postnet

    def forward(self, frames: torch.Tensor, mel_len: torch.Tensor) -> torch.Tensor:
        T = frames.shape[1]
        mask = torch.arange(T, device=frames.device)[None, :] >= mel_len.to(frames.device)[:, None]
        hs, _ = self.decoder(frames, mask)
        outs = self.feats_linear(hs)
        if self.postnet is not None:
            outs = self.postnet(outs) + outs
        return outs.transpose(1, 2)

"""Conformer variant of the acoustic decoder (`block_type: "conformer"`, config/model_config.yaml:17-33).

Module shells with the reference's constructor signatures, attribute names and ``state_dict()`` keys for
models/tts/fastspeech2/blocks/conformer.py:93-571 (``Decoder`` -> here ``ConformerDecoder``, ``ConformerBlock``,
``ResidualConnectionModule``, ``FeedForwardModule``, ``MultiHeadedSelfAttentionModule``, ``RelativeMultiHeadAttention``,
``ConformerConvModule``, ``PointwiseConv1d``, ``DepthwiseConv1d``) and blocks/utils.py:46-86 (``LinearNorm``, ``Swish``,
``GLU``).  Quirks kept: the decoder's position table is registered as a parameter of every block's attention module
(conformer.py:330 assigns the ``nn.Parameter``), and the blocks run their attention WITHOUT a padding mask
(``nn.Sequential`` passes one argument, conformer.py:252-256) -- padded frames take part in the softmax.

Synthesis (no autograd, CUDA): the convolution module runs on the kernels -- both pointwise convs on the tcgen05 conv
kernel (``vtts_conv_*``, channels-last, so the reference's Transpose pair disappears) and GLU + depthwise conv (k = 31) +
eval-mode BatchNorm + Swish in one fused CUDA kernel (``vtts_dwconv_glu_swish``).  The feed-forward modules and the
relative attention are Linear / matmul chains and stay PyTorch (cuBLAS) ops.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib
from .acoustic import DEFAULT_PRECISION, _TcConv, _kernel_path, _operand, get_sinusoid_encoding_table


class LinearNorm(nn.Module):
    def __init__(self, in_features, out_features, bias=False):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias)
        nn.init.xavier_uniform_(self.linear.weight)
        if bias:
            nn.init.constant_(self.linear.bias, 0.0)

    def forward(self, x):
        return self.linear(x)


class Swish(nn.Module):
    def forward(self, inputs):
        return inputs * inputs.sigmoid()


class GLU(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, inputs):
        outputs, gate = inputs.chunk(2, dim=self.dim)
        return outputs * gate.sigmoid()


class Transpose(nn.Module):
    def __init__(self, shape: tuple):
        super().__init__()
        self.shape = shape

    def forward(self, x: Tensor) -> Tensor:
        return x.transpose(*self.shape)


class PointwiseConv1d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, stride: int = 1, padding: int = 0, bias: bool = True) -> None:
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=1, stride=stride, padding=padding, bias=bias)

    def forward(self, inputs: Tensor) -> Tensor:
        return self.conv(inputs)


class DepthwiseConv1d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1, padding: int = 0,
                 bias: bool = False) -> None:
        super().__init__()
        assert out_channels % in_channels == 0, "out_channels should be constant multiple of in_channels"
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, groups=in_channels, stride=stride,
                              padding=padding, bias=bias)

    def forward(self, inputs: Tensor) -> Tensor:
        return self.conv(inputs)


class ResidualConnectionModule(nn.Module):
    """outputs = module(inputs) * module_factor + inputs * input_factor (conformer.py:259-271)."""

    def __init__(self, module: nn.Module, module_factor: float = 1.0, input_factor: float = 1.0):
        super().__init__()
        self.module = module
        self.module_factor = module_factor
        self.input_factor = input_factor

    def forward(self, inputs: Tensor) -> Tensor:
        return (self.module(inputs) * self.module_factor) + (inputs * self.input_factor)


class FeedForwardModule(nn.Module):
    def __init__(self, encoder_dim: int = 512, expansion_factor: int = 4, dropout_p: float = 0.1) -> None:
        super().__init__()
        self.sequential = nn.Sequential(
            nn.LayerNorm(encoder_dim),
            LinearNorm(encoder_dim, encoder_dim * expansion_factor, bias=True),
            Swish(),
            nn.Dropout(p=dropout_p),
            LinearNorm(encoder_dim * expansion_factor, encoder_dim, bias=True),
            nn.Dropout(p=dropout_p),
        )

    def forward(self, inputs: Tensor) -> Tensor:
        return self.sequential(inputs)


class RelativeMultiHeadAttention(nn.Module):
    """Transformer-XL style relative attention (conformer.py:357-441), PyTorch ops."""

    def __init__(self, d_model: int = 512, num_heads: int = 16, dropout_p: float = 0.1):
        super().__init__()
        assert d_model % num_heads == 0, "d_model % num_heads should be zero."
        self.d_model = d_model
        self.d_head = int(d_model / num_heads)
        self.num_heads = num_heads
        self.sqrt_dim = math.sqrt(d_model)
        self.query_proj = LinearNorm(d_model, d_model)
        self.key_proj = LinearNorm(d_model, d_model)
        self.value_proj = LinearNorm(d_model, d_model)
        self.pos_proj = LinearNorm(d_model, d_model, bias=False)
        self.dropout = nn.Dropout(p=dropout_p)
        self.u_bias = nn.Parameter(torch.Tensor(self.num_heads, self.d_head))
        self.v_bias = nn.Parameter(torch.Tensor(self.num_heads, self.d_head))
        nn.init.xavier_uniform_(self.u_bias)
        nn.init.xavier_uniform_(self.v_bias)
        self.out_proj = LinearNorm(d_model, d_model)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, pos_embedding: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        batch_size = value.size(0)
        query = self.query_proj(query).view(batch_size, -1, self.num_heads, self.d_head)
        key = self.key_proj(key).view(batch_size, -1, self.num_heads, self.d_head).permute(0, 2, 1, 3)
        value = self.value_proj(value).view(batch_size, -1, self.num_heads, self.d_head).permute(0, 2, 1, 3)
        pos_embedding = self.pos_proj(pos_embedding).view(batch_size, -1, self.num_heads, self.d_head)
        content_score = torch.matmul((query + self.u_bias).transpose(1, 2), key.transpose(2, 3))
        pos_score = torch.matmul((query + self.v_bias).transpose(1, 2), pos_embedding.permute(0, 2, 3, 1))
        pos_score = self._relative_shift(pos_score)
        score = (content_score + pos_score) / self.sqrt_dim
        if mask is not None:
            score.masked_fill_(mask.unsqueeze(1), -1e9)
        attn = self.dropout(F.softmax(score, -1))
        context = torch.matmul(attn, value).transpose(1, 2)
        context = context.contiguous().view(batch_size, -1, self.d_model)
        return self.out_proj(context)

    def _relative_shift(self, pos_score: Tensor) -> Tensor:
        batch_size, num_heads, seq_length1, seq_length2 = pos_score.size()
        zeros = pos_score.new_zeros(batch_size, num_heads, seq_length1, 1)
        padded_pos_score = torch.cat([zeros, pos_score], dim=-1)
        padded_pos_score = padded_pos_score.view(batch_size, num_heads, seq_length2 + 1, seq_length1)
        return padded_pos_score[:, :, 1:].view_as(pos_score)


class MultiHeadedSelfAttentionModule(nn.Module):
    def __init__(self, d_model: int, num_heads: int, dropout_p: float = 0.1, position_enc: Optional[Tensor] = None,
                 max_seq_len: int = 10000):
        super().__init__()
        self.d_model = d_model
        self.max_seq_len = max_seq_len
        self.positional_encoding = position_enc      # an nn.Parameter here registers it on this module, like the reference
        self.layer_norm = nn.LayerNorm(d_model)
        self.attention = RelativeMultiHeadAttention(d_model, num_heads, dropout_p)
        self.dropout = nn.Dropout(p=dropout_p)

    def forward(self, inputs: Tensor, mask: Optional[Tensor] = None):
        batch_size, seq_length, _ = inputs.size()
        if not self.training and seq_length > self.max_seq_len:
            pos_embedding = get_sinusoid_encoding_table(seq_length, self.d_model)[:seq_length, :].unsqueeze(0) \
                .expand(batch_size, -1, -1).to(inputs.device)
        else:
            pos_embedding = self.positional_encoding[:, :seq_length, :].expand(batch_size, -1, -1)
        inputs = self.layer_norm(inputs)
        outputs = self.attention(inputs, inputs, inputs, pos_embedding=pos_embedding, mask=mask)
        return self.dropout(outputs)


class ConformerConvModule(nn.Module):
    """LayerNorm -> pointwise (C -> 2C) -> GLU -> depthwise (k) -> BatchNorm -> Swish -> pointwise (C -> C) (conformer.py:444-482)."""

    def __init__(self, in_channels: int, kernel_size: int = 31, expansion_factor: int = 2, dropout_p: float = 0.1) -> None:
        super().__init__()
        assert (kernel_size - 1) % 2 == 0, "kernel_size should be a odd number for 'SAME' padding"
        assert expansion_factor == 2, "Currently, Only Supports expansion_factor 2"
        self.sequential = nn.Sequential(
            nn.LayerNorm(in_channels),
            Transpose(shape=(1, 2)),
            PointwiseConv1d(in_channels, in_channels * expansion_factor, stride=1, padding=0, bias=True),
            GLU(dim=1),
            DepthwiseConv1d(in_channels, in_channels, kernel_size, stride=1, padding=(kernel_size - 1) // 2),
            nn.BatchNorm1d(in_channels),
            Swish(),
            PointwiseConv1d(in_channels, in_channels, stride=1, padding=0, bias=True),
            nn.Dropout(p=dropout_p),
        )
        self.precision = DEFAULT_PRECISION
        self.__dict__["_tc"] = None

    def _convs(self):
        tc = self.__dict__.get("_tc")
        pw1, pw2 = self.sequential[2].conv, self.sequential[7].conv
        if tc is None or tc[0].conv is not pw1 or tc[1].conv is not pw2:
            tc = (_TcConv(pw1), _TcConv(pw2))
            self.__dict__["_tc"] = tc
        return tc

    def _forward_eager(self, inputs: Tensor) -> Tensor:
        return self.sequential(inputs).transpose(1, 2)

    def forward(self, inputs: Tensor) -> Tensor:
        dw, bn = self.sequential[4].conv, self.sequential[5]
        if self.training or dw.kernel_size[0] > 31 or not _kernel_path(inputs, self):
            return self._forward_eager(inputs)
        lib = _lib.load()
        dev = inputs.device
        c1, c2 = self._convs()
        x = self.sequential[0](inputs.detach().to(torch.float32))                     # LayerNorm (PyTorch)
        a = _operand(x, self.precision, c1.padded_channels(dev))
        pw, _ = c1.run(a, self.precision, want_x=True, want_a=False)                   # (B, L, 2C) fp32
        B, L, C2 = pw.shape
        C = C2 // 2
        # eval-mode BatchNorm folded into the depthwise taps: y = (conv - mean) * gamma / sqrt(var + eps) + beta
        s = (bn.weight.detach() if bn.weight is not None else torch.ones(C, device=dev)) / torch.sqrt(bn.running_var.detach() + bn.eps)
        w = (dw.weight.detach().reshape(C, -1) * s[:, None]).to(torch.float32).contiguous()
        b = -bn.running_mean.detach() * s
        if dw.bias is not None:
            b = b + dw.bias.detach() * s
        if bn.bias is not None:
            b = b + bn.bias.detach()
        b = b.to(torch.float32).contiguous()
        dt = torch.float16 if self.precision == "fp16" else torch.bfloat16
        width = c2.padded_channels(dev)
        h = torch.empty((B, L, C), dtype=dt, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.vtts_dwconv_glu_swish(pw.data_ptr(), w.data_ptr(), b.data_ptr(), h.data_ptr(), _lib.PRECISION[self.precision],
                                                 B, L, C, dw.kernel_size[0], _lib.current_stream(dev)))
        if width != C:
            h = F.pad(h, (0, width - C)).contiguous()
        y, _ = c2.run(h, self.precision, want_x=True, want_a=False)
        return y.to(inputs.dtype)


class ConformerBlock(nn.Module):
    def __init__(self, encoder_dim: int = 512, num_attention_heads: int = 8, feed_forward_expansion_factor: int = 4,
                 conv_expansion_factor: int = 2, feed_forward_dropout_p: float = 0.1, attention_dropout_p: float = 0.1,
                 conv_dropout_p: float = 0.1, conv_kernel_size: int = 31, half_step_residual: bool = True,
                 position_enc: Tensor = None, max_seq_len: int = 10000):
        super().__init__()
        self.feed_forward_residual_factor = 0.5 if half_step_residual else 1
        self.sequential = nn.Sequential(
            ResidualConnectionModule(module=FeedForwardModule(encoder_dim=encoder_dim, expansion_factor=feed_forward_expansion_factor,
                                                              dropout_p=feed_forward_dropout_p),
                                     module_factor=self.feed_forward_residual_factor),
            ResidualConnectionModule(module=MultiHeadedSelfAttentionModule(d_model=encoder_dim, num_heads=num_attention_heads,
                                                                           dropout_p=attention_dropout_p, position_enc=position_enc,
                                                                           max_seq_len=max_seq_len)),
            ResidualConnectionModule(module=ConformerConvModule(in_channels=encoder_dim, kernel_size=conv_kernel_size,
                                                                expansion_factor=conv_expansion_factor, dropout_p=conv_dropout_p)),
            ResidualConnectionModule(module=FeedForwardModule(encoder_dim=encoder_dim, expansion_factor=feed_forward_expansion_factor,
                                                              dropout_p=feed_forward_dropout_p),
                                     module_factor=self.feed_forward_residual_factor),
            nn.LayerNorm(encoder_dim),
        )

    def forward(self, inputs: Tensor, mask: Tensor) -> Tensor:
        output = self.sequential(inputs)
        if mask is not None:
            output = output.masked_fill(mask.unsqueeze(-1), 0)
        return output


class ConformerDecoder(nn.Module):
    """blocks/conformer.py:93-169 (``Decoder``): position table + ``layers`` Conformer blocks over the expanded frames."""

    def __init__(self, layers: int, hidden_dim: int, max_seq_len: int, config: dict) -> None:
        super().__init__()
        self.config = config
        self.max_seq_len = max_seq_len
        self.d_model = hidden_dim
        self.position_enc = nn.Parameter(get_sinusoid_encoding_table(max_seq_len + 1, hidden_dim).unsqueeze(0), requires_grad=False)
        self.layer_stack = nn.ModuleList([
            ConformerBlock(encoder_dim=hidden_dim, num_attention_heads=config["decoder_head"],
                           feed_forward_expansion_factor=config["ffn_expansion_factor"],
                           conv_expansion_factor=config["conv_expansion_factor"],
                           feed_forward_dropout_p=config["decoder_dropout"], attention_dropout_p=config["decoder_dropout"],
                           conv_dropout_p=config["decoder_dropout"], conv_kernel_size=config["conv_kernel_size"],
                           half_step_residual=config["half_step_residual"], position_enc=self.position_enc,
                           max_seq_len=self.max_seq_len) for _ in range(layers)])

    def forward(self, enc_seq, mask):
        batch_size, max_len = enc_seq.shape[0], enc_seq.shape[1]
        if not self.training and enc_seq.shape[1] > self.max_seq_len:
            dec_output = enc_seq + get_sinusoid_encoding_table(enc_seq.shape[1], self.d_model)[: enc_seq.shape[1], :] \
                .unsqueeze(0).expand(batch_size, -1, -1).to(enc_seq.device)
        else:
            max_len = min(max_len, self.max_seq_len)
            dec_output = enc_seq[:, :max_len, :] + self.position_enc[:, :max_len, :].expand(batch_size, -1, -1)
            mask = mask[:, :max_len]
        for dec_layer in self.layer_stack:
            dec_output = dec_layer(dec_output, mask=mask)
        return dec_output, mask

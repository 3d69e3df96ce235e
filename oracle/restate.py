"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the synthesis hot path.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
(``viet-transformer-tts_b200/``) never does and fails loudly without its CUDA library.

Pinning: the reference ships no golden vectors, so this restatement is
pinned by executing the unmodified reference modules in the build container
(``oracle/ref_loader.py``) -- ``tests/test_oracle_pinned.py`` compares them directly when
``/root/reference`` exists, and ``tests/golden/*.npz`` (made by ``tests/golden/make_golden.py``
from the reference itself) travel to the GPU box.  JETS-specific composition is *parity
unpinned* (espnet is not vendored).

Every function cites the reference lines it restates (paths relative to the reference root).
The LengthRegulator is integer/byte work and is restated with numpy index arithmetic (and in
plain C in ``oracle/lr_oracle.c``); the generator is floating point and is restated with
``torch.nn.functional`` on the CPU in fp32 or fp64.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# LengthRegulator  (models/tts/fastspeech2/layers.py:434-462, pad_list function.py:97-124)
# --------------------------------------------------------------------------------------------


def lr_scale_durations(ds: torch.Tensor, alpha: float) -> torch.Tensor:
    """layers.py:446-448 -- ``ds = torch.round(ds.float() * alpha).long()`` (half-to-even)."""
    assert alpha > 0
    d32 = ds.detach().cpu().numpy().astype(np.float32)
    scaled = (d32 * np.float32(alpha)).astype(np.float32)
    return torch.from_numpy(np.rint(scaled).astype(np.int64))


def lr_expand(xs: torch.Tensor, ds: torch.Tensor, alpha: float = 1.0, pad_value: float = 0.0):
    """Restates ``LengthRegulator.forward`` (layers.py:434-462).

    Returns ``(out, ds_used)``.  ``ds`` is mutated IN PLACE on the all-zero-batch path exactly
    like the reference (layers.py:450-458): when the *whole batch* sums to zero every element
    of every all-zero row becomes 1.  ``out[b, t] = xs[b, j]`` with
    ``j = #{i : cumsum(ds[b])[i] <= t}`` for ``t < sum(ds[b])`` else ``pad_value``;
    ``T_out = max_b sum(ds[b])`` (pad_list, function.py:117-122).
    """
    if alpha != 1.0:
        ds = lr_scale_durations(ds, alpha)
    if int(ds.sum()) == 0:
        logging.warning(
            "predicted durations includes all 0 sequences. fill the first element with 1."
        )
        ds[ds.sum(dim=1).eq(0)] = 1
    d = ds.detach().cpu().numpy().astype(np.int64)
    if (d < 0).any():
        raise RuntimeError("repeats can not be negative")  # torch.repeat_interleave contract
    x = xs.detach().cpu().numpy()
    B, Tmax = d.shape
    lens = d.sum(axis=1)
    T_out = int(lens.max()) if B > 0 else 0
    out = np.full((B, T_out) + x.shape[2:], pad_value, dtype=x.dtype)
    for b in range(B):
        cum = np.cumsum(d[b])
        t = np.arange(int(lens[b]), dtype=np.int64)
        j = np.searchsorted(cum, t, side="right")  # first index with cum > t
        out[b, : int(lens[b])] = x[b, j]
    return torch.from_numpy(out), ds


def lr_mel_len(ds: torch.Tensor) -> torch.Tensor:
    """``mel_lens = torch.sum(duration_rounded, dim=1)`` -- VarianceAdaptor, layers.py:209."""
    return torch.from_numpy(ds.detach().cpu().numpy().astype(np.int64).sum(axis=1))


def durations_from_log(log_d: torch.Tensor, d_control: float = 1.0) -> torch.Tensor:
    """layers.py:205-208 -- ``clamp(round(exp(logd) - 1) * d_control, min=0).long()``."""
    return torch.clamp(torch.round(torch.exp(log_d) - 1) * d_control, min=0).long()


# --------------------------------------------------------------------------------------------
# weight-norm folding  (generator.py:185-195 -> torch.nn.utils.weight_norm, dim=0)
# --------------------------------------------------------------------------------------------


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """``w = g * v / ||v||`` with the norm over dims (1, 2) per dim-0 index.

    dim 0 is the *out* channel for Conv1d and the *in* channel for ConvTranspose1d
    (SURVEY.md appendix 9.5).
    """
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
    return v * (g / norm)


def _weight(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    return fold_weight_norm(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"])


def _bias(sd: Dict[str, torch.Tensor], prefix: str) -> Optional[torch.Tensor]:
    return sd.get(prefix + ".bias")


# --------------------------------------------------------------------------------------------
# HiFiGAN (ESPnet skin)  generator.py:132-156, ResidualBlock layers.py:83-98
# --------------------------------------------------------------------------------------------


def residual_block(
    sd, prefix: str, x: torch.Tensor, kernel_size: int, dilations: Sequence[int],
    slope: float = 0.1, use_additional_convs: bool = True,
) -> torch.Tensor:
    """layers.py:93-97: ``xt = convs2[i](convs1[i](x)); x = xt + x`` (each = LeakyReLU + Conv1d)."""
    for idx, d in enumerate(dilations):
        xt = F.conv1d(
            F.leaky_relu(x, slope), _weight(sd, f"{prefix}.convs1.{idx}.1"),
            _bias(sd, f"{prefix}.convs1.{idx}.1"), padding=(kernel_size - 1) // 2 * d, dilation=d,
        )
        if use_additional_convs:
            xt = F.conv1d(
                F.leaky_relu(xt, slope), _weight(sd, f"{prefix}.convs2.{idx}.1"),
                _bias(sd, f"{prefix}.convs2.{idx}.1"), padding=(kernel_size - 1) // 2,
            )
        x = xt + x
    return x


def hifigan_forward(
    sd: Dict[str, torch.Tensor], c: torch.Tensor, g: Optional[torch.Tensor] = None,
    upsample_scales: Sequence[int] = (8, 8, 2, 2),
    resblock_kernel_sizes: Sequence[int] = (3, 7, 11),
    resblock_dilations: Sequence[Sequence[int]] = ((1, 3, 5), (1, 3, 5), (1, 3, 5)),
    kernel_size: int = 7, slope: float = 0.1, use_additional_convs: bool = True,
    dtype: torch.dtype = torch.float32, return_stages: bool = False,
):
    """Restates ``HiFiGAN.forward`` (generator.py:132-156) from a reference ``state_dict``.

    (B, in_ch, T) -> (B, out_ch, T * prod(scales)).  The final activation slope is the
    ``nn.LeakyReLU()`` default 0.01 (generator.py:111), not ``slope``.
    """
    sd = {k: v.detach().to("cpu", dtype) for k, v in sd.items()}
    c = c.detach().to("cpu", dtype)
    stages: List[torch.Tensor] = []
    c = F.conv1d(c, _weight(sd, "input_conv"), _bias(sd, "input_conv"), padding=(kernel_size - 1) // 2)
    if g is not None:
        c = c + F.conv1d(g.detach().to("cpu", dtype), _weight(sd, "global_conv"), _bias(sd, "global_conv"))
    stages.append(c)
    nb = len(resblock_kernel_sizes)
    for i, s in enumerate(upsample_scales):
        c = F.conv_transpose1d(
            F.leaky_relu(c, slope), _weight(sd, f"upsamples.{i}.1"), _bias(sd, f"upsamples.{i}.1"),
            stride=s, padding=s // 2 + s % 2, output_padding=s % 2,
        )
        stages.append(c)
        cs = 0.0
        for j in range(nb):
            cs = cs + residual_block(
                sd, f"blocks.{i * nb + j}", c, resblock_kernel_sizes[j], resblock_dilations[j],
                slope, use_additional_convs,
            )
        c = cs / nb
        stages.append(c)
    c = torch.tanh(
        F.conv1d(F.leaky_relu(c, 0.01), _weight(sd, "output_conv.1"), _bias(sd, "output_conv.1"),
                 padding=(kernel_size - 1) // 2)
    )
    return (c, stages) if return_stages else c


def hifigan_inference(sd, c: torch.Tensor, g: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
    """generator.py:197-213: (T, in_ch) -> (T * upsample_factor, out_ch)."""
    if g is not None:
        g = g.unsqueeze(0)
    y = hifigan_forward(sd, c.transpose(1, 0).unsqueeze(0), g=g, **kw)
    return y.squeeze(0).transpose(1, 0)


# --------------------------------------------------------------------------------------------
# vits2 Generator skin  (vits2/layers.py:159-177, ResBlock1/2 sublayers.py:293-303, 341-349)
# --------------------------------------------------------------------------------------------


def vits2_generator_forward(
    sd: Dict[str, torch.Tensor], x: torch.Tensor, g: Optional[torch.Tensor] = None,
    resblock: str = "1", upsample_rates: Sequence[int] = (8, 8, 2, 2),
    upsample_kernel_sizes: Sequence[int] = (16, 16, 4, 4),
    resblock_kernel_sizes: Sequence[int] = (3, 7, 11),
    resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 3, 5), (1, 3, 5), (1, 3, 5)),
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """Restates vits2 ``Generator.forward``: conv_pre/conv_post carry no weight-norm, conv_post
    has no bias (layers.py:153), ups padding is ``(k - u) // 2`` (layers.py:140)."""
    sd = {k: v.detach().to("cpu", dtype) for k, v in sd.items()}
    x = x.detach().to("cpu", dtype)
    x = F.conv1d(x, _weight(sd, "conv_pre"), _bias(sd, "conv_pre"), padding=3)
    if g is not None:
        x = x + F.conv1d(g.detach().to("cpu", dtype), _weight(sd, "cond"), _bias(sd, "cond"))
    nk = len(resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, _weight(sd, f"ups.{i}"), _bias(sd, f"ups.{i}"), stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            p = f"resblocks.{i * nk + j}"
            ks, dil = resblock_kernel_sizes[j], resblock_dilation_sizes[j]
            y = x
            if resblock == "1":
                for m, d in enumerate(dil):
                    xt = F.conv1d(F.leaky_relu(y, 0.1), _weight(sd, f"{p}.convs1.{m}"), _bias(sd, f"{p}.convs1.{m}"),
                                  padding=(ks * d - d) // 2, dilation=d)
                    xt = F.conv1d(F.leaky_relu(xt, 0.1), _weight(sd, f"{p}.convs2.{m}"), _bias(sd, f"{p}.convs2.{m}"),
                                  padding=(ks - 1) // 2)
                    y = xt + y
            else:
                for m, d in enumerate(dil):
                    xt = F.conv1d(F.leaky_relu(y, 0.1), _weight(sd, f"{p}.convs.{m}"), _bias(sd, f"{p}.convs.{m}"),
                                  padding=(ks * d - d) // 2, dilation=d)
                    y = xt + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01, layers.py:174
    x = F.conv1d(x, _weight(sd, "conv_post"), None, padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------------------------
# synthetic state_dicts (same key set / shapes as the reference constructors)
# --------------------------------------------------------------------------------------------


def make_hifigan_state_dict(
    in_channels: int = 80, out_channels: int = 1, channels: int = 512, global_channels: int = -1,
    kernel_size: int = 7, upsample_scales: Sequence[int] = (8, 8, 2, 2),
    resblock_kernel_sizes: Sequence[int] = (3, 7, 11),
    resblock_dilations: Sequence[Sequence[int]] = ((1, 3, 5), (1, 3, 5), (1, 3, 5)),
    use_additional_convs: bool = True, seed: int = 1234, weight_norm: bool = True,
) -> Dict[str, torch.Tensor]:
    """Random-init ``state_dict`` with the reference's key names and shapes (generator.py:70-123).

    Values are Kaiming-uniform-like (what the reference effectively has) but are
    drawn by this function, not by the reference constructor -- use the reference's own
    ``state_dict()`` (tests/golden) when bit-identical weights matter.
    """
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(prefix, cout, cin, k, transpose=False):
        shape = (cin, cout, k) if transpose else (cout, cin, k)
        fan_in = shape[1] * k
        bound = 1.0 / np.sqrt(fan_in)
        w = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        b = (torch.rand(cout, generator=gen) * 2 - 1) * bound
        if weight_norm:
            sd[prefix + ".weight_g"] = w.reshape(shape[0], -1).norm(dim=1).reshape(-1, 1, 1) * (
                0.75 + 0.5 * torch.rand(shape[0], 1, 1, generator=gen))
            sd[prefix + ".weight_v"] = w
        else:
            sd[prefix + ".weight"] = w
        sd[prefix + ".bias"] = b

    conv("input_conv", channels, in_channels, kernel_size)
    ch = channels
    for i, s in enumerate(upsample_scales):
        conv(f"upsamples.{i}.1", ch // 2, ch, 2 * s, transpose=True)
        ch //= 2
        for j, k in enumerate(resblock_kernel_sizes):
            n = i * len(resblock_kernel_sizes) + j
            for m in range(len(resblock_dilations[j])):
                conv(f"blocks.{n}.convs1.{m}.1", ch, ch, k)
                if use_additional_convs:
                    conv(f"blocks.{n}.convs2.{m}.1", ch, ch, k)
    conv("output_conv.1", out_channels, ch, kernel_size)
    if global_channels > 0:
        conv("global_conv", channels, global_channels, 1)
    return sd


# --------------------------------------------------------------------------------------------
# GaussianUpsampling  (models/tts/fastspeech2/layers.py:465-520; same body in jets/alignments.py:168-222)
# --------------------------------------------------------------------------------------------


def gaussian_upsampling(hs: torch.Tensor, ds: torch.Tensor, h_masks: Optional[torch.Tensor] = None,
                        d_masks: Optional[torch.Tensor] = None, delta: float = 0.1):
    """Restates ``GaussianUpsampling.forward`` (layers.py:476-520) with explicit fp32 steps.

    ``ds`` is mutated in place on the all-zero-batch path (layers.py:492-499).  Without ``h_masks`` the
    number of output frames is the duration sum over the WHOLE batch (layers.py:501-502) -- a reference
    quirk that is preserved, not fixed.  Returns (B, T_feats, adim).
    """
    B = ds.size(0)
    if int(ds.sum()) == 0:
        logging.warning(
            "predicted durations includes all 0 sequences. fill the first element with 1."
        )
        ds[ds.sum(dim=1).eq(0)] = 1
    T_feats = int(ds.sum()) if h_masks is None else h_masks.size(-1)
    t = torch.arange(0, T_feats).unsqueeze(0).repeat(B, 1).float()          # layers.py:505
    if h_masks is not None:
        t = t * h_masks.float()                                             # layers.py:506-507
    c = ds.cumsum(dim=-1).float() - ds.float() / 2                          # layers.py:509 (int64 - float32 -> float32)
    energy = (-1 * delta) * (t.unsqueeze(-1) - c.unsqueeze(1)) ** 2         # layers.py:510
    if d_masks is not None:
        energy = energy.masked_fill(~(d_masks.unsqueeze(1).repeat(1, T_feats, 1)), -float("inf"))
    p_attn = torch.softmax(energy, dim=2)                                   # layers.py:516
    return torch.matmul(p_attn, hs.float())                                 # layers.py:517


# --------------------------------------------------------------------------------------------
# vits2 monotonic duration path  (models/gan_tts/vits2/utils.py:104-126; call site vits2/generator.py:251-259)
# --------------------------------------------------------------------------------------------


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Restates ``generate_path`` (utils.py:111-126) with explicit index arithmetic.

    duration [b, 1, t_x], mask [b, 1, t_y, t_x] -> path [b, 1, t_y, t_x]:
    ``path[b,0,y,x] = ((y < cum[b,x]) - (y < cum[b,x-1])) * mask[b,0,y,x]`` with ``cum = cumsum(duration)``
    (``sequence_mask`` compares an arange of duration's dtype, utils.py:104-108).
    """
    b, _, t_y, t_x = mask.shape
    cum = torch.cumsum(duration, -1).view(b, t_x)                       # utils.py:120-122
    y = torch.arange(t_y, dtype=cum.dtype).view(1, t_y, 1)
    below = (y < cum.view(b, 1, t_x)).to(mask.dtype)                    # sequence_mask, transposed to (b, t_y, t_x)
    prev = torch.zeros_like(below)
    prev[:, :, 1:] = below[:, :, :-1]                                   # F.pad(path, [[0,0],[1,0],[0,0]])[:, :-1]
    return (below - prev).unsqueeze(1) * mask                           # utils.py:124-125


def expand_by_path(x: torch.Tensor, duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """``torch.matmul(attn.squeeze(1), x.transpose(1, 2)).transpose(1, 2)`` (vits2/generator.py:256-259)."""
    attn = generate_path(duration, mask)
    return torch.matmul(attn.squeeze(1), x.transpose(1, 2)).transpose(1, 2)


# --------------------------------------------------------------------------------------------
# Acoustic decoder + Postnet
#   Decoder / FFTBlock / MultiHeadAttention / PositionwiseFeedForward: models/tts/fastspeech2/blocks/transformer.py:90-298
#   Postnet: models/tts/fastspeech2/layers.py:571-625 (eval mode: BatchNorm1d uses its running statistics, dropout off)
# --------------------------------------------------------------------------------------------


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """blocks/utils.py:14-35."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    idx = np.arange(d_hid)[None, :]
    t = pos / np.power(10000, 2 * (idx // 2) / d_hid)
    t[:, 0::2] = np.sin(t[:, 0::2])
    t[:, 1::2] = np.cos(t[:, 1::2])
    return torch.FloatTensor(t)


def fft_decoder_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, mask: torch.Tensor, n_head: int,
                        prefix: str = "") -> torch.Tensor:
    """Decoder.forward (transformer.py:132-166, eval, T <= max_seq_len) -> (B, T, d_model).  mask: True on the padding."""
    B, T, D = x.shape
    out = x + sd[prefix + "position_enc"][:, :T, :]
    attn_mask = mask.unsqueeze(1).expand(-1, T, -1)
    d_k = D // n_head
    layer = 0
    while f"{prefix}layer_stack.{layer}.slf_attn.w_qs.weight" in sd:
        p = f"{prefix}layer_stack.{layer}."
        # MultiHeadAttention (transformer.py:215-243)
        res = out
        q = F.linear(out, sd[p + "slf_attn.w_qs.weight"], sd[p + "slf_attn.w_qs.bias"]).view(B, T, n_head, d_k)
        k = F.linear(out, sd[p + "slf_attn.w_ks.weight"], sd[p + "slf_attn.w_ks.bias"]).view(B, T, n_head, d_k)
        v = F.linear(out, sd[p + "slf_attn.w_vs.weight"], sd[p + "slf_attn.w_vs.bias"]).view(B, T, n_head, d_k)
        q, k, v = (t.permute(2, 0, 1, 3).reshape(-1, T, d_k) for t in (q, k, v))
        a = torch.bmm(q, k.transpose(1, 2)) / float(np.power(d_k, 0.5))
        a = a.masked_fill(attn_mask.repeat(n_head, 1, 1), -np.inf)
        a = torch.softmax(a, dim=2)
        o = torch.bmm(a, v).view(n_head, B, T, d_k).permute(1, 2, 0, 3).reshape(B, T, -1)
        o = F.linear(o, sd[p + "slf_attn.fc.weight"], sd[p + "slf_attn.fc.bias"])
        out = F.layer_norm(o + res, (D,), sd[p + "slf_attn.layer_norm.weight"], sd[p + "slf_attn.layer_norm.bias"])
        out = out.masked_fill(mask.unsqueeze(-1), 0)
        # PositionwiseFeedForward (transformer.py:288-298)
        res = out
        w1, w2 = sd[p + "pos_ffn.w_1.weight"], sd[p + "pos_ffn.w_2.weight"]
        h = F.conv1d(out.transpose(1, 2), w1, sd[p + "pos_ffn.w_1.bias"], padding=(w1.shape[-1] - 1) // 2)
        h = F.conv1d(F.relu(h), w2, sd[p + "pos_ffn.w_2.bias"], padding=(w2.shape[-1] - 1) // 2).transpose(1, 2)
        out = F.layer_norm(h + res, (D,), sd[p + "pos_ffn.layer_norm.weight"], sd[p + "pos_ffn.layer_norm.bias"])
        out = out.masked_fill(mask.unsqueeze(-1), 0)
        layer += 1
    return out


def postnet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "", eps: float = 1e-5) -> torch.Tensor:
    """Postnet.forward (layers.py:614-621) in eval mode; x and the result are (B, T, n_mel)."""
    y = x.transpose(1, 2)
    n = 0
    while f"{prefix}convolutions.{n}.0.conv.weight" in sd:
        n += 1
    for i in range(n):
        p = f"{prefix}convolutions.{i}."
        w = sd[p + "0.conv.weight"]
        y = F.conv1d(y, w, sd.get(p + "0.conv.bias"), padding=(w.shape[-1] - 1) // 2)
        y = F.batch_norm(y, sd[p + "1.running_mean"], sd[p + "1.running_var"], sd[p + "1.weight"], sd[p + "1.bias"], False, 0.0, eps)
        if i < n - 1:
            y = torch.tanh(y)
    return y.transpose(1, 2)


def acoustic_tail_forward(sd: Dict[str, torch.Tensor], frames: torch.Tensor, mel_len: torch.Tensor, n_head: int) -> torch.Tensor:
    """FastSpeech2.inference tail (model.py:250-257) with keys ``decoder.*``, ``feats_linear.*``, ``postnet.*`` -> (B, n_mel, T)."""
    T = frames.shape[1]
    mask = torch.arange(T)[None, :] >= mel_len[:, None]
    hs = fft_decoder_forward(sd, frames, mask, n_head, prefix="decoder.")
    outs = F.linear(hs, sd["feats_linear.weight"], sd["feats_linear.bias"])
    if "postnet.convolutions.0.0.conv.weight" in sd:
        outs = postnet_forward(sd, outs, prefix="postnet.") + outs
    return outs.transpose(1, 2)


# --------------------------------------------------------------------------------------------
# Conformer decoder (models/tts/fastspeech2/blocks/conformer.py:93-571), eval mode, T <= max_seq_len
# --------------------------------------------------------------------------------------------


def _rel_shift(s: torch.Tensor) -> torch.Tensor:
    """conformer.py:432-441."""
    b, h, t1, t2 = s.shape
    z = s.new_zeros(b, h, t1, 1)
    return torch.cat([z, s], dim=-1).view(b, h, t2 + 1, t1)[:, :, 1:].view_as(s)


def conformer_decoder_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, mask: torch.Tensor, n_head: int,
                              half_step: bool = True, prefix: str = "", bn_eps: float = 1e-5) -> torch.Tensor:
    """Decoder.forward (conformer.py:144-169): position table, then per block FFN/2 + rel-MHSA (UNMASKED: the Sequential
    passes no mask, :252-253) + conv module + FFN/2 + LayerNorm, padded frames zeroed after every block (:254-255)."""
    B, T, D = x.shape
    out = x + sd[prefix + "position_enc"][:, :T, :]
    d_head = D // n_head
    ffn_f = 0.5 if half_step else 1.0

    def ln(v, p):
        return F.layer_norm(v, (D,), sd[p + ".weight"], sd[p + ".bias"])

    def ffn(v, p):
        h = F.linear(ln(v, p + ".sequential.0"), sd[p + ".sequential.1.linear.weight"], sd[p + ".sequential.1.linear.bias"])
        h = h * torch.sigmoid(h)
        return F.linear(h, sd[p + ".sequential.4.linear.weight"], sd[p + ".sequential.4.linear.bias"])

    layer = 0
    while f"{prefix}layer_stack.{layer}.sequential.4.weight" in sd:
        p = f"{prefix}layer_stack.{layer}.sequential."
        out = ffn(out, p + "0.module") * ffn_f + out
        # MultiHeadedSelfAttentionModule (:336-354) + RelativeMultiHeadAttention (:400-430)
        a = p + "1.module"
        pos = sd[a + ".positional_encoding"][:, :T, :].expand(B, -1, -1)
        xin = ln(out, a + ".layer_norm")
        q = F.linear(xin, sd[a + ".attention.query_proj.linear.weight"]).view(B, T, n_head, d_head)
        k = F.linear(xin, sd[a + ".attention.key_proj.linear.weight"]).view(B, T, n_head, d_head).permute(0, 2, 1, 3)
        v = F.linear(xin, sd[a + ".attention.value_proj.linear.weight"]).view(B, T, n_head, d_head).permute(0, 2, 1, 3)
        pe = F.linear(pos, sd[a + ".attention.pos_proj.linear.weight"]).view(B, T, n_head, d_head)
        cs = torch.matmul((q + sd[a + ".attention.u_bias"]).transpose(1, 2), k.transpose(2, 3))
        ps = _rel_shift(torch.matmul((q + sd[a + ".attention.v_bias"]).transpose(1, 2), pe.permute(0, 2, 3, 1)))
        att = torch.softmax((cs + ps) / float(np.sqrt(D)), -1)
        ctx = torch.matmul(att, v).transpose(1, 2).reshape(B, T, D)
        out = F.linear(ctx, sd[a + ".attention.out_proj.linear.weight"]) + out
        # ConformerConvModule (:470-482)
        c = p + "2.module.sequential."
        h = ln(out, c + "0").transpose(1, 2)
        h = F.conv1d(h, sd[c + "2.conv.weight"], sd[c + "2.conv.bias"])
        h = h[:, :D] * torch.sigmoid(h[:, D:])
        wd = sd[c + "4.conv.weight"]
        h = F.conv1d(h, wd, None, padding=(wd.shape[-1] - 1) // 2, groups=D)
        h = F.batch_norm(h, sd[c + "5.running_mean"], sd[c + "5.running_var"], sd[c + "5.weight"], sd[c + "5.bias"], False, 0.0, bn_eps)
        h = h * torch.sigmoid(h)
        h = F.conv1d(h, sd[c + "7.conv.weight"], sd[c + "7.conv.bias"]).transpose(1, 2)
        out = h + out
        out = ffn(out, p + "3.module") * ffn_f + out
        out = ln(out, p + "4")
        out = out.masked_fill(mask.unsqueeze(-1), 0)
        layer += 1
    return out

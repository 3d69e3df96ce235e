"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests or fixtures, so these files -- produced by
executing the reference's own modules (oracle/ref_loader.py) on seeded synthetic inputs -- are
what pins both the CPU oracle (oracle/restate.py, oracle/lr_oracle.c) and the CUDA path.  They
travel to the GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
logging.disable(logging.WARNING)


def sd_to_np(sd, prefix="sd."):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def lr_cases():
    LR = ref_loader.load_length_regulator()
    g = torch.Generator().manual_seed(0)
    cases = {}

    def add(name, xs, ds, alpha=1.0, pad=0.0):
        ds_in = ds.clone()
        ds_work = ds.clone()
        out = LR(pad_value=pad)(xs, ds_work, alpha)
        cases[name] = dict(xs=xs.numpy(), ds=ds_in.numpy(), ds_after=ds_work.numpy(), out=out.numpy(),
                           alpha=np.float64(alpha), pad=np.float64(pad))

    # plain ragged batch, zeros on the padding
    B, T, D = 4, 13, 80
    ds = torch.randint(1, 12, (B, T), generator=g)
    for b, n in enumerate([13, 9, 5, 1]):
        ds[b, n:] = 0
    add("ragged", torch.randn(B, T, D, generator=g), ds)
    # zeros inside rows + one all-zero row inside a non-zero batch, non-zero pad value
    ds = torch.randint(0, 4, (3, 10), generator=g)
    ds[1] = 0
    add("zero_row", torch.randn(3, 10, 192, generator=g), ds, pad=-1.5)
    # whole batch zero: in-place fix-up to ones, T_out = Tmax
    add("all_zero_batch", torch.randn(2, 7, 256, generator=g), torch.zeros(2, 7, dtype=torch.long))
    # alpha != 1 (round half to even): appendix 9.6 example plus a random one
    add("alpha_1p5_small", torch.randn(2, 3, 4, generator=g), torch.tensor([[2, 1, 1], [1, 1, 1]]), alpha=1.5)
    add("alpha_0p5", torch.randn(3, 11, 384, generator=g), torch.randint(0, 9, (3, 11), generator=g), alpha=0.5)
    add("alpha_1p5", torch.randn(3, 11, 384, generator=g), torch.randint(0, 9, (3, 11), generator=g), alpha=1.5)
    # alpha that rounds everything to zero -> fix-up happens on the scaled copy, caller's ds untouched
    add("alpha_to_zero", torch.randn(2, 5, 16, generator=g), torch.ones(2, 5, dtype=torch.long), alpha=0.25)
    # odd feature width (no 16-byte rows), Tmax not a multiple of 4
    add("odd_width", torch.randn(2, 7, 5, generator=g), torch.randint(0, 5, (2, 7), generator=g))
    add("width_1", torch.randn(2, 6, 1, generator=g), torch.randint(1, 3, (2, 6), generator=g))
    # single long row
    add("long_row", torch.randn(1, 300, 64, generator=g), torch.randint(0, 7, (1, 300), generator=g))
    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(OUT, "lr_cases.npz"), **flat)
    print("lr_cases:", list(cases))


def hifigan_small():
    HiFiGAN, _ = ref_loader.load_hifigan()
    torch.manual_seed(7)
    cfg = dict(in_channels=8, out_channels=1, channels=32, global_channels=4, kernel_size=7,
               upsample_scales=[4, 2], upsample_kernel_sizes=[8, 4], resblock_kernel_sizes=[3, 5],
               resblock_dilations=[[1, 3], [1, 2]])
    m = HiFiGAN(**cfg).eval()
    g = torch.Generator().manual_seed(1)
    c = torch.randn(3, 8, 37, generator=g)
    gc = torch.randn(3, 4, 1, generator=g)
    with torch.no_grad():
        y = m(c, gc)
        y_nog = m(c)
    np.savez_compressed(os.path.join(OUT, "hifigan_small.npz"), c=c.numpy(), g=gc.numpy(), y=y.numpy(),
                        y_nog=y_nog.numpy(), **sd_to_np(m.state_dict()))
    # use_additional_convs=False, bias=False, odd upsample scale (output_padding path), no weight norm
    torch.manual_seed(8)
    cfg2 = dict(in_channels=6, out_channels=1, channels=16, kernel_size=5, upsample_scales=[3, 2],
                upsample_kernel_sizes=[6, 4], resblock_kernel_sizes=[3], resblock_dilations=[[1, 2, 3]],
                use_additional_convs=False, bias=False, use_weight_norm=False)
    m2 = HiFiGAN(**cfg2).eval()
    c2 = torch.randn(2, 6, 21, generator=g)
    with torch.no_grad():
        y2 = m2(c2)
    np.savez_compressed(os.path.join(OUT, "hifigan_small2.npz"), c=c2.numpy(), y=y2.numpy(), **sd_to_np(m2.state_dict()))
    print("hifigan_small:", tuple(y.shape), tuple(y2.shape))


def hifigan_v1():
    HiFiGAN, _ = ref_loader.load_hifigan()
    torch.manual_seed(1234)  # config/train_config.yaml:1
    m = HiFiGAN().eval()
    sd = m.state_dict()
    g = torch.Generator().manual_seed(0)
    c = torch.randn(2, 80, 24, generator=g)
    with torch.no_grad():
        y = m(c)
        inf = m.inference(c[0].transpose(0, 1))
    keys = sorted(sd.keys())
    sums = np.array([float(sd[k].double().sum()) for k in keys])
    abss = np.array([float(sd[k].double().abs().sum()) for k in keys])
    np.savez_compressed(os.path.join(OUT, "hifigan_v1.npz"), c=c.numpy(), y=y.numpy(), inference=inf.numpy(),
                        keys=np.array(keys), shapes=np.array([str(tuple(sd[k].shape)) for k in keys]),
                        sums=sums, abs_sums=abss, seed=np.int64(1234))
    print("hifigan_v1:", tuple(y.shape), len(keys))


def vits2_small():
    Generator, _, _ = ref_loader.load_vits2_generator()
    g = torch.Generator().manual_seed(2)
    out = {}
    for tag, rb, dil in (("rb1", "1", [[1, 3, 5], [1, 2, 4]]), ("rb2", "2", [[1, 3], [1, 2]])):
        torch.manual_seed(11)
        import contextlib, io
        m = Generator(12, resblock=rb, resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=dil,
                      upsample_rates=[4, 2], upsample_initial_channel=32, upsample_kernel_sizes=[8, 4],
                      gin_channels=5).eval()
        x = torch.randn(2, 12, 19, generator=g)
        gc = torch.randn(2, 5, 1, generator=g)
        with torch.no_grad():
            y = m(x, gc)
        out.update({f"{tag}.x": x.numpy(), f"{tag}.g": gc.numpy(), f"{tag}.y": y.numpy()})
        out.update(sd_to_np(m.state_dict(), prefix=f"{tag}.sd."))
    np.savez_compressed(os.path.join(OUT, "vits2_small.npz"), **out)
    print("vits2_small ok")


def gaussian_cases():
    GU = ref_loader.load_gaussian_upsampling()
    g = torch.Generator().manual_seed(3)
    out = {}

    def add(name, hs, ds, h_masks=None, d_masks=None, delta=0.1):
        ds_work = ds.clone()
        y = GU(delta=delta)(hs, ds_work, h_masks, d_masks)
        out[f"{name}.hs"] = hs.numpy(); out[f"{name}.ds"] = ds.numpy(); out[f"{name}.ds_after"] = ds_work.numpy()
        out[f"{name}.y"] = y.numpy(); out[f"{name}.delta"] = np.float64(delta)
        if h_masks is not None: out[f"{name}.h_masks"] = h_masks.numpy()
        if d_masks is not None: out[f"{name}.d_masks"] = d_masks.numpy()

    # the VarianceAdaptor call (layers.py:231): h_masks = ~mel_mask, d_masks = ~txt_mask
    B, T, D = 3, 11, 48
    tl = torch.tensor([11, 7, 4])
    ds = torch.randint(1, 6, (B, T), generator=g)
    ds[torch.arange(T)[None] >= tl[:, None]] = 0
    ml = ds.sum(1)
    h = torch.arange(int(ml.max()))[None] < ml[:, None]
    d = torch.arange(T)[None] < tl[:, None]
    add("masked", torch.randn(B, T, D, generator=g), ds, h, d)
    # no masks: T_feats = sum over the WHOLE batch (quirk)
    add("nomask", torch.randn(2, 5, 16, generator=g), torch.randint(0, 4, (2, 5), generator=g))
    # zero durations inside rows, other delta, D not a multiple of 32
    ds2 = torch.tensor([[0, 3, 0, 2, 1], [2, 0, 0, 0, 4]])
    add("zeros", torch.randn(2, 5, 37, generator=g), ds2, torch.ones(2, 9, dtype=torch.bool), torch.ones(2, 5, dtype=torch.bool), delta=0.35)
    # all-zero batch -> in-place fix-up
    add("all_zero", torch.randn(2, 4, 8, generator=g), torch.zeros(2, 4, dtype=torch.long), torch.ones(2, 4, dtype=torch.bool), None)
    np.savez_compressed(os.path.join(OUT, "gaussian_cases.npz"), **out)
    print("gaussian_cases ok")


def path_cases():
    """vits2 generate_path + the attn matmuls of VITS2.inference (vits2/generator.py:246-259), run on the reference."""
    U = ref_loader.load_vits2_utils()
    g = torch.Generator().manual_seed(7)
    out = {}

    def add(name, w, x_len, d_ch, feats_lengths=None):
        b, _, t_x = w.shape
        x_mask = U.sequence_mask(x_len, t_x).unsqueeze(1).to(w.dtype)                    # [b,1,t_x]
        w = w * x_mask
        w_ceil = torch.ceil(w)
        if feats_lengths is None:
            feats_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()          # generator.py:251
        y_mask = torch.unsqueeze(U.sequence_mask(feats_lengths, None), 1).to(x_mask.dtype)
        attn_mask = torch.unsqueeze(x_mask, 2) * torch.unsqueeze(y_mask, -1)
        attn = U.generate_path(w_ceil, attn_mask)
        m_p = torch.randn(b, d_ch, t_x, generator=g)
        m_out = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)
        out.update({f"{name}.w_ceil": w_ceil.numpy(), f"{name}.attn_mask": attn_mask.numpy(), f"{name}.attn": attn.numpy(),
                    f"{name}.m_p": m_p.numpy(), f"{name}.m_out": m_out.contiguous().numpy()})

    add("basic", torch.rand(3, 1, 13, generator=g) * 4, torch.tensor([13, 9, 5]), 24)
    add("zeros_inside", torch.tensor([[[0.0, 2.2, 0.0, 0.0, 3.0, 0.4]], [[1.0, 0.0, 0.0, 0.0, 0.0, 5.5]]]), torch.tensor([6, 6]), 7)
    add("all_zero_row", torch.tensor([[[0.0, 0.0, 0.0]], [[2.0, 1.0, 3.0]]]), torch.tensor([3, 3]), 5)      # clamp_min(…, 1)
    add("long", torch.rand(4, 1, 120, generator=g) * 9, torch.tensor([120, 77, 40, 101]), 40)
    add("short_y", torch.rand(2, 1, 10, generator=g) * 5, torch.tensor([10, 10]), 16, feats_lengths=torch.tensor([9, 4]))
    np.savez_compressed(os.path.join(OUT, "path_cases.npz"), **out)
    print("path_cases ok")


def fastspeech2_capture():
    """a9: tensors captured INSIDE the reference's FastSpeech2.inference (fastspeech2/model.py:194-257) on the C2 recipe
    (4-layer / 256-hidden transformer blocks, duration predictor bias = log 7), reduced to 3 utterances so the
    fixture stays small: the regulator's inputs and output as VarianceAdaptor.forward passes them (layers.py:226-233),
    the final mel, feats_lengths and the reference HiFiGAN's waveform for that mel (Text2Wav.inference order,
    text2wav/model.py:139-167)."""
    import math
    import yaml

    ref_loader.load_length_regulator()
    from models.tts.fastspeech2.model import FastSpeech2  # type: ignore
    HiFiGAN, _ = ref_loader.load_hifigan()
    out = {}
    for tag, use_gaussian in (("lr", False), ("gauss", True)):
        with open(os.path.join(ref_loader.REF_ROOT, "config/model_config.yaml")) as fh:
            cfg = yaml.safe_load(fh)["fastspeech2"]
        cfg["use_cvae"] = False
        cfg["encoder_layers"] = cfg["decoder_layers"] = 4
        cfg["encoder_hidden"] = cfg["decoder_hidden"] = 256
        cfg["building_block"]["block_type"] = "transformer"
        cfg["variance"]["learn_alignment"] = False
        cfg["variance"]["duration_modelling"]["use_gaussian"] = use_gaussian
        stats = {"pitch": {"min": -2.0, "max": 8.0}, "energy": {"min": -1.5, "max": 7.0}}
        torch.manual_seed(1234)
        m = FastSpeech2(n_symbols=131, n_channels=80, hparams=cfg, stats=stats, n_speakers=6).eval()
        m.variance_adaptor.duration_predictor.linear.bias.data.fill_(math.log(7.0))
        g = torch.Generator().manual_seed(0)
        B, T = 3, 24
        text = torch.randint(1, 131, (B, T), generator=g)
        tl = torch.tensor([24, 17, 9])
        text[torch.arange(T)[None] >= tl[:, None]] = 0
        sids = torch.randint(0, 6, (B,), generator=g)
        cap = {}

        def pre(mod, args):
            cap["args"] = [a.clone() if torch.is_tensor(a) else a for a in args]

        def post(mod, args, res):
            cap["out"] = res.clone()
            cap["ds_after"] = args[1].clone()

        reg = m.variance_adaptor.length_regulator
        h1, h2 = reg.register_forward_pre_hook(pre), reg.register_forward_hook(post)
        with torch.no_grad():
            mel, feats_lengths, _ = m.inference(sids, text, tl)
        h1.remove(); h2.remove()
        out[f"{tag}.xs"] = cap["args"][0].numpy()
        out[f"{tag}.ds"] = cap["args"][1].numpy()
        out[f"{tag}.ds_after"] = cap["ds_after"].numpy()
        if use_gaussian:
            out[f"{tag}.h_masks"] = cap["args"][2].numpy()
            out[f"{tag}.d_masks"] = cap["args"][3].numpy()
        out[f"{tag}.out"] = cap["out"].numpy()
        out[f"{tag}.feats_lengths"] = feats_lengths.numpy()
        out[f"{tag}.mel"] = mel.contiguous().numpy()
        if not use_gaussian:
            torch.manual_seed(1234)
            voc = HiFiGAN().eval()
            with torch.no_grad():
                wav = voc(mel)
            out[f"{tag}.wav"] = wav.numpy()
        print("fastspeech2_capture", tag, tuple(cap["out"].shape), tuple(mel.shape), feats_lengths.tolist())
    np.savez_compressed(os.path.join(OUT, "fastspeech2_capture.npz"), **out)


def hifigan_v1_long():
    """One full-size comparison for the tile schedule at large L: the reference HiFiGAN V1 (seed 1234) on (4, 80, 1000)
    (BASELINE config C4's length).  Only a strided sample of the waveform is stored (every 97th sample + the first and
    last 2,048 of every row); the input is re-drawn from its seed by the test."""
    HiFiGAN, _ = ref_loader.load_hifigan()
    torch.manual_seed(1234)
    m = HiFiGAN().eval()
    g = torch.Generator().manual_seed(5)
    c = torch.randn(4, 80, 1000, generator=g)
    with torch.no_grad():
        y = m(c)
    L = y.shape[-1]
    idx = np.unique(np.concatenate([np.arange(0, 2048), np.arange(0, L, 97), np.arange(L - 2048, L)]))
    np.savez_compressed(os.path.join(OUT, "hifigan_v1_long.npz"), seed=np.int64(5), shape=np.array(c.shape), idx=idx,
                        y=y[:, 0, idx].numpy(), y_norm=np.float64(y.double().norm()), c_sum=np.float64(c.double().sum()))
    print("hifigan_v1_long", tuple(y.shape), idx.size)


def _randomize_batchnorm(postnet):
    """Non-trivial running statistics / affine parameters, as a trained Postnet has (deterministic formula, also used by the tests)."""
    for i, seq in enumerate(postnet.convolutions):
        bn = seq[1]
        n = bn.num_features
        t = torch.arange(n, dtype=torch.float32)
        bn.running_mean.copy_(0.2 * torch.sin(0.37 * t + i))
        bn.running_var.copy_(1.0 + 0.5 * torch.cos(0.11 * t + 2 * i))
        bn.weight.data.copy_(1.0 + 0.3 * torch.sin(0.05 * t + 3 * i))
        bn.bias.data.copy_(0.1 * torch.cos(0.23 * t + i))


def acoustic_tail():
    """f3: the reference's Decoder (transformer FFT blocks) + feats_linear + Postnet, built stand-alone from the reference's own
    classes.  `small`: full state dict stored.  `c2`: the C2 width (4 layers, 256 hidden, 1024 filter, Postnet 512) drawn with
    torch.manual_seed(1234) in a fixed construction order (decoder, feats_linear, postnet) -- the test re-draws the same
    parameters through the drop-in classes (construction-order parity) and only the checksums + outputs are stored."""
    ref_loader.load_length_regulator()
    from models.tts.fastspeech2.blocks.transformer import Decoder  # type: ignore
    from models.tts.fastspeech2.layers import Postnet  # type: ignore

    out = {}
    g = torch.Generator().manual_seed(9)
    for tag, (layers, hidden, filt, emb, B, T) in (("small", (2, 64, 128, 64, 3, 37)), ("c2", (4, 256, 1024, 512, 2, 50))):
        cfg = {"decoder_head": 2, "conv_filter_size": filt, "conv_kernel_size": [9, 1], "decoder_dropout": 0.2}
        torch.manual_seed(1234)
        dec = Decoder(layers, hidden, 1000, cfg).eval()
        lin = torch.nn.Linear(hidden, 80).eval()
        post = Postnet(80, {"embedding_dim": emb, "conv_layers": 5, "kernel_size": 5}).eval()
        with torch.no_grad():
            _randomize_batchnorm(post)
        frames = torch.randn(B, T, hidden, generator=g)
        mel_len = torch.tensor([T, T - 9, 5][:B])
        mask = torch.arange(T)[None] >= mel_len[:, None]
        frames = frames.masked_fill(mask.unsqueeze(-1), 0.0)      # the regulator pads with zeros
        with torch.no_grad():
            hs, _ = dec(frames, mask)
            outs = lin(hs)
            mel = (post(outs) + outs).transpose(1, 2)
            pn = post(outs)
        out.update({f"{tag}.frames": frames.numpy(), f"{tag}.mel_len": mel_len.numpy(), f"{tag}.dec": hs.numpy(),
                    f"{tag}.postnet": pn.numpy(), f"{tag}.mel": mel.contiguous().numpy()})
        sd = {}
        sd.update({"decoder." + k: v for k, v in dec.state_dict().items()})
        sd.update({"feats_linear." + k: v for k, v in lin.state_dict().items()})
        sd.update({"postnet." + k: v for k, v in post.state_dict().items()})
        keys = sorted(sd)
        out[f"{tag}.keys"] = np.array(keys)
        out[f"{tag}.sums"] = np.array([float(sd[k].double().sum()) for k in keys])
        if tag == "small":
            out.update(sd_to_np({k: v for k, v in sd.items() if k != "decoder.position_enc"}, prefix="small.sd."))
        print("acoustic_tail", tag, tuple(mel.shape), len(keys))
    np.savez_compressed(os.path.join(OUT, "acoustic_tail.npz"), **out)


def conformer_decoder():
    """f3, conformer variant: the reference's conformer Decoder (blocks/conformer.py:93-169) built stand-alone.  `small`: full
    state dict stored (without the position tables); `c4`: the shipped size (6 layers, 384 hidden, 8 heads, k = 31,
    model_config.yaml:3-6,25-32) drawn with torch.manual_seed(1234): checksums + outputs only."""
    ref_loader.load_length_regulator()
    from models.tts.fastspeech2.blocks.conformer import Decoder  # type: ignore

    out = {}
    g = torch.Generator().manual_seed(21)
    for tag, (layers, hidden, heads, B, T) in (("small", (2, 64, 2, 3, 45)), ("c4", (6, 384, 8, 2, 70))):
        cfg = {"decoder_head": heads, "ffn_expansion_factor": 4, "conv_expansion_factor": 2, "conv_kernel_size": 31,
               "half_step_residual": True, "decoder_dropout": 0.1}
        torch.manual_seed(1234)
        dec = Decoder(layers, hidden, 1000, cfg).eval()
        with torch.no_grad():
            for i, blk in enumerate(dec.layer_stack):            # non-trivial BatchNorm statistics, as after training
                bn = blk.sequential[2].module.sequential[5]
                t = torch.arange(bn.num_features, dtype=torch.float32)
                bn.running_mean.copy_(0.2 * torch.sin(0.37 * t + i)); bn.running_var.copy_(1.0 + 0.5 * torch.cos(0.11 * t + 2 * i))
                bn.weight.data.copy_(1.0 + 0.3 * torch.sin(0.05 * t + 3 * i)); bn.bias.data.copy_(0.1 * torch.cos(0.23 * t + i))
        frames = torch.randn(B, T, hidden, generator=g)
        mel_len = torch.tensor([T, T - 13, 9][:B])
        mask = torch.arange(T)[None] >= mel_len[:, None]
        frames = frames.masked_fill(mask.unsqueeze(-1), 0.0)
        with torch.no_grad():
            hs, _ = dec(frames, mask)
        sd = dec.state_dict()
        keys = sorted(sd)
        out.update({f"{tag}.frames": frames.numpy(), f"{tag}.mel_len": mel_len.numpy(), f"{tag}.dec": hs.numpy(),
                    f"{tag}.keys": np.array(keys), f"{tag}.sums": np.array([float(sd[k].double().sum()) for k in keys])})
        if tag == "small":
            out.update(sd_to_np({k: v for k, v in sd.items() if "position" not in k}, prefix="small.sd."))
        print("conformer_decoder", tag, tuple(hs.shape), len(keys))
    np.savez_compressed(os.path.join(OUT, "conformer_decoder.npz"), **out)


if __name__ == "__main__":
    assert ref_loader.reference_available(), "needs /root/reference"
    only = sys.argv[1:]
    if only:
        for name in only:
            globals()[name]()
        sys.exit(0)
    path_cases()
    gaussian_cases()
    lr_cases()
    hifigan_small()
    hifigan_v1()
    vits2_small()
    fastspeech2_capture()
    hifigan_v1_long()
    acoustic_tail()
    conformer_decoder()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))

"""f3: acoustic decoder (FFT blocks) + Postnet between the LengthRegulator and the generator.

Golden: ``tests/golden/acoustic_tail.npz``, produced by ``make_golden.py::acoustic_tail`` from the reference's own
``Decoder`` (blocks/transformer.py:90-166), ``nn.Linear`` and ``Postnet`` (layers.py:571-625) in eval mode.
CPU tests: state-dict keys / seeded parameter draw of the drop-in classes, the oracle restatement, the drop-ins'
autograd (PyTorch) path.  GPU tests: the conv-kernel path (fp16 operands, fp32 accumulate) against the golden.
Tolerance of the 16-bit path (the north_star's waveform bar applied to the mel): rel-L2 <= 1e-3, max-abs <= 1e-2 against
the fp32 reference itself, not a rounded emulation (measured: decoder 1.8e-4, Postnet 7e-4, mel 6.5e-4 / 3.6e-3).
"""
import numpy as np
import pytest
import torch

import restate
import vtts_b200
from conftest import load_golden, max_abs, rel_l2, state_dict_from

DEV = "cuda:0"
SHAPES = {"small": (2, 64, 128, 64), "c2": (4, 256, 1024, 512)}


def randomize_batchnorm(postnet):
    """Same deterministic formula as make_golden.py::_randomize_batchnorm."""
    with torch.no_grad():
        for i, seq in enumerate(postnet.convolutions):
            bn = seq[1]
            t = torch.arange(bn.num_features, dtype=torch.float32)
            bn.running_mean.copy_(0.2 * torch.sin(0.37 * t + i))
            bn.running_var.copy_(1.0 + 0.5 * torch.cos(0.11 * t + 2 * i))
            bn.weight.data.copy_(1.0 + 0.3 * torch.sin(0.05 * t + 3 * i))
            bn.bias.data.copy_(0.1 * torch.cos(0.23 * t + i))


def build_tail(tag):
    layers, hidden, filt, emb = SHAPES[tag]
    cfg = {"decoder_head": 2, "conv_filter_size": filt, "conv_kernel_size": [9, 1], "decoder_dropout": 0.2}
    torch.manual_seed(1234)
    dec = vtts_b200.Decoder(layers, hidden, 1000, cfg)
    lin = torch.nn.Linear(hidden, 80)
    post = vtts_b200.Postnet(80, {"embedding_dim": emb, "conv_layers": 5, "kernel_size": 5})
    randomize_batchnorm(post)
    return vtts_b200.AcousticTail(dec, lin, post).eval()


@pytest.fixture(scope="module")
def gold():
    return load_golden("acoustic_tail.npz")


@pytest.mark.parametrize("tag", ["small", "c2"])
def test_keys_and_seeded_draw_match_the_reference_classes(gold, tag):
    m = build_tail(tag)
    sd = m.state_dict()
    assert sorted(sd) == [str(k) for k in gold[f"{tag}.keys"]]
    sums = np.array([float(sd[k].double().sum()) for k in sorted(sd)])
    assert np.allclose(sums, gold[f"{tag}.sums"], rtol=0, atol=1e-6 * np.maximum(1.0, np.abs(gold[f"{tag}.sums"])))


def test_oracle_matches_the_reference_on_the_stored_state_dict(gold):
    sd = state_dict_from(gold, prefix="small.sd.")
    sd["decoder.position_enc"] = restate.sinusoid_table(1001, 64).unsqueeze(0)
    frames, mel_len = torch.from_numpy(gold["small.frames"]), torch.from_numpy(gold["small.mel_len"])
    mask = torch.arange(frames.shape[1])[None] >= mel_len[:, None]
    dec = restate.fft_decoder_forward(sd, frames, mask, 2, prefix="decoder.")
    assert max_abs(dec, torch.from_numpy(gold["small.dec"])) < 1e-5
    mel = restate.acoustic_tail_forward(sd, frames, mel_len, 2)
    assert max_abs(mel, torch.from_numpy(gold["small.mel"])) < 1e-5


def test_autograd_path_equals_the_reference_and_loads_its_state_dict(gold):
    m = build_tail("small")
    sd = state_dict_from(gold, prefix="small.sd.")
    sd["decoder.position_enc"] = m.decoder.position_enc.detach().clone()
    assert not m.load_state_dict(sd, strict=True).missing_keys
    frames, mel_len = torch.from_numpy(gold["small.frames"]), torch.from_numpy(gold["small.mel_len"])
    mel = m(frames, mel_len)                    # grad mode + parameters that require grad -> PyTorch formula
    assert mel.requires_grad and max_abs(mel, torch.from_numpy(gold["small.mel"])) < 1e-5
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(frames, mel_len)                      # synthesis path: no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "c2"])
def test_conv_kernel_path_against_the_reference_golden(gold, tag):
    m = build_tail(tag).to(DEV)
    frames, mel_len = torch.from_numpy(gold[f"{tag}.frames"]).to(DEV), torch.from_numpy(gold[f"{tag}.mel_len"]).to(DEV)
    mask = torch.arange(frames.shape[1], device=DEV)[None] >= mel_len[:, None]
    with torch.no_grad():
        dec, _ = m.decoder(frames, mask)
        outs = torch.from_numpy(gold[f"{tag}.mel"]).to(DEV).transpose(1, 2) - torch.from_numpy(gold[f"{tag}.postnet"]).to(DEV)
        pn = m.postnet(outs)                   # the Postnet alone on the reference's own pre-Postnet mel
        mel = m(frames, mel_len)
    for name, got, ref in (("decoder", dec, gold[f"{tag}.dec"]), ("postnet", pn, gold[f"{tag}.postnet"]), ("mel", mel, gold[f"{tag}.mel"])):
        ref = torch.from_numpy(ref)
        r, a = rel_l2(got, ref), max_abs(got, ref)
        print(f"{tag} {name}: rel-L2 {r:.3e} max-abs {a:.3e}")
        assert r <= 1e-3 and a <= 1e-2, (tag, name, r, a)


@pytest.mark.gpu
def test_single_conv_layer_entry_point():
    """vtts_conv_forward: bias, residual, ReLU / tanh 16-bit copies, cout that is not a multiple of 128, cin padding."""
    from vtts_b200.acoustic import _TcConv, _operand
    g = torch.Generator().manual_seed(0)
    for cin, cout, k, L in ((80, 512, 5, 203), (256, 1024, 9, 300), (1024, 256, 1, 97), (512, 80, 5, 64)):
        conv = torch.nn.Conv1d(cin, cout, k, padding=(k - 1) // 2)
        x = torch.randn(2, L, cin, generator=g)
        res = torch.randn(2, L, cout, generator=g)
        tc = _TcConv(conv.to(DEV))
        a = _operand(x.to(DEV), "fp16", tc.padded_channels(torch.device(DEV)))
        want_a = cout % 32 == 0
        y, ya = tc.run(a, "fp16", want_x=True, want_a=False, res=res.to(DEV).contiguous())
        ref = torch.nn.functional.conv1d(x.half().float().transpose(1, 2).double(), conv.weight.detach().cpu().half().double(),
                                         conv.bias.detach().cpu().double(), padding=(k - 1) // 2).transpose(1, 2).float() + res
        assert max_abs(y, ref) < 2e-3 * max(1.0, float(ref.abs().max())), (cin, cout, k)
        if want_a:
            _, yt = tc.run(a, "fp16", want_x=False, want_a=True, act_tanh=True)
            _, yr = tc.run(a, "fp16", want_x=False, want_a=True, slope_out=0.0)
            assert max_abs(yt.float(), torch.tanh(ref - res)) < 2e-3 and max_abs(yr.float(), torch.relu(ref - res)) < 4e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.gpu
def test_synthesizer_with_the_acoustic_tail(gold):
    """regulator -> decoder + Postnet (conv kernels) -> vocoder through the front door; oracle composition as the checker."""
    tail = build_tail("c2").to(DEV)
    torch.manual_seed(1234)
    voc = vtts_b200.HiFiGAN().to(DEV).eval()
    g = torch.Generator().manual_seed(4)
    hs = torch.randn(2, 9, 256, generator=g)
    ds = torch.tensor([[3, 2, 4, 1, 5, 2, 3, 4, 2], [2, 3, 1, 4, 0, 0, 0, 0, 0]])
    synth = vtts_b200.Synthesizer(voc, vtts_b200.LengthRegulator(), acoustic_tail=tail)
    wav, wav_len = synth(hs, ds)
    frames, _ = restate.lr_expand(hs, ds.clone())
    sd_t = {k: v.detach().cpu() for k, v in tail.state_dict().items()}
    mel = restate.acoustic_tail_forward(sd_t, frames, ds.sum(1), 2)
    ref = restate.hifigan_forward({k: v.detach().cpu() for k, v in voc.state_dict().items()}, mel)
    for b, n in enumerate(wav_len.tolist()):
        assert rel_l2(wav[b, 0, :n], ref[b, 0, :n]) <= 5e-3 and max_abs(wav[b, 0, :n], ref[b, 0, :n]) <= 2e-2

"""a9: the hot path on tensors captured INSIDE the reference's ``FastSpeech2.inference``.

``tests/golden/fastspeech2_capture.npz`` (made by ``make_golden.py::fastspeech2_capture`` from the unmodified
reference: fastspeech2/model.py:194-257 -> VarianceAdaptor.forward layers.py:226-233 -> decoder / Postnet ->
the reference HiFiGAN, text2wav/model.py:139-167) holds what the regulator really receives in that call (hidden
states after the pitch / energy embeddings, ``duration_rounded``, the masks), what it returned, the final mel and
the reference vocoder's waveform for that mel.  The GPU tests replay those tensors through the CUDA kernels: the
frame expansion must be bit-exact, the Gaussian upsampling within 1e-5, the fp16 waveform within the north_star
tolerance on the model's own mel.  The CPU tests pin the oracle on the same tensors and, where ``/root/reference``
exists, check that ``vtts_b200.install()`` really lands inside a reference-built ``FastSpeech2``.
"""
import numpy as np
import pytest
import torch

import ref_loader
import restate
import vtts_b200
from conftest import load_golden, max_abs, rel_l2, split_cases

DEV = "cuda:0"


@pytest.fixture(scope="module")
def cap():
    return split_cases(load_golden("fastspeech2_capture.npz"))


# ---- CPU: the oracle on the captured tensors --------------------------------------------------------
def test_oracle_expansion_matches_the_captured_call(cap):
    c = cap["lr"]
    ds = torch.from_numpy(c["ds"]).clone()
    out, ds_used = restate.lr_expand(torch.from_numpy(c["xs"]), ds)
    assert torch.equal(out, torch.from_numpy(c["out"]))
    assert torch.equal(restate.lr_mel_len(ds_used), torch.from_numpy(c["feats_lengths"]))
    assert torch.equal(ds, torch.from_numpy(c["ds_after"]))


def test_oracle_vocoder_matches_the_captured_waveform(cap):
    c = cap["lr"]
    torch.manual_seed(1234)                                  # the shell draws the reference's weights (test_modules_cpu.py)
    sd = {k: v.detach() for k, v in vtts_b200.HiFiGAN().state_dict().items()}
    with torch.no_grad():
        y = restate.hifigan_forward(sd, torch.from_numpy(c["mel"][2:3]))   # the shortest row keeps this quick
    ref = torch.from_numpy(c["wav"][2:3])
    assert y.shape == ref.shape
    assert rel_l2(y, ref) < 1e-5 and max_abs(y, ref) < 1e-5


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference (build container only)")
def test_install_lands_inside_a_reference_built_fastspeech2():
    """models/tts/fastspeech2/layers.py:168-172 builds the regulator from the module-level names install() rebinds."""
    import yaml

    ref_loader.load_length_regulator()
    import models.tts.fastspeech2.layers as ref_layers  # type: ignore
    from models.tts.fastspeech2.model import FastSpeech2  # type: ignore

    saved = (ref_layers.LengthRegulator, ref_layers.GaussianUpsampling)
    try:
        vtts_b200.install(import_missing=False)
        for use_gaussian, cls in ((False, vtts_b200.LengthRegulator), (True, vtts_b200.GaussianUpsampling)):
            with open(ref_loader.REF_ROOT + "/config/model_config.yaml") as fh:
                cfg = yaml.safe_load(fh)["fastspeech2"]
            cfg["use_cvae"] = False
            cfg["encoder_layers"] = cfg["decoder_layers"] = 1
            cfg["encoder_hidden"] = cfg["decoder_hidden"] = 32
            cfg["building_block"]["block_type"] = "transformer"
            cfg["variance"]["learn_alignment"] = False
            cfg["variance"]["duration_modelling"]["use_gaussian"] = use_gaussian
            stats = {"pitch": {"min": -2.0, "max": 8.0}, "energy": {"min": -1.5, "max": 7.0}}
            m = FastSpeech2(n_symbols=131, n_channels=80, hparams=cfg, stats=stats, n_speakers=6).eval()
            assert type(m.variance_adaptor.length_regulator) is cls
            # the synthesis path has no CPU fallback: the reference wrapper reaches the drop-in and it refuses loudly
            with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
                m.inference(torch.zeros(1, dtype=torch.long), torch.randint(1, 131, (1, 5)), torch.tensor([5]))
    finally:
        vtts_b200.uninstall()
        assert (ref_layers.LengthRegulator, ref_layers.GaussianUpsampling) == saved


# ---- GPU: the kernels on the captured tensors -------------------------------------------------------
@pytest.mark.gpu
def test_expansion_bit_exact_on_the_models_own_hidden_states(cap):
    c = cap["lr"]
    xs, ds = torch.from_numpy(c["xs"]).to(DEV), torch.from_numpy(c["ds"]).to(DEV)
    out, mel_len = vtts_b200.LengthRegulator().forward_with_lengths(xs, ds)
    assert torch.equal(out.cpu(), torch.from_numpy(c["out"]))
    assert torch.equal(mel_len.cpu(), torch.from_numpy(c["feats_lengths"]))
    assert torch.equal(ds.cpu(), torch.from_numpy(c["ds_after"]))


@pytest.mark.gpu
def test_gaussian_upsampling_on_the_models_own_call(cap):
    c = cap["gauss"]
    hs, ds = torch.from_numpy(c["xs"]).to(DEV), torch.from_numpy(c["ds"]).to(DEV)
    hm, dm = torch.from_numpy(c["h_masks"]).to(DEV), torch.from_numpy(c["d_masks"]).to(DEV)
    with torch.no_grad():
        y = vtts_b200.GaussianUpsampling()(hs, ds, hm, dm)
    assert max_abs(y, torch.from_numpy(c["out"])) <= 1e-5 * max(1.0, float(np.abs(c["out"]).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("precision,rel_tol,abs_tol", [("fp32", 2e-5, 2e-5), ("fp16", 1e-3, 1e-2)])
def test_waveform_on_the_models_own_mel(cap, precision, rel_tol, abs_tol):
    """north_star tolerance (rel-L2 <= 1e-3, max-abs <= 1e-2 for the 16-bit path) on a mel FastSpeech2 produced."""
    c = cap["lr"]
    torch.manual_seed(1234)
    m = vtts_b200.HiFiGAN()
    m.precision = precision
    m = m.to(DEV).eval()
    mel = torch.from_numpy(c["mel"]).to(DEV)
    ref = torch.from_numpy(c["wav"])
    with torch.no_grad():
        y = m(mel)
        yt = m.forward_trimmed(mel, torch.from_numpy(c["feats_lengths"]).to(DEV))
    assert rel_l2(y, ref) <= rel_tol and max_abs(y, ref) <= abs_tol
    for b, n in enumerate(c["feats_lengths"].tolist()):           # padding trim: valid samples identical
        assert torch.equal(yt[b, :, : n * 256], y[b, :, : n * 256])


@pytest.mark.gpu
def test_synthesizer_front_door_on_the_captured_call(cap):
    """regulator -> (mel as the reference decoder produced it) -> vocoder through Synthesizer, host buffers in / out."""
    c = cap["lr"]
    torch.manual_seed(1234)
    m = vtts_b200.HiFiGAN().to(DEV).eval()
    mel_ref = torch.from_numpy(c["mel"]).to(DEV)                  # (B, 80, T): stands in for decoder + Postnet

    def frames_to_mel(frames):
        assert torch.equal(frames.cpu(), torch.from_numpy(c["out"]))
        return mel_ref

    synth = vtts_b200.Synthesizer(m, vtts_b200.LengthRegulator(), frames_to_mel=frames_to_mel)
    wav, wav_len = synth(torch.from_numpy(c["xs"]), torch.from_numpy(c["ds"]))
    assert wav_len.tolist() == [int(n) * 256 for n in c["feats_lengths"]]
    ref = torch.from_numpy(c["wav"])
    for b, n in enumerate(wav_len.tolist()):
        got = wav[b].reshape(-1)[:n]
        assert rel_l2(got, ref[b, 0, :n]) <= 1e-3 and max_abs(got, ref[b, 0, :n]) <= 1e-2

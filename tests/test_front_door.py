"""f4 front door: phoneme ids -> per-utterance PCM (vtts_b200.TwoStageTTS / OneStageTTS), the role of the
``OneStageTTS / TwoStageTTS`` objects ``test.py:36-38,63`` drives (their module is absent upstream)."""
import wave

import numpy as np
import pytest
import torch

import restate
import vtts_b200
from conftest import max_abs, rel_l2


class ToyAcoustic(torch.nn.Module):
    """inference(sids, text, text_lengths) -> (mel (B, 80, T), mel_len): embedding -> fixed durations -> regulator."""

    def __init__(self, regulator):
        super().__init__()
        torch.manual_seed(5)
        self.emb = torch.nn.Embedding(32, 80, padding_idx=0)
        self.spk = torch.nn.Embedding(4, 80)
        self.length_regulator = regulator

    def durations(self, text):
        return (text % 3 + 1) * (text > 0)

    def inference(self, sids, text, text_lengths):
        hs = self.emb(text) + self.spk(sids)[:, None, :] * (text > 0)[..., None]
        ds = self.durations(text)
        frames = self.length_regulator(hs, ds)
        return frames.transpose(1, 2).contiguous(), ds.sum(1), None


class CpuRegulator(torch.nn.Module):
    def forward(self, xs, ds):
        return restate.lr_expand(xs, ds.clone())[0]


class RecordingVocoder(torch.nn.Module):
    upsample_factor = 4

    def forward(self, mel):
        return mel[:, :1, :].repeat_interleave(4, dim=2)


def test_order_batching_and_wav_writer(tmp_path):
    tts = vtts_b200.TwoStageTTS(ToyAcoustic(CpuRegulator()), RecordingVocoder(), max_batch=2)
    texts = [torch.tensor([3, 5, 7]), torch.tensor([2, 2, 2, 9, 4, 1]), torch.tensor([11]), torch.tensor([6, 6])]
    outs = tts(texts, speaker_id=1)
    assert len(outs) == 4
    single = [tts([t], speaker_id=1)[0] for t in texts]          # batching and ordering do not change any utterance
    for a, b, t in zip(outs, single, texts):
        assert a.dtype == np.float32 and a.shape == b.shape and np.array_equal(a, b)
        assert a.shape[0] == int(((t % 3 + 1) * (t > 0)).sum()) * 4
    path = str(tmp_path / "a.wav")
    vtts_b200.save_wav(path, outs[1], 22050)
    with wave.open(path, "rb") as fh:
        assert fh.getnchannels() == 1 and fh.getframerate() == 22050 and fh.getnframes() == outs[1].shape[0]


def test_one_stage_shape():
    class Joint(torch.nn.Module):
        def inference(self, sids, text, text_lengths):
            return text.float().unsqueeze(1).repeat_interleave(2, dim=2), text_lengths * 2

    outs = vtts_b200.OneStageTTS(Joint())([torch.tensor([1, 2, 3]), torch.tensor([4])])
    assert [o.tolist() for o in outs] == [[1, 1, 2, 2, 3, 3], [4, 4]]


@pytest.mark.gpu
def test_two_stage_on_the_kernels_against_the_oracle():
    dev = "cuda:0"
    ac = ToyAcoustic(vtts_b200.LengthRegulator()).to(dev)
    torch.manual_seed(1234)
    voc = vtts_b200.HiFiGAN().to(dev).eval()
    tts = vtts_b200.TwoStageTTS(ac, voc, device=dev, max_batch=3)
    g = torch.Generator().manual_seed(0)
    texts = [torch.randint(1, 32, (int(n),), generator=g) for n in (9, 4, 13, 6, 2)]
    outs = tts(texts, speaker_id=2)
    sd = {k: v.detach().cpu() for k, v in voc.state_dict().items()}
    ac_cpu = ToyAcoustic(CpuRegulator())
    # oracle on the same padded batches (like Text2Wav.inference, text2wav/model.py:139-167: the generator sees the padded
    # mel and the waveform is cut afterwards, so the last frames of a short utterance depend on its batch)
    from vtts_b200.tts import _pad_batch
    for idx in tts._batches(texts):
        text, lens = _pad_batch([texts[i] for i in idx])
        mel, mel_len, _ = ac_cpu.inference(torch.full((len(idx),), 2), text, lens)
        with torch.no_grad():
            ref = restate.hifigan_forward(sd, mel)
        for row, i in enumerate(idx):
            n = int(mel_len[row]) * 256
            got = torch.from_numpy(outs[i])
            assert got.shape[0] == n
            assert rel_l2(got, ref[row, 0, :n]) <= 1e-3 and max_abs(got, ref[row, 0, :n]) <= 1e-2

"""Synthesis front door: phoneme ids in, PCM out.

The reference's ``test.py:6,36-38,63`` drives ``OneStageTTS / TwoStageTTS`` from ``src/api/modules/tts.py`` -- a module
the upstream repository does not ship -- as ``nnet(texts=..., speaker_id=...)`` followed by a wav writer (``test.py:7,73``).
Text normalisation and G2P are remote services there (``test.py:51-56``) and out of scope here; this wrapper covers the
part that runs on the GPU: batched phoneme-id sequences -> acoustic model ``inference`` -> vocoder -> per-utterance PCM,
with the same two entry shapes:

  ``TwoStageTTS(acoustic, vocoder)``   acoustic.inference(sids, text, text_lengths) -> (mel (B, n_mel, T), mel_len, ...)
                                       (FastSpeech2.inference, fastspeech2/model.py:194-257), then the vocoder
                                       (Text2Wav.inference order, text2wav/model.py:139-167)
  ``OneStageTTS(model)``               model.inference(sids, text, text_lengths) -> (wav, wav_len)   (Text2Wav / JETS / VITS2)

Both take already-built modules (a reference model after ``vtts_b200.install()``, or the drop-in shells) instead of
checkpoint paths, batch the utterances longest-first, and return one float32 numpy array per input in the caller's order.
"""
from __future__ import annotations

import wave
from typing import List, Optional, Sequence

import numpy as np
import torch


def _pad_batch(seqs: Sequence[torch.Tensor], pad_id: int = 0):
    lens = torch.tensor([int(s.numel()) for s in seqs], dtype=torch.long)
    out = torch.full((len(seqs), int(lens.max())), pad_id, dtype=torch.long)
    for i, s in enumerate(seqs):
        out[i, : s.numel()] = s.to(torch.long).view(-1)
    return out, lens


class _FrontDoor:
    sample_rate = 22050

    def __init__(self, device: Optional[torch.device] = None, max_batch: int = 16, speakers: Optional[Sequence] = None):
        self.device = torch.device(device) if device is not None else None
        self.max_batch = max_batch
        self.speakers = list(speakers) if speakers is not None else [0]
        self.accents = None                      # test.py:58 branches on this attribute

    def _batches(self, seqs):
        order = sorted(range(len(seqs)), key=lambda i: -int(seqs[i].numel()))      # longest first: least padding per batch
        for k in range(0, len(order), self.max_batch):
            yield order[k: k + self.max_batch]

    def _synthesize(self, sids, text, text_lengths):
        raise NotImplementedError

    @torch.no_grad()
    def __call__(self, texts: Sequence[torch.Tensor], speaker_id: int = 0) -> List[np.ndarray]:
        """texts: phoneme-id sequences (1-D LongTensors).  Returns float32 waveforms in [-1, 1], one per input."""
        if len(texts) == 0:
            return []
        out: List[Optional[np.ndarray]] = [None] * len(texts)
        for idx in self._batches(texts):
            text, lens = _pad_batch([texts[i] for i in idx])
            sids = torch.full((len(idx),), int(speaker_id), dtype=torch.long)
            if self.device is not None:
                text, lens, sids = text.to(self.device), lens.to(self.device), sids.to(self.device)
            wav, wav_len = self._synthesize(sids, text, lens)
            wav = wav.reshape(wav.shape[0], -1).float().cpu().numpy()
            wav_len = wav_len.cpu().tolist()
            for row, i in enumerate(idx):
                out[i] = wav[row, : int(wav_len[row])].copy()
        return out  # type: ignore[return-value]


class TwoStageTTS(_FrontDoor):
    def __init__(self, acoustic: torch.nn.Module, vocoder: torch.nn.Module, **kw):
        super().__init__(**kw)
        self.acoustic, self.vocoder = acoustic.eval(), vocoder.eval()
        self.hop = int(getattr(vocoder, "upsample_factor", 256))

    def _synthesize(self, sids, text, text_lengths):
        mel, mel_len = self.acoustic.inference(sids, text, text_lengths)[:2]
        if hasattr(self.vocoder, "forward_trimmed"):
            wav = self.vocoder.forward_trimmed(mel, mel_len)            # skip the padded tail of shorter utterances
        else:
            wav = self.vocoder(mel)
        return wav, mel_len * self.hop


class OneStageTTS(_FrontDoor):
    def __init__(self, model: torch.nn.Module, **kw):
        super().__init__(**kw)
        self.model = model.eval()

    def _synthesize(self, sids, text, text_lengths):
        wav, wav_len = self.model.inference(sids, text, text_lengths)[:2]
        return wav, wav_len


def save_wav(path: str, pcm: np.ndarray, sample_rate: int = 22050) -> None:
    """16-bit mono RIFF writer (the reference's ``save_to_local`` encodes through an external service module)."""
    x = np.clip(np.asarray(pcm, dtype=np.float32), -1.0, 1.0)
    with wave.open(path, "wb") as fh:
        fh.setnchannels(1)
        fh.setsampwidth(2)
        fh.setframerate(sample_rate)
        fh.writeframes((x * 32767.0).round().astype("<i2").tobytes())

"""``Synthesizer`` -- the call a user makes: host buffers in, waveforms out.

Plays the role of the reference's ``Text2Wav.inference`` tail
(models/gan_tts/text2wav/model.py:139-167) for the part of the path this repository owns:
token-level hidden states + integer durations -> LengthRegulator -> (acoustic decoder, which
stays PyTorch and is supplied by the caller as ``frames_to_mel``) -> HiFi-GAN generator ->
``(wav, wav_len)`` with ``wav_len = mel_len * upsample_factor`` (text2wav/model.py:165).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from .length_regulator import LengthRegulator


def slice_decoder_stand_in(in_channels: int) -> Callable[[torch.Tensor], torch.Tensor]:
    """Stand-in for the acoustic decoder (out of scope): first ``in_channels`` features as mel."""

    def f(frames: torch.Tensor) -> torch.Tensor:  # (B, T, D) -> (B, in_channels, T)
        return frames[..., :in_channels].transpose(1, 2)

    return f


class Synthesizer:
    def __init__(self, generator, length_regulator: Optional[LengthRegulator] = None,
                 frames_to_mel: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 device: Optional[torch.device] = None, trim_padding: bool = True):
        self.generator = generator
        self.length_regulator = length_regulator or LengthRegulator()
        cfg = generator._gen_config()
        self.frames_to_mel = frames_to_mel or slice_decoder_stand_in(cfg.in_channels)
        self.device = torch.device(device) if device is not None else next(generator.parameters()).device
        self._pinned_out = None
        # skip generator work on the padded tail of shorter utterances (valid samples are unaffected)
        self.trim_padding = trim_padding

    @torch.no_grad()
    def __call__(self, hs: torch.Tensor, ds: torch.Tensor, alpha: float = 1.0, to_host: bool = True
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """hs (B,Tmax,D) float, ds (B,Tmax) int64 -- host (ideally pinned) or device tensors."""
        dev = self.device
        hs_d = hs.to(dev, non_blocking=True)
        ds_d = ds.to(dev, non_blocking=True)
        frames, mel_len = self.length_regulator.forward_with_lengths(hs_d, ds_d, alpha)
        mel = self.frames_to_mel(frames)
        if self.trim_padding and hasattr(self.generator, "forward_trimmed"):
            wav = self.generator.forward_trimmed(mel, mel_len)
        else:
            wav = self.generator(mel)
        wav_len = mel_len * self.generator.upsample_factor
        if not to_host:
            return wav, wav_len
        if self._pinned_out is None or self._pinned_out.numel() < wav.numel():
            self._pinned_out = torch.empty(wav.numel(), dtype=wav.dtype, pin_memory=True)
        out = self._pinned_out[: wav.numel()].view(wav.shape)
        out.copy_(wav, non_blocking=True)
        wav_len_h = wav_len.to("cpu")
        torch.cuda.current_stream(dev).synchronize()
        return out, wav_len_h

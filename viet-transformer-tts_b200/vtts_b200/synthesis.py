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


class PendingSynthesis:
    """Handle of one ``Synthesizer.submit`` call: the waveform is on its way to a pinned host buffer."""

    def __init__(self, wav_dev, wav_host, wav_len_dev, wav_len_host, copied):
        self._wav_dev, self._wav_len_dev = wav_dev, wav_len_dev  # kept alive until the copy has finished
        self._wav_host, self._wav_len_host, self._copied = wav_host, wav_len_host, copied

    def done(self) -> bool:
        return self._copied.query()

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(wav (B,1,T) pinned host tensor, wav_len (B,) host tensor).  The host buffer is reused by the submit
        after next (two buffers rotate): consume or copy it before submitting two more batches."""
        self._copied.synchronize()
        self._wav_dev = self._wav_len_dev = None
        return self._wav_host, self._wav_len_host


class Synthesizer:
    def __init__(self, generator, length_regulator: Optional[LengthRegulator] = None,
                 frames_to_mel: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 device: Optional[torch.device] = None, trim_padding: bool = True,
                 acoustic_tail: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
        """``acoustic_tail(frames (B,T,D), mel_len (B,)) -> mel (B, n_mel, T)`` (e.g. :class:`vtts_b200.AcousticTail`:
        decoder + Postnet on the conv kernels) takes precedence over ``frames_to_mel(frames)``; without either the
        first ``in_channels`` features stand in for the decoder."""
        self.generator = generator
        self.length_regulator = length_regulator or LengthRegulator()
        cfg = generator._gen_config()
        self.frames_to_mel = frames_to_mel or slice_decoder_stand_in(cfg.in_channels)
        self.acoustic_tail = acoustic_tail
        self.device = torch.device(device) if device is not None else next(generator.parameters()).device
        self._pinned_out = None
        self._copy_stream = None            # submit(): device->host copies run here, behind the next batch's kernels
        self._pinned_ring = [None, None]
        self._pinned_len_ring = [None, None]
        self._ring_pos = 0
        self._host_wav_len = None
        # skip generator work on the padded tail of shorter utterances (valid samples are unaffected)
        self.trim_padding = trim_padding

    @torch.no_grad()
    def __call__(self, hs: torch.Tensor, ds: torch.Tensor, alpha: float = 1.0, to_host: bool = True
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """hs (B,Tmax,D) float, ds (B,Tmax) int64 -- host (ideally pinned) or device tensors.

        With ``to_host`` (default) the returned waveform is a VIEW of one pinned host buffer that this object reuses:
        the next ``__call__`` overwrites it.  Consume it (write the file, ``.clone()``) before calling again, or use
        ``to_host=False`` for a device tensor you own; ``submit()`` rotates two buffers (see ``PendingSynthesis.result``)."""
        dev = self.device
        # Durations that arrive on the host are summed on the host (a few hundred integers): the output length is then
        # known without the LengthRegulator's device->host read, and the whole call is queued without a sync.
        max_len, host_len = None, None
        if not ds.is_cuda and ds.dtype == torch.int64 and ds.numel():
            ds_eff = ds if alpha == 1.0 else torch.round(ds.float() * alpha).long()   # layers.py:446-448
            sums = ds_eff.sum(1)
            if int(ds_eff.min()) >= 0 and int(sums.sum()) > 0:   # otherwise: the module's own error / fix-up path
                max_len, host_len = int(sums.max()), sums
        hs_d = hs.to(dev, non_blocking=True)
        ds_d = ds.to(dev, non_blocking=True)
        frames, mel_len = self.length_regulator.forward_with_lengths(hs_d, ds_d, alpha, max_len=max_len)
        mel = self.acoustic_tail(frames, mel_len) if self.acoustic_tail is not None else self.frames_to_mel(frames)
        if self.trim_padding and hasattr(self.generator, "forward_trimmed"):
            wav = self.generator.forward_trimmed(mel, mel_len)
        else:
            wav = self.generator(mel)
        wav_len = mel_len * self.generator.upsample_factor
        self._host_wav_len = None if host_len is None else host_len * self.generator.upsample_factor
        if not to_host:
            return wav, wav_len
        if self._pinned_out is None or self._pinned_out.numel() < wav.numel():
            self._pinned_out = torch.empty(wav.numel(), dtype=wav.dtype, pin_memory=True)
        out = self._pinned_out[: wav.numel()].view(wav.shape)
        out.copy_(wav, non_blocking=True)
        wav_len_h = self._host_wav_len if self._host_wav_len is not None else wav_len.to("cpu")
        torch.cuda.current_stream(dev).synchronize()
        return out, wav_len_h

    @torch.no_grad()
    def submit(self, hs: torch.Tensor, ds: torch.Tensor, alpha: float = 1.0) -> PendingSynthesis:
        """Throughput form of ``__call__``: returns as soon as the batch is queued; the device->host copy of its
        waveform runs on a second stream, so it overlaps the kernels of the NEXT submitted batch.

            pending = synth.submit(hs0, ds0)
            for hs, ds in batches:            # steady state: copy of batch k overlaps compute of batch k+1
                nxt = synth.submit(hs, ds)
                wav, wav_len = pending.result()
                ...
                pending = nxt
        """
        dev = self.device
        wav, wav_len = self(hs, ds, alpha, to_host=False)
        cur = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        i = self._ring_pos
        self._ring_pos ^= 1
        if self._pinned_ring[i] is None or self._pinned_ring[i].numel() < wav.numel():
            self._pinned_ring[i] = torch.empty(wav.numel(), dtype=wav.dtype, pin_memory=True)
        if self._pinned_len_ring[i] is None or self._pinned_len_ring[i].numel() < wav_len.numel():
            self._pinned_len_ring[i] = torch.empty(wav_len.numel(), dtype=wav_len.dtype, pin_memory=True)
        out = self._pinned_ring[i][: wav.numel()].view(wav.shape)
        out_len = self._pinned_len_ring[i][: wav_len.numel()].view(wav_len.shape)
        ready = torch.cuda.Event()
        ready.record(cur)
        copied = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            out.copy_(wav, non_blocking=True)
            out_len.copy_(wav_len, non_blocking=True)
            copied.record(self._copy_stream)
        wav.record_stream(self._copy_stream)
        wav_len.record_stream(self._copy_stream)
        return PendingSynthesis(wav, out, wav_len, out_len, copied)

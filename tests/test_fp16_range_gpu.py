"""fp16 operand range of the 16-bit path (weak point: ``cvt.rn.satfinite.f16x2`` clips silently at 65504).

Inputs so far were ``randn`` mels; a log-mel has mean about -5 and a wide one-sided range, and a trained checkpoint can
carry much larger gains than the random init.  These tests (1) run the tolerance check on log-mel statistics, (2) scale
the input conv so that stage-0 activations pass 1e4 and check both the tolerance and that the range probe
(``HiFiGAN.activation_range`` -> ``vtts_gen_set_range_probe``) reports the headroom, (3) push past 65504 and check
that the probe flags it (fp16 saturates there by construction; bf16 / fp32 are the documented way out).
"""
import pytest
import torch

import restate
import vtts_b200
from conftest import max_abs, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def log_mel_like(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, 80, T, generator=g) * 2.0 - 5.0
    tilt = torch.linspace(1.5, -2.5, 80).view(1, 80, 1)            # spectral tilt: low bins louder
    silence = (torch.rand(B, 1, T, generator=g) < 0.1).float() * -6.0
    return (base + tilt + silence).clamp(-11.5, 2.0)                # log(clip(mel, 1e-5)) floor = -11.5


def v1():
    torch.manual_seed(1234)
    return vtts_b200.HiFiGAN().to(DEV).eval()


def oracle_wave(m, c):
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        return restate.hifigan_forward(sd, c)


def test_log_mel_statistics_within_tolerance_and_range():
    m = v1()
    c = log_mel_like(2, 40, seed=3)
    ref = oracle_wave(m, c)
    with torch.no_grad():
        y = m(c.to(DEV))
    assert rel_l2(y, ref) <= 1e-3 and max_abs(y, ref) <= 1e-2, (rel_l2(y, ref), max_abs(y, ref))
    r = m.activation_range(c.to(DEV))
    assert abs(r["input"] - float(c.abs().max())) < 1e-5, (r["input"], float(c.abs().max()))
    # 78 conv layers; the output conv's result (the waveform) is not a 16-bit operand and is not probed
    assert len(r["layers"]) == 78 and sum(1 for v in r["layers"] if v == 0) == 1
    assert r["fp16_headroom"] > 100                                  # random-init V1 on a log-mel: far from 65504


def scaled(m, gain):
    """A checkpoint with large internal gains and the SAME waveform: LeakyReLU is positively homogeneous, so scaling the
    input conv and every later bias by `gain` and the output conv's weight by 1 / gain multiplies every internal
    activation by `gain` and leaves the output unchanged."""
    with torch.no_grad():
        m.input_conv.weight_g.mul_(gain)
        for name, p in m.named_parameters():
            if name.endswith(".bias") and not name.startswith("output_conv"):
                p.mul_(gain)
        m.output_conv[1].weight_g.mul_(1.0 / gain)
    m.invalidate()
    return m


def test_large_gain_below_fp16_max_keeps_the_tolerance():
    m = v1()
    c = log_mel_like(1, 32, seed=4)
    base = m.activation_range(c.to(DEV))
    gain = 2.0e4 / base["layers"][0]                                 # input conv output reaches 2e4
    m = scaled(m, gain)
    r = m.activation_range(c.to(DEV))
    assert 1.0e4 < r["layers"][0] < 6.0e4 and r["max"] < 65504.0 and r["fp16_headroom"] > 1.0
    ref = oracle_wave(m, c)
    with torch.no_grad():
        y = m(c.to(DEV))
    assert rel_l2(y, ref) <= 1e-3 and max_abs(y, ref) <= 1e-2, (rel_l2(y, ref), max_abs(y, ref))


def test_probe_flags_a_checkpoint_that_would_saturate_fp16():
    m = v1()
    c = log_mel_like(1, 32, seed=5)
    base = m.activation_range(c.to(DEV))
    m = scaled(m, 4.0e5 / base["layers"][0])
    r = m.activation_range(c.to(DEV))
    assert r["max"] > 65504.0 and r["fp16_headroom"] < 1.0
    ref = oracle_wave(m, c)
    with torch.no_grad():
        m.precision = "fp32"
        y32 = m(c.to(DEV))
    assert rel_l2(y32, ref) <= 1e-4                                  # the fp32 kernels are unaffected

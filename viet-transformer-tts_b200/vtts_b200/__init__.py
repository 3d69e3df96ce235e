"""vtts_b200 -- B200-native (sm_100a) synthesis hot path for Viet-Transformer-TTS.

Host-side mirror of the reference's module interface for ONE path: LengthRegulator ->
HiFi-GAN generator.  Everything numeric happens in libvtts_b200.so (hand-written CUDA behind
the C ABI in include/vtts_b200.h); this package is plumbing: parameter containers with the
reference's ``state_dict`` keys, ctypes calls, batch sharding across ranks.
"""
from . import _lib
from .acoustic import AcousticTail, ConvNorm, Decoder, FFTBlock, MultiHeadAttention, PositionwiseFeedForward, Postnet
from .conformer import ConformerBlock, ConformerConvModule, ConformerDecoder
from .dropin import install, uninstall
from .gaussian_upsampling import GaussianUpsampling
from .hifigan import GraphedForward, HiFiGAN, ResidualBlock
from .length_regulator import LengthRegulator
from .sharding import gather_waveforms, plan_shards, shard_batch
from .synthesis import PendingSynthesis, Synthesizer
from .training import conv1d_tc, conv_transpose1d_tc, hifigan_forward_tc
from .tts import OneStageTTS, TwoStageTTS, save_wav
from .vits2 import Generator, ResBlock1, ResBlock2
from .vits2_path import expand_by_path, generate_path

__all__ = [
    "HiFiGAN", "ResidualBlock", "GraphedForward", "LengthRegulator", "GaussianUpsampling", "Generator", "ResBlock1", "ResBlock2",
    "Synthesizer", "PendingSynthesis", "plan_shards", "shard_batch", "gather_waveforms", "install", "uninstall", "generate_path",
    "expand_by_path", "Decoder", "FFTBlock", "MultiHeadAttention", "PositionwiseFeedForward", "Postnet", "ConvNorm", "AcousticTail", "ConformerDecoder", "ConformerBlock", "ConformerConvModule", "OneStageTTS", "TwoStageTTS", "save_wav", "conv1d_tc", "conv_transpose1d_tc", "hifigan_forward_tc",
]
__version__ = "0.1.0"

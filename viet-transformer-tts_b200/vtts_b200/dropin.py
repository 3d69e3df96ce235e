"""Swap the reference's classes for the B200 ones without editing the reference.

``install()`` makes ``test.py`` / ``train.py``-style code pick up the kernels unchanged:

    import sys; sys.path.insert(0, "<repo>/viet-transformer-tts_b200")
    import vtts_b200; vtts_b200.install()
    from models.gan_tts.text2wav.model import Text2Wav   # now builds vtts_b200.HiFiGAN inside

Replaced symbols (reference file:line):
  models/gan_tts/hifigan/generator.py:16   HiFiGAN
  models/gan_tts/hifigan/layers.py:16      ResidualBlock
  models/tts/fastspeech2/layers.py:410     LengthRegulator      :465 GaussianUpsampling      :571 Postnet
  models/tts/fastspeech2/blocks/transformer.py:265   PositionwiseFeedForward (FFT-block convs of encoder and decoder)
  models/tts/fastspeech2/blocks/conformer.py:444     ConformerConvModule (convolution module of encoder and decoder blocks)
  espnet2.gan_tts.hifigan.HiFiGANGenerator / espnet ...length_regulator.LengthRegulator  (JETS: jets/model.py:12,18,433,481-496)
  models/gan_tts/vits2/layers.py:107       Generator
  models/gan_tts/vits2/sublayers.py:215    ResBlock1      :312 ResBlock2
  models/gan_tts/vits2/utils.py:111        generate_path (also the name imported into vits2/generator.py:8)
Every module already imported that holds a reference to one of the original classes (e.g.
``from models.gan_tts.hifigan import HiFiGAN`` in text2wav/model.py:5) is patched too.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, Tuple

_TARGETS = {
    "models.gan_tts.hifigan.generator": ("HiFiGAN",),
    "models.gan_tts.hifigan.layers": ("ResidualBlock",),
    "models.tts.fastspeech2.layers": ("LengthRegulator", "GaussianUpsampling", "Postnet"),
    "models.tts.fastspeech2.blocks.transformer": ("PositionwiseFeedForward",),
    "models.tts.fastspeech2.blocks.conformer": ("ConformerConvModule",),
    "models.gan_tts.jets.alignments": ("GaussianUpsampling",),
    # JETS takes both classes from espnet (jets/model.py:12,18); the local ones are declared copies of them
    # (hifigan/generator.py:3, fastspeech2/layers.py:410-462), so the same drop-ins apply when espnet is installed
    "espnet2.gan_tts.hifigan": ("HiFiGANGenerator",),
    "espnet2.gan_tts.hifigan.hifigan": ("HiFiGANGenerator",),
    "espnet.nets.pytorch_backend.fastspeech.length_regulator": ("LengthRegulator",),
    "models.gan_tts.vits2.layers": ("Generator",),
    "models.gan_tts.vits2.sublayers": ("ResBlock1", "ResBlock2"),
    "models.gan_tts.vits2.utils": ("generate_path",),
}
_saved: Dict[Tuple[str, str], object] = {}


def _replacements():
    from . import acoustic, conformer, gaussian_upsampling, hifigan, length_regulator, vits2, vits2_path

    return {
        "HiFiGAN": hifigan.HiFiGAN, "HiFiGANGenerator": hifigan.HiFiGAN, "ResidualBlock": hifigan.ResidualBlock,
        "LengthRegulator": length_regulator.LengthRegulator, "Generator": vits2.Generator,
        "GaussianUpsampling": gaussian_upsampling.GaussianUpsampling,
        "Postnet": acoustic.Postnet, "PositionwiseFeedForward": acoustic.PositionwiseFeedForward,
        "ConformerConvModule": conformer.ConformerConvModule,
        "ResBlock1": vits2.ResBlock1, "ResBlock2": vits2.ResBlock2, "generate_path": vits2_path.generate_path,
    }


def install(import_missing: bool = True) -> int:
    """Patch the reference modules; returns the number of attributes replaced."""
    repl = _replacements()
    originals = {}
    for modname, names in _TARGETS.items():
        mod = sys.modules.get(modname)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                mod = None  # reference (or its espnet deps) not importable: patch what exists
        if mod is None:
            continue
        for n in names:
            cur = getattr(mod, n, None)
            if cur is not None and cur is not repl[n]:
                originals[cur] = repl[n]
    count = 0
    for modname, mod in list(sys.modules.items()):
        if mod is None or modname.startswith("vtts_b200"):
            continue
        for attr, val in list(getattr(mod, "__dict__", {}).items()):
            try:
                new = originals.get(val)
            except TypeError:
                continue
            if new is not None:
                _saved.setdefault((modname, attr), val)
                setattr(mod, attr, new)
                count += 1
    return count


def uninstall() -> int:
    n = 0
    for (modname, attr), val in list(_saved.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, attr, val)
            n += 1
    _saved.clear()
    return n

"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules from /root/reference.

Only ``tests/golden/make_golden.py`` and the ``-m "not gpu"`` pinning tests import this file,
and only inside the build container: ``/root/reference`` does not exist on the GPU box, so
every caller must check :func:`reference_available` first.  Nothing under
``viet-transformer-tts_b200/`` may import it.

Recipe: the reference package ``__init__`` files drag in espnet2
(models/gan_tts/hifigan/__init__.py:7-12 -> loss.py:15-17), so the package inits are bypassed
with empty placeholder packages and the leaf files are loaded by path.  For the
LengthRegulator, fastspeech2/layers.py:10-11 imports exactly two espnet symbols, which are
stubbed (``pad_list`` is loaded from the reference's own byte-identical local copy,
fastspeech2/function.py:97-124, by executing that file's function body from its source span).
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("VTTS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models/gan_tts/hifigan/generator.py"))


def _placeholder(name: str, path: str | None = None) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = [path] if path else []  # mark as package
        sys.modules[name] = mod
    return mod


def _load_file(modname: str, relpath: str) -> types.ModuleType:
    if modname in sys.modules and getattr(sys.modules[modname], "__file__", None):
        return sys.modules[modname]
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def _function_from_source(relpath: str, fname: str, globs: dict):
    """Compile ONE top-level function of a reference file without importing the file.

    Used for ``pad_list`` (fastspeech2/function.py:97-124): importing function.py pulls numba
    jit-compilation of unrelated alignment helpers; the function itself only needs torch.
    """
    path = os.path.join(REF_ROOT, relpath)
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == fname:
            node.decorator_list = []
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            ns = dict(globs)
            exec(code, ns)
            return ns[fname]
    raise ImportError(f"{fname} not found in {path}")


def load_hifigan():
    """Return (HiFiGAN, ResidualBlock) -- models/gan_tts/hifigan/generator.py:16, layers.py:16."""
    _placeholder("models", os.path.join(REF_ROOT, "models"))
    _placeholder("models.gan_tts", os.path.join(REF_ROOT, "models/gan_tts"))
    _placeholder("models.gan_tts.hifigan", os.path.join(REF_ROOT, "models/gan_tts/hifigan"))
    layers = _load_file("models.gan_tts.hifigan.layers", "models/gan_tts/hifigan/layers.py")
    gen = _load_file("models.gan_tts.hifigan.generator", "models/gan_tts/hifigan/generator.py")
    return gen.HiFiGAN, layers.ResidualBlock


def load_vits2_generator():
    """Return (Generator, ResBlock1, ResBlock2) -- vits2/layers.py:107, sublayers.py:215,312."""
    _placeholder("models", os.path.join(REF_ROOT, "models"))
    _placeholder("models.gan_tts", os.path.join(REF_ROOT, "models/gan_tts"))
    _placeholder("models.gan_tts.vits2", os.path.join(REF_ROOT, "models/gan_tts/vits2"))
    for leaf in ("utils", "transforms", "sublayers", "attentions", "layers"):
        _load_file(f"models.gan_tts.vits2.{leaf}", f"models/gan_tts/vits2/{leaf}.py")
    layers = sys.modules["models.gan_tts.vits2.layers"]
    sub = sys.modules["models.gan_tts.vits2.sublayers"]
    return layers.Generator, sub.ResBlock1, sub.ResBlock2


def load_length_regulator():
    """Return the reference ``LengthRegulator`` class (fastspeech2/layers.py:410-462)."""
    import torch

    pad_list = _function_from_source("models/tts/fastspeech2/function.py", "pad_list", {"torch": torch})
    for name in ("espnet", "espnet.nets", "espnet.nets.pytorch_backend"):
        _placeholder(name)
    nets_utils = _placeholder("espnet.nets.pytorch_backend.nets_utils")
    nets_utils.pad_list = pad_list
    for name in ("espnet2", "espnet2.tts", "espnet2.tts.gst"):
        _placeholder(name)
    se = _placeholder("espnet2.tts.gst.style_encoder")
    se.ReferenceEncoder = type("ReferenceEncoder", (), {})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # the real (namespace-free) package import now works: models/tts/fastspeech2/__init__.py
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        if not getattr(sys.modules[k], "__file__", None) and k in ("models", "models.tts"):
            del sys.modules[k]
    try:
        from models.tts.fastspeech2.layers import LengthRegulator  # type: ignore
    except Exception:
        # fall back to loading only layers.py by path with placeholder parents
        _placeholder("models", os.path.join(REF_ROOT, "models"))
        _placeholder("models.tts", os.path.join(REF_ROOT, "models/tts"))
        _placeholder("models.tts.fastspeech2", os.path.join(REF_ROOT, "models/tts/fastspeech2"))
        layers = _load_file("models.tts.fastspeech2.layers", "models/tts/fastspeech2/layers.py")
        LengthRegulator = layers.LengthRegulator
    return LengthRegulator


def load_gaussian_upsampling():
    """Return the reference ``GaussianUpsampling`` class (fastspeech2/layers.py:465-520)."""
    load_length_regulator()
    import sys as _sys

    return _sys.modules["models.tts.fastspeech2.layers"].GaussianUpsampling


def load_vits2_utils():
    """Return the reference module ``models/gan_tts/vits2/utils.py`` (generate_path :111-126, sequence_mask :104-108)."""
    _placeholder("models", os.path.join(REF_ROOT, "models"))
    _placeholder("models.gan_tts", os.path.join(REF_ROOT, "models/gan_tts"))
    _placeholder("models.gan_tts.vits2", os.path.join(REF_ROOT, "models/gan_tts/vits2"))
    return _load_file("models.gan_tts.vits2.utils", "models/gan_tts/vits2/utils.py")

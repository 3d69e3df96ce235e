"""Host-side module shells: reference-compatible trees, keys, policies (no GPU needed)."""
import copy

import numpy as np
import pytest
import torch

import restate
import vtts_b200
from conftest import load_golden, max_abs, state_dict_from


def test_hifigan_state_dict_keys_match_reference_v1():
    z = load_golden("hifigan_v1.npz")
    torch.manual_seed(int(z["seed"]))
    m = vtts_b200.HiFiGAN()
    sd = m.state_dict()
    assert sorted(sd) == [str(k) for k in z["keys"]]
    assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == [str(s) for s in z["shapes"]]
    # same seed -> same parameter draw as the reference constructor (construction order parity)
    sums = np.array([float(sd[k].double().sum()) for k in sorted(sd)])
    assert np.allclose(sums, z["sums"], rtol=0, atol=1e-9)
    assert m.upsample_factor == 256 and m.num_upsamples == 4 and m.num_blocks == 3


def test_hifigan_loads_reference_checkpoint_and_eager_path_matches_golden():
    z = load_golden("hifigan_small.npz")
    m = vtts_b200.HiFiGAN(in_channels=8, out_channels=1, channels=32, global_channels=4, kernel_size=7,
                          upsample_scales=[4, 2], upsample_kernel_sizes=[8, 4], resblock_kernel_sizes=[3, 5],
                          resblock_dilations=[[1, 3], [1, 2]])
    missing = m.load_state_dict(state_dict_from(z), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    # the autograd (training) path is the module tree itself and must equal the reference
    y = m(torch.from_numpy(z["c"]), torch.from_numpy(z["g"]))
    assert y.requires_grad
    assert max_abs(y, torch.from_numpy(z["y"])) < 1e-6
    y.sum().backward()
    assert m.input_conv.weight_v.grad is not None


def test_remove_weight_norm_changes_keys_like_reference():
    m = vtts_b200.HiFiGAN(in_channels=8, channels=16, upsample_scales=[2], upsample_kernel_sizes=[4],
                          resblock_kernel_sizes=[3], resblock_dilations=[[1]])
    assert "input_conv.weight_g" in m.state_dict()
    m.remove_weight_norm()
    assert "input_conv.weight" in m.state_dict() and "input_conv.weight_g" not in m.state_dict()
    m.apply_weight_norm()
    assert "blocks.0.convs1.0.1.weight_v" in m.state_dict()


def test_vits2_generator_keys_and_eager_path():
    z = load_golden("vits2_small.npz")
    for tag, rb, dil in (("rb1", "1", [[1, 3, 5], [1, 2, 4]]), ("rb2", "2", [[1, 3], [1, 2]])):
        m = vtts_b200.Generator(12, resblock=rb, resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=dil,
                                upsample_rates=[4, 2], upsample_initial_channel=32, upsample_kernel_sizes=[8, 4],
                                gin_channels=5)
        sd = state_dict_from(z, prefix=f"{tag}.sd.")
        assert sorted(m.state_dict()) == sorted(sd)
        m.load_state_dict(sd)
        y = m(torch.from_numpy(z[f"{tag}.x"]), torch.from_numpy(z[f"{tag}.g"]))
        assert max_abs(y, torch.from_numpy(z[f"{tag}.y"])) < 1e-6


def test_synthesis_path_refuses_cpu_loudly():
    m = vtts_b200.HiFiGAN(in_channels=8, channels=16, upsample_scales=[2], upsample_kernel_sizes=[4],
                          resblock_kernel_sizes=[3], resblock_dilations=[[1]])
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 8, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.inference(torch.randn(5, 8))
    lr = vtts_b200.LengthRegulator()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lr(torch.randn(1, 3, 4), torch.ones(1, 3, dtype=torch.long))


def test_missing_library_fails_loudly(monkeypatch):
    from vtts_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvtts_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_deepcopy_does_not_share_device_handles():
    # (old-style weight_norm modules are not deep-copyable in torch itself, same as the reference)
    m = vtts_b200.HiFiGAN(in_channels=8, channels=16, upsample_scales=[2], upsample_kernel_sizes=[4],
                          resblock_kernel_sizes=[3], resblock_dilations=[[1]], use_weight_norm=False)
    m._handles[0] = 12345  # pretend
    c = copy.deepcopy(m)
    assert c._handles == {}
    m._handles.clear()


def test_layer_order_matches_handle_enumeration():
    import ctypes
    from vtts_b200 import _lib

    lib = _lib.load()
    m = vtts_b200.HiFiGAN(global_channels=16)
    cfg = m._gen_config()
    h = ctypes.c_void_p()
    _lib.check(lib.vtts_gen_create(ctypes.byref(cfg), ctypes.byref(h)))
    try:
        mods = m._layer_modules()
        assert lib.vtts_gen_num_layers(h) == len(mods) == 1 + 4 * 19 + 1 + 1
        info = _lib.VttsLayerInfo()
        for i, mod in enumerate(mods):
            _lib.check(lib.vtts_gen_layer_info(h, i, ctypes.byref(info)))
            w = mod.weight_v
            want = (info.cout, info.cin, info.ksize) if info.kind == 0 else (info.cin, info.cout, info.ksize)
            assert tuple(w.shape) == want, i
            if info.kind == 0:
                assert mod.dilation[0] == info.dilation
    finally:
        lib.vtts_gen_destroy(h)


def test_dropin_install_patches_reference_modules():
    ref_loader = pytest.importorskip("ref_loader")
    if not ref_loader.reference_available():
        pytest.skip("/root/reference not present")
    import sys
    RefHiFiGAN, _ = ref_loader.load_hifigan()
    holder = type(sys)("fake_text2wav")
    holder.HiFiGAN = RefHiFiGAN  # `from models.gan_tts.hifigan import HiFiGAN` in a caller module
    sys.modules["fake_text2wav"] = holder
    try:
        n = vtts_b200.install(import_missing=False)
        assert n >= 2
        assert holder.HiFiGAN is vtts_b200.HiFiGAN
        assert sys.modules["models.gan_tts.hifigan.generator"].HiFiGAN is vtts_b200.HiFiGAN
    finally:
        vtts_b200.uninstall()
        del sys.modules["fake_text2wav"]
    assert sys.modules["models.gan_tts.hifigan.generator"].HiFiGAN is RefHiFiGAN


def test_dropin_install_covers_the_espnet_names_jets_imports():
    """jets/model.py:12,18 takes HiFiGANGenerator and LengthRegulator from espnet (absent here): with fake espnet modules in
    place, install() rebinds those names - in the defining modules and in a module that imported them - and uninstall() restores."""
    import sys
    import types

    class FakeGen:  # stands for espnet2.gan_tts.hifigan.HiFiGANGenerator
        pass

    class FakeLR:
        pass

    names = ["espnet2", "espnet2.gan_tts", "espnet2.gan_tts.hifigan", "espnet", "espnet.nets", "espnet.nets.pytorch_backend",
             "espnet.nets.pytorch_backend.fastspeech", "espnet.nets.pytorch_backend.fastspeech.length_regulator", "fake_jets_model"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        for n in names:
            sys.modules[n] = types.ModuleType(n)
        sys.modules["espnet2.gan_tts.hifigan"].HiFiGANGenerator = FakeGen
        sys.modules["espnet.nets.pytorch_backend.fastspeech.length_regulator"].LengthRegulator = FakeLR
        sys.modules["fake_jets_model"].HiFiGANGenerator = FakeGen          # `from espnet2.gan_tts.hifigan import HiFiGANGenerator`
        sys.modules["fake_jets_model"].LengthRegulator = FakeLR
        vtts_b200.install(import_missing=False)
        assert sys.modules["fake_jets_model"].HiFiGANGenerator is vtts_b200.HiFiGAN
        assert sys.modules["fake_jets_model"].LengthRegulator is vtts_b200.LengthRegulator
        assert sys.modules["espnet2.gan_tts.hifigan"].HiFiGANGenerator is vtts_b200.HiFiGAN
        vtts_b200.uninstall()
        assert sys.modules["fake_jets_model"].HiFiGANGenerator is FakeGen and sys.modules["fake_jets_model"].LengthRegulator is FakeLR
    finally:
        vtts_b200.uninstall()
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m

"""GPU parity: HiFi-GAN generator kernels vs the CPU oracle and the reference's golden vectors.

Tolerances: fp32 path <= 2e-5 max-abs / 1e-5 rel-L2 against the fp32 CPU oracle; bf16 tensor-core
path rel-L2 <= 1e-3 and max-abs <= 1e-2 on the waveform (BASELINE.json north_star).
"""
import numpy as np
import pytest
import torch

import restate
import vtts_b200
from conftest import load_golden, max_abs, rel_l2, state_dict_from

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_MAXABS, FP32_REL = 2e-5, 1e-5
BF16_MAXABS, BF16_REL = 1e-2, 1e-3


def small_model(precision="fp32"):
    z = load_golden("hifigan_small.npz")
    m = vtts_b200.HiFiGAN(in_channels=8, out_channels=1, channels=32, global_channels=4, kernel_size=7,
                          upsample_scales=[4, 2], upsample_kernel_sizes=[8, 4], resblock_kernel_sizes=[3, 5],
                          resblock_dilations=[[1, 3], [1, 2]])
    m.load_state_dict(state_dict_from(z))
    m.precision = precision
    return m.to(DEV).eval(), z


def v1_model(precision):
    z = load_golden("hifigan_v1.npz")
    torch.manual_seed(int(z["seed"]))
    m = vtts_b200.HiFiGAN()
    sums = np.array([float(v.double().sum()) for _, v in sorted(m.state_dict().items())])
    assert np.allclose(sums, z["sums"], rtol=0, atol=1e-9), "seeded weights differ from the reference draw"
    m.precision = precision
    return m.to(DEV).eval(), z


def test_dbg_conv1d_fp32_matches_torch_cpu():
    from vtts_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (B, cin, cout, L, k, d) in [(2, 8, 16, 50, 3, 1), (1, 33, 70, 301, 7, 3), (3, 64, 64, 129, 11, 5), (1, 32, 1, 257, 7, 1)]:
        x = torch.randn(B, cin, L, generator=g)
        w = torch.randn(cout, cin, k, generator=g) * 0.1
        b = torch.randn(cout, generator=g)
        r = torch.randn(B, cout, L, generator=g)
        ref = torch.nn.functional.conv1d(torch.nn.functional.leaky_relu(x, 0.1), w, b, padding=(k - 1) // 2 * d, dilation=d) + r
        xd, wd, bd, rd = (t.to(DEV).contiguous() for t in (x, w, b, r))
        y = torch.empty(B, cout, L, device=DEV)
        _lib.check(lib.vtts_dbg_conv1d_fp32(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), y.data_ptr(),
                                            B, cin, cout, L, k, d, 0.1, 0, torch.cuda.current_stream().cuda_stream))
        assert max_abs(y, ref) < 2e-5


def test_fp32_small_golden_with_and_without_global_conditioning():
    m, z = small_model("fp32")
    with torch.no_grad():
        y = m(torch.from_numpy(z["c"]).to(DEV), torch.from_numpy(z["g"]).to(DEV))
        y0 = m(torch.from_numpy(z["c"]).to(DEV))
    assert max_abs(y, torch.from_numpy(z["y"])) < FP32_MAXABS
    assert max_abs(y0, torch.from_numpy(z["y_nog"])) < FP32_MAXABS
    assert m.last_launch_count > 0


def test_fp32_small2_resblock_without_additional_convs_odd_scale():
    z = load_golden("hifigan_small2.npz")
    m = vtts_b200.HiFiGAN(in_channels=6, out_channels=1, channels=16, kernel_size=5, upsample_scales=[3, 2],
                          upsample_kernel_sizes=[6, 4], resblock_kernel_sizes=[3], resblock_dilations=[[1, 2, 3]],
                          use_additional_convs=False, bias=False, use_weight_norm=False)
    m.load_state_dict(state_dict_from(z))
    m.precision = "fp32"
    m = m.to(DEV).eval()
    with torch.no_grad():
        y = m(torch.from_numpy(z["c"]).to(DEV))
    assert tuple(y.shape) == z["y"].shape
    assert max_abs(y, torch.from_numpy(z["y"])) < FP32_MAXABS


def test_fp32_stage_dumps_match_oracle_stages():
    m, z = small_model("fp32")
    sd = state_dict_from(z)
    c, g = torch.from_numpy(z["c"]), torch.from_numpy(z["g"])
    _, stages = restate.hifigan_forward(sd, c, g, upsample_scales=(4, 2), resblock_kernel_sizes=(3, 5),
                                        resblock_dilations=((1, 3), (1, 2)), return_stages=True)
    with torch.no_grad():
        for s, ref in enumerate(stages):
            got = m.debug_stage(c.to(DEV), s, g.to(DEV))
            assert got.shape == ref.shape
            assert max_abs(got, ref) < FP32_MAXABS, f"stage {s}"


def test_fp32_v1_golden_and_inference_layout():
    m, z = v1_model("fp32")
    c = torch.from_numpy(z["c"]).to(DEV)
    with torch.no_grad():
        y = m(c)
    ref = torch.from_numpy(z["y"])
    assert max_abs(y, ref) < FP32_MAXABS and rel_l2(y, ref) < FP32_REL
    inf = m.inference(c[0].transpose(0, 1))
    assert tuple(inf.shape) == z["inference"].shape
    assert max_abs(inf, torch.from_numpy(z["inference"])) < FP32_MAXABS
    # remove_weight_norm leaves the function unchanged (generator.py:173-183) and invalidates the pack
    m.remove_weight_norm()
    with torch.no_grad():
        y2 = m(c)
    assert max_abs(y2, ref) < FP32_MAXABS


def test_fp32_weight_update_invalidates_packed_weights():
    m, z = small_model("fp32")
    c = torch.from_numpy(z["c"]).to(DEV)
    with torch.no_grad():
        y0 = m(c)
        m.output_conv[1].bias.add_(0.25)
        y1 = m(c)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = restate.hifigan_forward(sd, c.cpu(), None, upsample_scales=(4, 2), resblock_kernel_sizes=(3, 5),
                                  resblock_dilations=((1, 3), (1, 2)))
    assert max_abs(y1, ref) < FP32_MAXABS
    assert max_abs(y1, y0) > 1e-3


def test_fp32_vits2_skins_match_golden():
    z = load_golden("vits2_small.npz")
    for tag, rb, dil in (("rb1", "1", [[1, 3, 5], [1, 2, 4]]), ("rb2", "2", [[1, 3], [1, 2]])):
        m = vtts_b200.Generator(12, resblock=rb, resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=dil,
                                upsample_rates=[4, 2], upsample_initial_channel=32, upsample_kernel_sizes=[8, 4],
                                gin_channels=5)
        m.load_state_dict(state_dict_from(z, prefix=f"{tag}.sd."))
        m.precision = "fp32"
        m = m.to(DEV).eval()
        with torch.no_grad():
            y = m(torch.from_numpy(z[f"{tag}.x"]).to(DEV), torch.from_numpy(z[f"{tag}.g"]).to(DEV))
        assert max_abs(y, torch.from_numpy(z[f"{tag}.y"])) < FP32_MAXABS


def test_fp32_ragged_lengths_and_tile_edges_vs_oracle():
    """T values around the 128-wide time tile; batch rows are independent."""
    sd = restate.make_hifigan_state_dict(in_channels=16, channels=64, upsample_scales=(4, 2),
                                         resblock_kernel_sizes=(3, 11), resblock_dilations=((1, 3, 5), (1, 3, 5)), seed=3)
    m = vtts_b200.HiFiGAN(in_channels=16, channels=64, upsample_scales=[4, 2], upsample_kernel_sizes=[8, 4],
                          resblock_kernel_sizes=[3, 11], resblock_dilations=[[1, 3, 5], [1, 3, 5]])
    m.load_state_dict(sd)
    m.precision = "fp32"
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    for T in (1, 2, 15, 16, 17, 33):
        c = torch.randn(2, 16, T, generator=g)
        ref = restate.hifigan_forward(sd, c, None, upsample_scales=(4, 2), resblock_kernel_sizes=(3, 11),
                                      resblock_dilations=((1, 3, 5), (1, 3, 5)))
        with torch.no_grad():
            y = m(c.to(DEV))
            y_row0 = m(c[:1].to(DEV))
        assert max_abs(y, ref) < FP32_MAXABS, T
        assert torch.equal(y[:1], y_row0)


def test_autograd_path_on_gpu_matches_kernels():
    m, z = small_model("fp32")
    c = torch.from_numpy(z["c"]).to(DEV)
    y_train = m(c)  # grad enabled + trainable parameters -> torch autograd path (training policy)
    assert y_train.requires_grad
    with torch.no_grad():
        y_kernel = m(c)
    assert max_abs(y_train, y_kernel) < 1e-4  # cuDNN may use TF32


def test_synthesizer_end_to_end_host_buffers():
    m, z = small_model("fp32")
    synth = vtts_b200.Synthesizer(m)
    g = torch.Generator().manual_seed(9)
    hs = torch.randn(3, 6, 16, generator=g).pin_memory()
    ds = torch.randint(1, 5, (3, 6), generator=g).pin_memory()
    wav, wav_len = synth(hs, ds)
    frames, _ = restate.lr_expand(hs, ds.clone())
    sd = state_dict_from(z)
    ref = restate.hifigan_forward(sd, frames[..., :8].transpose(1, 2), None, upsample_scales=(4, 2),
                                  resblock_kernel_sizes=(3, 5), resblock_dilations=((1, 3), (1, 2)))
    assert not wav.is_cuda and max_abs(wav, ref) < FP32_MAXABS
    assert torch.equal(wav_len, ds.sum(1) * 8)

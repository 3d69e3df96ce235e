"""f3, conformer variant of the acoustic decoder (`block_type: "conformer"`, the shipped default, model_config.yaml:17).

Golden: ``tests/golden/conformer_decoder.npz`` from the reference's own ``Decoder`` (blocks/conformer.py:93-169) in eval mode
(``make_golden.py::conformer_decoder``).  CPU: keys / seeded parameter draw of the drop-in classes, the oracle restatement,
the autograd (PyTorch) path.  GPU: the convolution module on the kernels (two pointwise convs on the tcgen05 conv kernel,
GLU + depthwise conv + BatchNorm + Swish in ``vtts_dwconv_glu_swish``) against the golden at rel-L2 <= 1e-3, and the
depthwise kernel alone against torch in fp32.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import restate
import vtts_b200
from conftest import load_golden, max_abs, rel_l2, state_dict_from
from vtts_b200.conformer import ConformerDecoder

DEV = "cuda:0"
SHAPES = {"small": (2, 64, 2), "c4": (6, 384, 8)}


def build(tag):
    layers, hidden, heads = SHAPES[tag]
    cfg = {"decoder_head": heads, "ffn_expansion_factor": 4, "conv_expansion_factor": 2, "conv_kernel_size": 31,
           "half_step_residual": True, "decoder_dropout": 0.1}
    torch.manual_seed(1234)
    dec = ConformerDecoder(layers, hidden, 1000, cfg).eval()
    with torch.no_grad():
        for i, blk in enumerate(dec.layer_stack):               # same formula as make_golden.py::conformer_decoder
            bn = blk.sequential[2].module.sequential[5]
            t = torch.arange(bn.num_features, dtype=torch.float32)
            bn.running_mean.copy_(0.2 * torch.sin(0.37 * t + i)); bn.running_var.copy_(1.0 + 0.5 * torch.cos(0.11 * t + 2 * i))
            bn.weight.data.copy_(1.0 + 0.3 * torch.sin(0.05 * t + 3 * i)); bn.bias.data.copy_(0.1 * torch.cos(0.23 * t + i))
    return dec


@pytest.fixture(scope="module")
def gold():
    return load_golden("conformer_decoder.npz")


@pytest.mark.parametrize("tag", ["small", "c4"])
def test_keys_and_seeded_draw_match_the_reference_class(gold, tag):
    sd = build(tag).state_dict()
    assert sorted(sd) == [str(k) for k in gold[f"{tag}.keys"]]
    sums = np.array([float(sd[k].double().sum()) for k in sorted(sd)])
    assert np.allclose(sums, gold[f"{tag}.sums"], rtol=0, atol=1e-6 * np.maximum(1.0, np.abs(gold[f"{tag}.sums"])))


def _small_sd(gold, m):
    sd = state_dict_from(gold, prefix="small.sd.")
    for k, v in m.state_dict().items():
        if "position" in k:
            sd[k] = v.detach().clone()
    return sd


def test_oracle_and_autograd_path_match_the_reference(gold):
    m = build("small")
    sd = _small_sd(gold, m)
    assert not m.load_state_dict(sd, strict=True).missing_keys
    frames, mel_len = torch.from_numpy(gold["small.frames"]), torch.from_numpy(gold["small.mel_len"])
    mask = torch.arange(frames.shape[1])[None] >= mel_len[:, None]
    ref = torch.from_numpy(gold["small.dec"])
    assert max_abs(restate.conformer_decoder_forward(sd, frames, mask, 2), ref) < 2e-5
    y, _ = m(frames, mask)                                      # grad mode: PyTorch formula
    assert y.requires_grad and max_abs(y, ref) < 2e-5
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(frames, mask)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "c4"])
def test_kernel_path_against_the_reference_golden(gold, tag):
    m = build(tag).to(DEV)
    frames, mel_len = torch.from_numpy(gold[f"{tag}.frames"]).to(DEV), torch.from_numpy(gold[f"{tag}.mel_len"]).to(DEV)
    mask = torch.arange(frames.shape[1], device=DEV)[None] >= mel_len[:, None]
    with torch.no_grad():
        y, _ = m(frames, mask)
    ref = torch.from_numpy(gold[f"{tag}.dec"])
    r, a = rel_l2(y, ref), max_abs(y, ref)
    print(f"conformer {tag}: rel-L2 {r:.3e} max-abs {a:.3e}")
    assert r <= 1e-3 and a <= 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("C,k,L,B", [(64, 31, 100, 2), (384, 31, 333, 1), (256, 7, 65, 3), (32, 1, 10, 1)])
def test_depthwise_glu_swish_kernel_vs_torch(C, k, L, B):
    from vtts_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(C + k)
    pw = torch.randn(B, L, 2 * C, generator=g).to(DEV)
    w = (torch.randn(C, k, generator=g) / k ** 0.5).to(DEV)
    bias = torch.randn(C, generator=g).to(DEV)
    out = torch.empty(B, L, C, dtype=torch.float16, device=DEV)
    _lib.check(lib.vtts_dwconv_glu_swish(pw.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), _lib.PRECISION["fp16"],
                                         B, L, C, k, torch.cuda.current_stream().cuda_stream))
    u = (pw[..., :C] * torch.sigmoid(pw[..., C:])).transpose(1, 2)
    v = F.conv1d(u, w.unsqueeze(1), bias, padding=(k - 1) // 2, groups=C)
    ref = (v * torch.sigmoid(v)).transpose(1, 2)
    assert max_abs(out.float(), ref) <= 2e-3 * max(1.0, float(ref.abs().max()))

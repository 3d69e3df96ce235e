"""f4: generator backward on the kernels (``HiFiGAN.train_backend = "tc"``, vtts_b200/training.py) against PyTorch
autograd of the same module tree in fp32 (TF32 off) -- what ``hifigan_trainer.py:143-167`` differentiates.

Forward and dgrad run on the tcgen05 conv kernel, wgrad on cuBLAS (16-bit operands, fp32 accumulation).  Tolerances:
single layers against F.conv1d on the SAME rounded operands are tight; the whole generator against the fp32 module tree:
waveform rel-L2 <= 2e-3, every parameter gradient rel-L2 <= 3e-2 (16-bit operand rounding of activations and gradients
through up to ~40 layers; measured values are printed).
"""
import pytest
import torch
import torch.nn.functional as F

import vtts_b200
from conftest import max_abs, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("cin,cout,k,d,L,B", [(64, 64, 3, 1, 200, 2), (128, 128, 7, 3, 333, 1), (32, 32, 11, 5, 500, 2),
                                              (80, 512, 7, 1, 40, 2), (32, 1, 7, 1, 300, 2)])
def test_conv1d_tc_forward_and_gradients(cin, cout, k, d, L, B):
    g = torch.Generator().manual_seed(cin + k)
    rnd = lambda t: t.half().float()
    x = rnd(torch.randn(B, cin, L, generator=g)).to(DEV).requires_grad_(True)
    w = rnd(torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(cout, generator=g).to(DEV).requires_grad_(True)
    dy = rnd(torch.randn(B, cout, L, generator=g)).to(DEV)
    y = vtts_b200.conv1d_tc(x, w, b, d)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    ref_y = F.conv1d(x, w, b, padding=(k - 1) // 2 * d, dilation=d)
    ref_y.backward(dy)
    for name, a, r in zip(("y", "dx", "dw", "db"), got, (ref_y.detach(), x.grad, w.grad, b.grad)):
        assert max_abs(a, r) <= 2e-3 * max(1.0, float(r.abs().max())), (name, max_abs(a, r), float(r.abs().max()))


@pytest.mark.parametrize("s,pad,op", [(8, 4, 0), (2, 1, 0), (3, 2, 1)])
def test_conv_transpose1d_tc_forward_and_gradients(s, pad, op):
    g = torch.Generator().manual_seed(s)
    rnd = lambda t: t.half().float()
    x = rnd(torch.randn(2, 64, 37, generator=g)).to(DEV).requires_grad_(True)
    w = rnd(torch.randn(64, 32, 2 * s, generator=g) / (64 * 2) ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(32, generator=g).to(DEV).requires_grad_(True)
    y = vtts_b200.conv_transpose1d_tc(x, w, b, s, pad, op)
    ref = F.conv_transpose1d(x, w, b, stride=s, padding=pad, output_padding=op)
    assert y.shape == ref.shape
    dy = rnd(torch.randn(ref.shape, generator=g)).to(DEV)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    ref.backward(dy)
    for name, a, r in zip(("y", "dx", "dw", "db"), got, (ref.detach(), x.grad, w.grad, b.grad)):
        assert max_abs(a, r) <= 2e-3 * max(1.0, float(r.abs().max())), (name, max_abs(a, r))


def _grads(m, c, gc, w_out):
    m.zero_grad(set_to_none=True)
    y = m(c, gc)
    (y * w_out).sum().backward()
    return y.detach(), {n: p.grad.detach().clone() for n, p in m.named_parameters()}, c.grad.detach().clone()


def test_generator_backward_on_the_kernels_matches_autograd_of_the_module_tree():
    torch.manual_seed(7)
    m = vtts_b200.HiFiGAN(in_channels=80, channels=128, global_channels=16, upsample_scales=[4, 2], upsample_kernel_sizes=[8, 4],
                          resblock_kernel_sizes=[3, 7], resblock_dilations=[[1, 3], [1, 3]]).to(DEV).train()
    g = torch.Generator().manual_seed(1)
    c = torch.randn(3, 80, 24, generator=g).to(DEV).requires_grad_(True)
    gc = torch.randn(3, 16, 1, generator=g).to(DEV)
    w_out = torch.randn(3, 1, 24 * 8, generator=g).to(DEV)
    m.train_backend = "eager"
    y_ref, g_ref, dc_ref = _grads(m, c, gc, w_out)
    c.grad = None
    m.train_backend = "tc"
    y, g_tc, dc = _grads(m, c, gc, w_out)
    assert rel_l2(y, y_ref) <= 2e-3, rel_l2(y, y_ref)
    errs = sorted(((rel_l2(g_tc[n], g_ref[n]), n) for n in g_ref), reverse=True)
    flat_tc = torch.cat([g_tc[n].flatten().double() for n in sorted(g_ref)])
    flat_ref = torch.cat([g_ref[n].flatten().double() for n in sorted(g_ref)])
    cos = float(torch.dot(flat_tc, flat_ref) / (flat_tc.norm() * flat_ref.norm()))
    whole = float((flat_tc - flat_ref).norm() / flat_ref.norm())
    print(f"train_backend=tc: waveform rel-L2 {rel_l2(y, y_ref):.2e}, input grad {rel_l2(dc, dc_ref):.2e}, whole gradient rel-L2 "
          f"{whole:.2e} cosine {cos:.6f}, worst tensors {errs[:2]}")
    assert set(g_tc) == set(g_ref)
    # The graph itself is exact (tools/dbg_train_backend.py: 5e-7 with fp32 convs in the same graph); rounding only the FORWARD
    # operands to fp16 already moves single tensors of this random-init GAN generator by up to 3e-2 (sums that partly cancel),
    # which is therefore the per-tensor bar; the update direction as a whole must agree much better.
    assert errs[0][0] <= 6e-2, errs[:3]
    assert whole <= 2.5e-2 and cos >= 0.9995
    assert rel_l2(dc, dc_ref) <= 3e-2


def test_a_training_step_updates_the_weights_the_synthesis_path_uses():
    torch.manual_seed(3)
    m = vtts_b200.HiFiGAN(channels=128, upsample_scales=[4, 2], upsample_kernel_sizes=[8, 4], resblock_kernel_sizes=[3],
                          resblock_dilations=[[1, 3]]).to(DEV)
    m.train_backend = "tc"
    opt = torch.optim.SGD(m.parameters(), lr=1e-2)
    c = torch.randn(2, 80, 16, generator=torch.Generator().manual_seed(0)).to(DEV)
    with torch.no_grad():
        y0 = m(c)
    loss = m(c).pow(2).mean()
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    opt.step()
    with torch.no_grad():
        y1 = m(c)                     # kernels re-upload the updated parameters (version signature)
        m.precision = "fp32"
        y1_ref = m(c)
    assert not torch.equal(y0, y1) and rel_l2(y1, y1_ref) <= 1e-3


def test_v1_backward_runs_at_the_trainer_segment_size():
    """hifigan_trainer.py feeds random segments (config segment_size: 64 frames); one backward of V1 at batch 2."""
    torch.manual_seed(1234)
    m = vtts_b200.HiFiGAN().to(DEV)
    m.train_backend = "tc"
    c = torch.randn(2, 80, 64, generator=torch.Generator().manual_seed(0)).to(DEV)
    y = m(c)
    assert y.shape == (2, 1, 64 * 256) and y.requires_grad
    y.abs().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0 for p in m.parameters())

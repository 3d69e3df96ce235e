"""GPU parity of the fused ResidualBlock chain kernel (csrc/chain_tc.cu: one launch = all 6 convs of a block,
models/gan_tts/hifigan/layers.py:83-98) and of the full-size V1 forward against the reference golden.

The single-block cases compare with a float64 torch evaluation of the same block whose conv operands are rounded to
16 bits at the points where the kernel rounds them (isolates kernel correctness from the number format; what is left
is fp32 accumulation order), and with the exact float64 block at the north_star tolerance.
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vtts_b200
from conftest import load_golden, max_abs, rel_l2
from vtts_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_block(C, k, n_convs, seed):
    g = torch.Generator().manual_seed(seed)
    ws = [torch.randn(C, C, k, generator=g) / (C * k) ** 0.5 for _ in range(n_convs)]
    bs = [0.1 * torch.randn(C, generator=g) for _ in range(n_convs)]
    return ws, bs


def block_f64(x, ws, bs, dil, has2, slope, rnd):
    """layers.py:83-98 in float64; rnd() models the kernel's 16-bit operand rounding (identity = exact)."""
    x = x.double()
    i = 0
    for d in dil:
        k = ws[i].shape[-1]
        xt = F.conv1d(rnd(F.leaky_relu(x, slope)), rnd(ws[i].double()), bs[i].double(), padding=(k - 1) // 2 * d, dilation=d)
        i += 1
        if has2:
            xt = F.conv1d(rnd(F.leaky_relu(xt, slope)), rnd(ws[i].double()), bs[i].double(), padding=(k - 1) // 2)
            i += 1
        x = xt + x
    return x


def run_chain(x, ws, bs, dil, has2, slope=0.1, fp16=1):
    lib = _lib.load()
    B, C, L = x.shape
    xd = x.to(DEV).contiguous()
    wd = [w.to(DEV).contiguous() for w in ws]
    bd = [b.to(DEV).contiguous() for b in bs]
    y = torch.empty_like(xd)
    n = len(ws)
    wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in wd])
    bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bd])
    dl = (ctypes.c_int * len(dil))(*dil)
    ms = ctypes.c_float(0.0)
    _lib.check(lib.vtts_dbg_resblock_chain(xd.data_ptr(), wp, bp, y.data_ptr(), B, C, L, ws[0].shape[-1], dl, len(dil),
                                           1 if has2 else 0, slope, fp16, 0, ctypes.byref(ms),
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return y.cpu()


CHAIN_CASES = [
    # C, k, L, B, dilations, has2     (tile = 480 positions at 32 channels, 240 at 64; valid part = tile - 2 * halo)
    (32, 3, 100, 1, (1, 3, 5), True),        # shorter than one tile
    (32, 11, 361, 3, (1, 3, 5), True),       # one valid tile + 1 position (halo 60 per side), three rows
    (32, 11, 2000, 2, (1, 3, 5), True),
    (32, 7, 5000, 1, (1, 3, 5), True),
    (32, 3, 457, 1, (1, 3, 5), True),
    (64, 3, 500, 2, (1, 3, 5), True),
    (64, 7, 1000, 1, (1, 3, 5), True),
    (32, 5, 900, 2, (1, 2, 4), True),        # vits2-style dilations
    (32, 3, 900, 1, (1, 1), False),          # no second conv (only dilation 1 is supported there)
    (32, 11, 40000, 1, (1, 3, 5), True),     # many tiles per CTA pair
]


@pytest.mark.parametrize("case", CHAIN_CASES)
def test_chain_block_vs_float64_block(case):
    C, k, L, B, dil, has2 = case
    n = len(dil) * (2 if has2 else 1)
    ws, bs = make_block(C, k, n, seed=C + k)
    x = torch.randn(B, C, L, generator=torch.Generator().manual_seed(L))
    y = run_chain(x, ws, bs, list(dil), has2)
    ref_q = block_f64(x, ws, bs, dil, has2, 0.1, lambda t: t.float().half().double()).float()
    ref = block_f64(x, ws, bs, dil, has2, 0.1, lambda t: t).float()
    # 16-bit rounding of an intermediate can flip by one ulp between two accumulation orders: 2e-3 absolute on O(1) data
    assert max_abs(y, ref_q) < 2.5e-3, case
    assert rel_l2(y, ref) <= 1e-3 and max_abs(y, ref) <= 1e-2, case
    # the block's edges (first / last 60 samples see the zero padding of every conv of the chain)
    e = min(64, L)
    assert max_abs(y[..., :e], ref_q[..., :e]) < 2.5e-3 and max_abs(y[..., -e:], ref_q[..., -e:]) < 2.5e-3


def test_chain_block_bf16_operands():
    ws, bs = make_block(32, 7, 6, seed=1)
    x = torch.randn(2, 32, 900, generator=torch.Generator().manual_seed(3))
    y = run_chain(x, ws, bs, [1, 3, 5], True, fp16=0)
    ref_q = block_f64(x, ws, bs, (1, 3, 5), True, 0.1, lambda t: t.float().bfloat16().double()).float()
    assert max_abs(y, ref_q) < 2e-2


def test_chain_rejects_blocks_it_cannot_run():
    lib = _lib.load()
    ws, bs = make_block(32, 3, 2, seed=0)
    with pytest.raises(RuntimeError):
        run_chain(torch.randn(1, 32, 64), ws, bs, [1, 3], False)        # dilated conv straight onto the residual
    ws, bs = make_block(16, 3, 6, seed=0)
    with pytest.raises(RuntimeError):
        run_chain(torch.randn(1, 16, 64), ws, bs, [1, 3, 5], True)      # 16 channels: per-unit kernels only
    assert lib is not None


@pytest.mark.parametrize("precision,rel_tol,abs_tol", [("fp32", 2e-5, 2e-5), ("fp16", 1e-3, 1e-2)])
def test_v1_full_size_against_reference_golden(precision, rel_tol, abs_tol):
    """(4, 80, 1000) -- BASELINE config C4's length -- against the unmodified reference (tests/golden/hifigan_v1_long.npz:
    a strided sample of the reference waveform): a tile-schedule bug at large L cannot hide in both of our own paths."""
    z = load_golden("hifigan_v1_long.npz")
    g = torch.Generator().manual_seed(int(z["seed"]))
    c = torch.randn(*[int(v) for v in z["shape"]], generator=g)
    assert abs(float(c.double().sum()) - float(z["c_sum"])) < 1e-6      # same input as the generating script drew
    torch.manual_seed(1234)
    m = vtts_b200.HiFiGAN()
    m.precision = precision
    m = m.to(DEV).eval()
    with torch.no_grad():
        y = m(c.to(DEV)).cpu()
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    got, ref = y[:, 0, idx], torch.from_numpy(z["y"])
    assert rel_l2(got, ref) <= rel_tol and max_abs(got, ref) <= abs_tol
    assert abs(float(y.double().norm()) - float(z["y_norm"])) <= 2 * rel_tol * float(z["y_norm"])


def test_chain_and_unit_kernels_mix_in_one_forward():
    """A 3-stage generator 128 / 64 / 32: the 32-channel stage runs the chain kernel, the 64-channel one keeps the
    per-unit kernels when one of its blocks does not fit (k = 11 at 64 channels); both must agree with the fp32 kernels."""
    torch.manual_seed(3)
    m = vtts_b200.HiFiGAN(channels=256, upsample_scales=[8, 4, 2], upsample_kernel_sizes=[16, 8, 4])
    m = m.to(DEV).eval()
    c = torch.randn(2, 80, 50, generator=torch.Generator().manual_seed(1)).to(DEV)
    with torch.no_grad():
        m.precision = "fp16"
        y = m(c)
        n_launch = m.last_launch_count
        m.precision = "fp32"
        ref = m(c)
    assert rel_l2(y, ref) <= 1e-3 and max_abs(y, ref) <= 1e-2
    assert n_launch <= 2 + 3 + 18 + 9 + 3 + 1      # layout + input conv, upsamples, stage 0 convs, stage 1 units, stage 2 chains, output conv

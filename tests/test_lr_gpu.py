"""GPU parity: LengthRegulator kernels vs the CPU oracle -- bit exact (integer/byte work)."""
import logging

import numpy as np
import pytest
import torch

import restate
import vtts_b200
from conftest import c_oracle_lr, load_golden, split_cases

pytestmark = pytest.mark.gpu
LR_CASES = split_cases(load_golden("lr_cases.npz"))
DEV = "cuda:0"


def bits(t: torch.Tensor) -> np.ndarray:
    a = t.detach().cpu().contiguous().numpy()
    return a.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


@pytest.mark.parametrize("name", sorted(LR_CASES))
def test_golden_cases_bit_exact(name):
    c = LR_CASES[name]
    lr = vtts_b200.LengthRegulator(pad_value=float(c["pad"]))
    ds = torch.from_numpy(c["ds"].copy()).to(DEV)
    out, mel_len = lr.forward_with_lengths(torch.from_numpy(c["xs"]).to(DEV), ds, float(c["alpha"]))
    assert tuple(out.shape) == c["out"].shape
    assert np.array_equal(bits(out), c["out"].view(np.uint32))
    # in-place mutation of the caller's ds on the all-zero path (layers.py:458)
    assert np.array_equal(ds.cpu().numpy(), c["ds_after"])
    out2 = lr(torch.from_numpy(c["xs"]).to(DEV), torch.from_numpy(c["ds"].copy()).to(DEV), float(c["alpha"]))
    assert torch.equal(out2, out)


@pytest.mark.parametrize("D", [80, 192, 256, 384, 3])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
def test_random_vs_numpy_and_c_oracle(D, dtype, lr_c_oracle):
    g = torch.Generator().manual_seed(D)
    B, T = 5, 37
    xs = torch.randn(B, T, D, generator=g).to(dtype)
    ds = torch.randint(0, 9, (B, T), generator=g)
    ds[2, 20:] = 0
    lr = vtts_b200.LengthRegulator(pad_value=0.5)
    out, mel_len = lr.forward_with_lengths(xs.to(DEV), ds.to(DEV))
    ref, _ = restate.lr_expand(xs.float() if dtype == torch.bfloat16 else xs, ds.clone(), 1.0, 0.5)
    assert torch.equal(out.cpu().float() if dtype == torch.bfloat16 else out.cpu(), ref)
    assert torch.equal(mel_len.cpu(), restate.lr_mel_len(ds))
    if dtype in (torch.float32, torch.float64):
        cref, _, clen = c_oracle_lr(lr_c_oracle, xs.numpy(), ds.numpy(), 1.0, 0.5)
        assert np.array_equal(bits(out), cref.view(bits(out).dtype))
        assert np.array_equal(mel_len.cpu().numpy(), clen)


def test_negative_duration_raises_like_repeat_interleave():
    lr = vtts_b200.LengthRegulator()
    ds = torch.tensor([[1, -1, 2]], device=DEV)
    with pytest.raises(RuntimeError, match="negative"):
        lr(torch.randn(1, 3, 4, device=DEV), ds)


def test_all_zero_batch_warns_and_mutates(caplog):
    lr = vtts_b200.LengthRegulator()
    ds = torch.zeros(3, 5, dtype=torch.long, device=DEV)
    xs = torch.randn(3, 5, 8, device=DEV)
    with caplog.at_level(logging.WARNING):
        out = lr(xs, ds)
    assert "all 0 sequences" in caplog.text
    assert torch.equal(ds, torch.ones_like(ds))
    assert torch.equal(out, xs)


def test_noncontiguous_inputs_and_max_len_extension():
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(4, 9, 32, generator=g).to(DEV)
    ds = torch.randint(0, 5, (9, 4), generator=g).to(DEV).t()  # non-contiguous view
    assert not ds.is_contiguous()
    lr = vtts_b200.LengthRegulator()
    out = lr(xs.transpose(1, 2).contiguous().transpose(1, 2), ds)
    ref, _ = restate.lr_expand(xs.cpu(), ds.cpu().clone())
    assert torch.equal(out.cpu(), ref)
    out2, ml = lr.forward_with_lengths(xs, ds, max_len=ref.shape[1] + 3)
    assert torch.equal(out2[:, : ref.shape[1]].cpu(), ref)
    assert torch.all(out2[:, ref.shape[1]:] == 0)


@pytest.mark.parametrize("B,Tmax,D", [(16, 120, 256), (64, 120, 384), (32, 170, 384), (256, 330, 256)])
def test_baseline_shapes_properties(B, Tmax, D):
    """BASELINE.json shapes: size-independent properties + sampled oracle comparison."""
    g = torch.Generator().manual_seed(B)
    xs = torch.randn(B, Tmax, D, generator=g)
    lens = torch.randint(40, Tmax + 1, (B,), generator=g)
    lens[0] = Tmax
    ds = torch.randint(1, 12, (B, Tmax), generator=g)
    ds[torch.arange(Tmax)[None, :] >= lens[:, None]] = 0
    lr = vtts_b200.LengthRegulator()
    out, mel_len = lr.forward_with_lengths(xs.to(DEV), ds.to(DEV))
    assert torch.equal(mel_len.cpu(), ds.sum(1))
    assert out.shape == (B, int(ds.sum(1).max()), D)
    # checksum property: sum over frames == sum_i d_i * x_i  (exact in float64 on the CPU)
    want = (xs.double() * ds[:, :, None].double()).sum(1)
    got = out.cpu().double().sum(1)
    assert torch.allclose(got, want, rtol=0, atol=1e-9)
    # every output frame equals a source frame, in non-decreasing token order
    ref, _ = restate.lr_expand(xs[:3], ds[:3].clone())
    assert torch.equal(out[:3, : ref.shape[1]].cpu(), ref)
    # idempotence with unit durations
    ones = (ds > 0).long()
    again = lr(xs.to(DEV), ones.to(DEV))
    for b in range(0, B, max(1, B // 4)):
        n = int(lens[b])
        assert torch.equal(again[b, :n].cpu(), xs[b, :n])


def test_expansion_is_differentiable_wrt_xs_like_the_reference():
    """layers.py:460-462 (repeat_interleave + pad_list) carries gradient from the decoder back to the encoder;
    the drop-in must too when it is installed under train.py (use_gaussian: false)."""
    import vtts_b200

    g = torch.Generator().manual_seed(11)
    B, Tmax, D = 3, 9, 16
    ds = torch.randint(0, 4, (B, Tmax), generator=g)
    ds[1, 5:] = 0
    xs = torch.randn(B, Tmax, D, generator=g)
    w = None
    grads = []
    for impl in ("ref", "new"):
        x = xs.clone().to("cuda:0").requires_grad_(True)
        d = ds.clone().to("cuda:0")
        if impl == "ref":
            rows = [torch.repeat_interleave(xi, di, dim=0) for xi, di in zip(x, d)]
            T = max(r.shape[0] for r in rows)
            out = torch.stack([torch.nn.functional.pad(r, (0, 0, 0, T - r.shape[0])) for r in rows])
        else:
            out = vtts_b200.LengthRegulator()(x, d)
            assert out.requires_grad
        if w is None:
            w = torch.randn(out.shape, generator=g).to("cuda:0")
        (out * w).sum().backward()
        grads.append((out.detach().cpu(), x.grad.cpu()))
    assert torch.equal(grads[0][0], grads[1][0])
    assert torch.allclose(grads[0][1], grads[1][1], atol=1e-6)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_randomised_shapes_bit_exact_vs_oracle(seed):
    """Random (B, Tmax, D, dtype, duration mix, pad) against the numpy oracle: the gather walks token runs per warp and
    switches vector width with the row size - every combination must stay bit-exact, including the padded tail."""
    import restate
    import vtts_b200

    g = torch.Generator().manual_seed(100 + seed)
    lr_mod = {}
    for case in range(25):
        B = int(torch.randint(1, 9, (1,), generator=g))
        Tmax = int(torch.randint(1, 70, (1,), generator=g))
        D = int(torch.randint(1, 400, (1,), generator=g)) if case % 3 else int(torch.randint(1, 12, (1,), generator=g)) * 32
        dmax = int(torch.randint(1, 40, (1,), generator=g))
        ds = torch.randint(0, dmax + 1, (B, Tmax), generator=g)
        ds[torch.rand(B, Tmax, generator=g) < 0.3] = 0
        if case % 5 == 0:
            ds[int(torch.randint(0, B, (1,), generator=g))] = 0            # an all-zero row next to non-zero ones
        if int(ds.sum()) == 0:
            ds[0, 0] = 1
        pad = float(torch.randn(1, generator=g)) if case % 2 else 0.0
        dtype = (torch.float32, torch.float16, torch.float64)[case % 3]
        xs = torch.randn(B, Tmax, D, generator=g).to(dtype)
        ref, _ = restate.lr_expand(xs, ds.clone(), pad_value=pad)
        m = lr_mod.setdefault(pad, vtts_b200.LengthRegulator(pad_value=pad))
        out = m(xs.to("cuda:0"), ds.to("cuda:0"))
        assert out.dtype == dtype and torch.equal(out.cpu(), ref), (case, B, Tmax, D, dtype, pad)

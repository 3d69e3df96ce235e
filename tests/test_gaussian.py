"""GaussianUpsampling: oracle pinned to the reference, CUDA kernel vs oracle.

Floating point: tolerance 1e-5 absolute on O(1) outputs (fp32 softmax + weighted sum; summation order differs).
"""
import logging

import numpy as np
import pytest
import torch

import restate
import vtts_b200
from conftest import load_golden, max_abs, split_cases

CASES = split_cases(load_golden("gaussian_cases.npz"))
TOL = 1e-5


def _args(c):
    hm = torch.from_numpy(c["h_masks"]) if "h_masks" in c else None
    dm = torch.from_numpy(c["d_masks"]) if "d_masks" in c else None
    return torch.from_numpy(c["hs"]), torch.from_numpy(c["ds"].copy()), hm, dm, float(c["delta"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    c = CASES[name]
    hs, ds, hm, dm, delta = _args(c)
    logging.disable(logging.WARNING)
    y = restate.gaussian_upsampling(hs, ds, hm, dm, delta)
    logging.disable(logging.NOTSET)
    assert y.shape == c["y"].shape
    assert max_abs(y, torch.from_numpy(c["y"])) < 1e-6
    assert np.array_equal(ds.numpy(), c["ds_after"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_kernel_matches_reference_golden(name):
    c = CASES[name]
    hs, ds, hm, dm, delta = _args(c)
    dev = "cuda:0"
    ds_d = ds.to(dev)
    logging.disable(logging.WARNING)
    y = vtts_b200.GaussianUpsampling(delta)(hs.to(dev), ds_d, None if hm is None else hm.to(dev),
                                            None if dm is None else dm.to(dev))
    logging.disable(logging.NOTSET)
    assert tuple(y.shape) == c["y"].shape
    assert max_abs(y, torch.from_numpy(c["y"])) < TOL
    assert np.array_equal(ds_d.cpu().numpy(), c["ds_after"])  # in-place fix-up on the caller's tensor


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,D", [(16, 120, 256), (8, 170, 384), (3, 37, 80)])
def test_kernel_vs_oracle_baseline_shapes(B, T, D):
    g = torch.Generator().manual_seed(B * T)
    hs = torch.randn(B, T, D, generator=g)
    tl = torch.randint(T // 3, T + 1, (B,), generator=g)
    tl[0] = T
    ds = torch.randint(1, 12, (B, T), generator=g)
    ds[torch.arange(T)[None] >= tl[:, None]] = 0
    ml = ds.sum(1)
    hm = torch.arange(int(ml.max()))[None] < ml[:, None]
    dm = torch.arange(T)[None] < tl[:, None]
    ref = restate.gaussian_upsampling(hs, ds.clone(), hm, dm, 0.1)
    dev = "cuda:0"
    y = vtts_b200.GaussianUpsampling(0.1)(hs.to(dev), ds.to(dev), hm.to(dev), dm.to(dev))
    assert max_abs(y, ref) < TOL
    # property: rows are convex combinations of token rows (weights sum to 1)
    ones = vtts_b200.GaussianUpsampling(0.1)(torch.ones(B, T, 32, device=dev), ds.to(dev), hm.to(dev), dm.to(dev))
    assert float((ones - 1).abs().max()) < 1e-5


def test_cpu_inputs_fail_loudly():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vtts_b200.GaussianUpsampling()(torch.randn(1, 3, 4), torch.ones(1, 3, dtype=torch.long))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_kernel_randomised_shapes_vs_oracle(seed):
    """Shapes that exercise every loop of the kernel: token windows longer than one staged chunk (32 tokens: short
    durations), feature widths beyond one pass (D > 256) and not a multiple of 32, small / large delta (window width),
    masks present or absent, frames beyond an utterance's length."""
    g = torch.Generator().manual_seed(50 + seed)
    for case in range(6):
        B = int(torch.randint(1, 5, (1,), generator=g))
        T = int(torch.randint(2, 150, (1,), generator=g))
        D = (17, 64, 256, 300, 384, 513)[case]
        dmax = (1, 2, 3, 8, 12, 30)[(case + seed) % 6]
        delta = (0.1, 0.02, 0.5)[(case + seed) % 3]
        tl = torch.randint(1, T + 1, (B,), generator=g)
        tl[0] = T
        ds = torch.randint(0, dmax + 1, (B, T), generator=g)
        ds[torch.arange(T)[None] >= tl[:, None]] = 0
        if int(ds.sum()) == 0:
            ds[0, 0] = 2
        hs = torch.randn(B, T, D, generator=g)
        masks = (case + seed) % 2 == 0
        ml = ds.sum(1)
        hm = (torch.arange(int(ml.max()))[None] < ml[:, None]) if masks else None
        dm = (torch.arange(T)[None] < tl[:, None]) if masks else None
        if masks and int(ml.min()) == 0:
            continue                      # a row without frames: softmax over an empty mask row is NaN in the reference too
        ref = restate.gaussian_upsampling(hs, ds.clone(), hm, dm, delta)
        with torch.no_grad():
            y = vtts_b200.GaussianUpsampling(delta=delta)(hs.to("cuda:0"), ds.to("cuda:0"), None if hm is None else hm.to("cuda:0"),
                                                          None if dm is None else dm.to("cuda:0"))
        assert y.shape == ref.shape
        assert max_abs(y, ref) <= TOL * max(1.0, float(ref.abs().max())), (case, B, T, D, dmax, delta, masks)

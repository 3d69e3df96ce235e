"""Pin the CPU oracle (oracle/restate.py, oracle/lr_oracle.c) to the reference.

Always: against the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py).  In the build container, additionally against the live reference.
"""
import logging

import numpy as np
import pytest
import torch

import restate
from conftest import c_oracle_lr, load_golden, max_abs, split_cases, state_dict_from

LR_CASES = split_cases(load_golden("lr_cases.npz"))


@pytest.mark.parametrize("name", sorted(LR_CASES))
def test_lr_numpy_oracle_matches_golden(name):
    c = LR_CASES[name]
    ds = torch.from_numpy(c["ds"].copy())
    out, _ = restate.lr_expand(torch.from_numpy(c["xs"]), ds, float(c["alpha"]), float(c["pad"]))
    assert out.shape == c["out"].shape
    assert np.array_equal(out.numpy().view(np.uint32), c["out"].view(np.uint32))  # bit exact
    assert np.array_equal(ds.numpy(), c["ds_after"])  # in-place quirk (layers.py:458)


@pytest.mark.parametrize("name", sorted(LR_CASES))
def test_lr_c_oracle_matches_golden(name, lr_c_oracle):
    c = LR_CASES[name]
    out, ds_after, mel_len = c_oracle_lr(lr_c_oracle, c["xs"], c["ds"], float(c["alpha"]), float(c["pad"]))
    assert out.shape == c["out"].shape
    assert np.array_equal(out.view(np.uint32), c["out"].view(np.uint32))
    if float(c["alpha"]) == 1.0:
        assert np.array_equal(ds_after, c["ds_after"])
    assert np.array_equal(mel_len, ds_after.sum(1))


def test_durations_from_log_matches_formula():
    # layers.py:205-208
    logd = torch.tensor([[-3.0, 0.0, 0.4, 1.0986123, 2.5]])
    d = restate.durations_from_log(logd, 1.0)
    assert d.tolist() == [[0, 0, 0, 2, 11]]
    assert restate.durations_from_log(logd, 1.5).tolist() == [[0, 0, 0, 3, 16]]


def test_hifigan_oracle_matches_golden_small():
    z = load_golden("hifigan_small.npz")
    sd = state_dict_from(z)
    kw = dict(upsample_scales=(4, 2), resblock_kernel_sizes=(3, 5), resblock_dilations=((1, 3), (1, 2)))
    y = restate.hifigan_forward(sd, torch.from_numpy(z["c"]), torch.from_numpy(z["g"]), **kw)
    assert max_abs(y, torch.from_numpy(z["y"])) < 1e-6
    y = restate.hifigan_forward(sd, torch.from_numpy(z["c"]), None, **kw)
    assert max_abs(y, torch.from_numpy(z["y_nog"])) < 1e-6


def test_hifigan_oracle_matches_golden_small2():
    z = load_golden("hifigan_small2.npz")
    y = restate.hifigan_forward(state_dict_from(z), torch.from_numpy(z["c"]), None, upsample_scales=(3, 2),
                                resblock_kernel_sizes=(3,), resblock_dilations=((1, 2, 3),), kernel_size=5,
                                use_additional_convs=False)
    assert y.shape == z["y"].shape
    assert max_abs(y, torch.from_numpy(z["y"])) < 1e-6


def test_vits2_oracle_matches_golden():
    z = load_golden("vits2_small.npz")
    for tag, rb, dil in (("rb1", "1", ((1, 3, 5), (1, 2, 4))), ("rb2", "2", ((1, 3), (1, 2)))):
        sd = state_dict_from(z, prefix=f"{tag}.sd.")
        y = restate.vits2_generator_forward(sd, torch.from_numpy(z[f"{tag}.x"]), torch.from_numpy(z[f"{tag}.g"]),
                                            resblock=rb, upsample_rates=(4, 2), upsample_kernel_sizes=(8, 4),
                                            resblock_kernel_sizes=(3, 7), resblock_dilation_sizes=dil)
        assert max_abs(y, torch.from_numpy(z[f"{tag}.y"])) < 1e-6


def test_fold_weight_norm_dims():
    # appendix 9.5: dim 0 = out-channels for Conv1d, in-channels for ConvTranspose1d
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(3, 5, 3))
    convt = torch.nn.utils.weight_norm(torch.nn.ConvTranspose1d(4, 2, 4, 2))
    for m in (conv, convt):
        w = restate.fold_weight_norm(m.weight_g.detach(), m.weight_v.detach())
        m(torch.zeros(1, m.in_channels, 8))  # run the pre-hook
        assert max_abs(w, m.weight) < 1e-7


# ---- live reference (build container only) ---------------------------------------------------
ref_loader = pytest.importorskip("ref_loader")
needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")


@needs_ref
def test_live_reference_lr_random():
    LR = ref_loader.load_length_regulator()
    g = torch.Generator().manual_seed(5)
    logging.disable(logging.WARNING)
    for trial in range(20):
        B, T, D = int(torch.randint(1, 6, (1,), generator=g)), int(torch.randint(1, 40, (1,), generator=g)), 8
        xs = torch.randn(B, T, D, generator=g)
        ds = torch.randint(0, 6, (B, T), generator=g)
        if trial % 5 == 0:
            ds.zero_()
        alpha = [1.0, 0.5, 1.5, 2.0][trial % 4]
        d1, d2 = ds.clone(), ds.clone()
        ref = LR(pad_value=0.25)(xs, d1, alpha)
        out, _ = restate.lr_expand(xs, d2, alpha, 0.25)
        assert torch.equal(ref, out)
        assert torch.equal(d1, d2)
    logging.disable(logging.NOTSET)


@needs_ref
def test_live_reference_hifigan_v1_matches_oracle_and_golden():
    HiFiGAN, _ = ref_loader.load_hifigan()
    z = load_golden("hifigan_v1.npz")
    torch.manual_seed(int(z["seed"]))
    m = HiFiGAN().eval()
    with torch.no_grad():
        y = m(torch.from_numpy(z["c"]))
    assert max_abs(y, torch.from_numpy(z["y"])) < 1e-6
    y2 = restate.hifigan_forward(m.state_dict(), torch.from_numpy(z["c"]))
    assert max_abs(y2, y) < 1e-5

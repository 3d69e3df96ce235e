"""vits2 duration path: generate_path + the attn matmuls of VITS2.inference.

Oracle pinned to the reference (golden vectors from models/gan_tts/vits2/utils.py run in the build container and, when
/root/reference is present, the live function); CUDA kernels vs golden and vs the oracle.  Bit-exact: the path is 0/1
valued and every output element of the expansion is a single product 1.0 * x (or 0).
"""
import numpy as np
import pytest
import torch

import ref_loader
import restate
import vtts_b200
from conftest import load_golden, split_cases

CASES = split_cases(load_golden("path_cases.npz"))
DEV = "cuda:0"


def _t(c, k):
    return torch.from_numpy(c[k])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    c = CASES[name]
    attn = restate.generate_path(_t(c, "w_ceil"), _t(c, "attn_mask"))
    assert torch.equal(attn, _t(c, "attn"))
    out = restate.expand_by_path(_t(c, "m_p"), _t(c, "w_ceil"), _t(c, "attn_mask"))
    assert torch.equal(out, _t(c, "m_out"))


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference")
def test_oracle_matches_live_reference_on_random_inputs():
    U = ref_loader.load_vits2_utils()
    g = torch.Generator().manual_seed(21)
    for b, t_x in [(1, 1), (2, 17), (5, 64)]:
        w = torch.ceil(torch.rand(b, 1, t_x, generator=g) * 6 - 1).clamp_min(0)
        t_y = int(w.sum(-1).max()) + 3
        mask = (torch.rand(b, 1, t_y, t_x, generator=g) > 0.2).float()
        assert torch.equal(restate.generate_path(w, mask), U.generate_path(w, mask))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_kernels_match_reference_golden(name):
    c = CASES[name]
    w, mask, m_p = _t(c, "w_ceil").to(DEV), _t(c, "attn_mask").to(DEV), _t(c, "m_p").to(DEV)
    attn = vtts_b200.generate_path(w, mask)
    assert tuple(attn.shape) == c["attn"].shape and attn.dtype == torch.float32
    assert torch.equal(attn.cpu(), _t(c, "attn"))
    out = vtts_b200.expand_by_path(m_p, w, mask)
    assert tuple(out.shape) == c["m_out"].shape
    assert torch.equal(out.cpu(), _t(c, "m_out"))


@pytest.mark.gpu
@pytest.mark.parametrize("b,t_x,d", [(16, 120, 192), (64, 200, 192), (1, 1, 3), (3, 300, 7)])
def test_kernels_vs_oracle_and_properties(b, t_x, d):
    g = torch.Generator().manual_seed(b * t_x)
    x_len = torch.randint(max(1, t_x // 3), t_x + 1, (b,), generator=g)
    x_len[0] = t_x
    x_mask = (torch.arange(t_x)[None] < x_len[:, None]).float().unsqueeze(1)
    w_ceil = torch.ceil(torch.rand(b, 1, t_x, generator=g) * 8) * x_mask
    y_len = w_ceil.sum([1, 2]).long().clamp_min(1)
    t_y = int(y_len.max())
    y_mask = (torch.arange(t_y)[None] < y_len[:, None]).float().unsqueeze(1)
    mask = x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)
    m_p = torch.randn(b, d, t_x, generator=g)
    attn = vtts_b200.generate_path(w_ceil.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(attn, restate.generate_path(w_ceil, mask))
    out = vtts_b200.expand_by_path(m_p.to(DEV), w_ceil.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(out, restate.expand_by_path(m_p, w_ceil, mask))
    # properties: every valid frame selects exactly one token; token x is repeated w_ceil[x] times, in order
    assert torch.equal(attn.sum(-1).squeeze(1), y_mask.squeeze(1))
    assert torch.equal(attn.sum(2).squeeze(1), w_ceil.squeeze(1))
    ids = torch.arange(t_x, dtype=torch.float32).view(1, 1, t_x).expand(b, 1, t_x).contiguous()
    idx = vtts_b200.expand_by_path(ids.to(DEV), w_ceil.to(DEV), mask.to(DEV)).cpu().squeeze(1)
    for r in range(b):
        want = torch.repeat_interleave(torch.arange(t_x, dtype=torch.float32), w_ceil[r, 0].long())
        assert torch.equal(idx[r, : int(y_len[r])], want[: int(y_len[r])]) or int(w_ceil[r].sum()) == 0


@pytest.mark.gpu
def test_negative_durations_follow_the_formula():
    """Outside the call site's contract (durations are ceil(exp(.)) >= 0) the path has +-1 entries; the kernels follow
    the same formula as the reference."""
    w = torch.tensor([[[2.0, -1.0, 3.0, 0.0, 1.0]]])
    mask = torch.ones(1, 1, 7, 5)
    x = torch.arange(10, dtype=torch.float32).view(1, 2, 5)
    assert torch.equal(vtts_b200.generate_path(w.to(DEV), mask.to(DEV)).cpu(), restate.generate_path(w, mask))
    got = vtts_b200.expand_by_path(x.to(DEV), w.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(got, restate.expand_by_path(x, w, mask))


def test_cpu_inputs_fail_loudly():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vtts_b200.generate_path(torch.ones(1, 1, 3), torch.ones(1, 1, 4, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vtts_b200.expand_by_path(torch.ones(1, 2, 3), torch.ones(1, 1, 3), torch.ones(1, 1, 4, 3))

"""One ResidualBlock through the fused chain kernel (vtts_dbg_resblock_chain) vs a CPU emulation that rounds the
operands to 16 bits at the same points (isolates kernel correctness from quantisation) and vs plain fp32.

usage: python tools/chain_block.py [C k L B [reps]]   (no arguments: a fixed list of cases)
"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import torch
import torch.nn.functional as F
from vtts_b200 import _lib

DEV = "cuda:0"


def make_block(C, k, n_units, has2, seed):
    g = torch.Generator().manual_seed(seed)
    n = n_units * (2 if has2 else 1)
    ws = [torch.randn(C, C, k, generator=g) / (C * k) ** 0.5 for _ in range(n)]
    bs = [0.1 * torch.randn(C, generator=g) for _ in range(n)]
    return ws, bs


def reference(x, ws, bs, dil, has2, slope, rnd):
    """layers.py:83-98 in fp64; rnd() models the 16-bit operand rounding of the kernel (identity = exact)."""
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


def run_chain(x, ws, bs, dil, has2, slope=0.1, fp16=1, reps=0):
    lib = _lib.load()
    B, C, L = x.shape
    k = ws[0].shape[-1]
    xd = x.to(DEV).contiguous()
    wd = [w.to(DEV).contiguous() for w in ws]
    bd = [b.to(DEV).contiguous() for b in bs]
    y = torch.empty_like(xd)
    n = len(ws)
    wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in wd])
    bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bd])
    dl = (ctypes.c_int * len(dil))(*dil)
    ms = ctypes.c_float(0.0)
    _lib.check(lib.vtts_dbg_resblock_chain(xd.data_ptr(), wp, bp, y.data_ptr(), B, C, L, k, dl, len(dil), 1 if has2 else 0,
                                           slope, fp16, reps, ctypes.byref(ms), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return y.cpu(), ms.value


def check(C, k, L, B, dil=(1, 3, 5), has2=True, fp16=1, seed=0):
    ws, bs = make_block(C, k, len(dil), has2, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, C, L, generator=g)
    y, _ = run_chain(x, ws, bs, list(dil), has2, fp16=fp16)
    rnd = (lambda t: t.float().half().double()) if fp16 else (lambda t: t.float().bfloat16().double())
    ref_q = reference(x, ws, bs, dil, has2, 0.1, rnd).float()
    ref = reference(x, ws, bs, dil, has2, 0.1, lambda t: t).float()
    eq = float((y - ref_q).abs().max())
    e = float((y - ref).abs().max())
    rel = float((y - ref).norm() / ref.norm())
    bad = (y - ref_q).abs() > 5e-3
    msg = f"C={C} k={k} L={L} B={B} dil={dil} has2={has2} fp16={fp16}: max|y-ref_q| {eq:.2e}  max|y-ref| {e:.2e} rel {rel:.2e}"
    if bad.any():
        idx = bad.nonzero()
        msg += f"  BAD {int(bad.sum())} first {idx[0].tolist()} last {idx[-1].tolist()}"
        pos = idx[:, 2]
        msg += f" pos range [{int(pos.min())},{int(pos.max())}] distinct mod4 {sorted(set((pos % 4).tolist()))}"
    print(msg, flush=True)
    return eq


if __name__ == "__main__":
    torch.zeros(1).to(DEV)
    if len(sys.argv) >= 5:
        C, k, L, B = (int(a) for a in sys.argv[1:5])
        reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
        ws, bs = make_block(C, k, 3, True, 0)
        x = torch.randn(B, C, L)
        y, ms = run_chain(x, ws, bs, [1, 3, 5], True, reps=reps)
        fl = 2.0 * C * C * k * 6 * B * L
        by = B * L * C * 8.0
        print(f"C={C} k={k} L={L} B={B}: {ms*1e3:.1f} us/launch, {fl/ms/1e9:.1f} TFLOP/s useful, {by/ms/1e6:.1f} GB/s (x in + y out)")
    else:
        worst = 0.0
        for case in [(32, 3, 100, 1), (32, 3, 1000, 2), (32, 7, 777, 1), (32, 11, 2000, 2), (32, 11, 361, 3),
                     (32, 3, 457, 1), (32, 7, 5000, 1), (64, 3, 500, 2), (64, 7, 1000, 1), (32, 11, 40000, 1)]:
            worst = max(worst, check(*case))
        worst = max(worst, check(32, 3, 900, 1, dil=(1, 1), has2=False))
        worst = max(worst, check(32, 5, 900, 2, dil=(1, 2, 4)))
        worst = max(worst, check(32, 7, 900, 2, fp16=0))
        print("WORST", worst)

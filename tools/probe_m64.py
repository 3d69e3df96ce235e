"""Hardware probe: where do the 64 rows of an M=64 tcgen05.mma accumulator live in TMEM, and does the assumed
tcgen05.ld 16x256b fragment layout hold?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
from vtts_b200 import _lib
lib = _lib.load()
N, K = 64, 64
g = torch.Generator().manual_seed(0)
a = torch.randn(128, K, generator=g).bfloat16().cuda()
b = torch.randn(64, K, generator=g).bfloat16().cuda()
ref = a.float() @ b.float().t()          # (128, N)
def run(variant):
    d = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(lib.vtts_dbg_umma_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), 128, N, K, 64, 0, variant, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d
d128 = run(0)
print("M=128 32x32b readout max err", (d128 - ref).abs().max().item())
d128f = run(8)
print("M=128 16x256b readout max err", (d128f - ref).abs().max().item())
for variant in (4, 12):
    d = run(variant)
    lanes = []
    for r in range(64):
        hit = [l for l in range(128) if torch.allclose(d[l], ref[r], atol=1e-2, rtol=1e-2)]
        lanes.append(hit)
    print(f"variant {variant}: M=64 row -> TMEM lane:", lanes)

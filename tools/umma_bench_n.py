"""tcgen05.mma cycles per instruction as a function of N (and M): is the cost linear in N?"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import torch
from vtts_b200 import _lib
lib = _lib.load()
torch.zeros(1).cuda()
out = np.zeros(2, dtype=np.int64)
reps = 2000
for rowb in (128, 64):
    ks = rowb // 32
    for M in (128, 64):
        for N in (256, 240, 224, 208, 192, 176, 160, 144, 128, 112, 96, 64, 32):
            _lib.check(lib.vtts_dbg_umma_bench(N, rowb, 1, reps, M, 0, out.ctypes.data))
            n = reps * ks
            print(f"rowb={rowb} M={M} N={N}: complete {out[1]/n:.1f} cyc/MMA ({out[1]/n/N:.3f} cyc/col)")

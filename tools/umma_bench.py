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
      for N in (256, 128, 64):
        for shift in (0, 1, 5):
            for two in (0, 1):
                _lib.check(lib.vtts_dbg_umma_bench(N, rowb, shift, reps, M, two, out.ctypes.data))
                n = reps * ks
                print(f"rowb={rowb} M={M} N={N} shift={shift} two_acc={two}: issue {out[0]/n:.1f} cyc/MMA, complete {out[1]/n:.1f} cyc/MMA "
                      f"-> {2*M*N*16/(out[1]/n):.0f} FLOP/cyc/SM (floor 128*N/256={128*N/256:.0f} cyc)")

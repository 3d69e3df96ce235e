"""Per-tile pipeline trace (block 0) of single-conv launches inside one untrimmed HiFi-GAN V1 forward, e.g. the upsamples.

usage: python tools/trace_upsample.py [launch ordinals ...]     (default: the four upsamples; ordinals as in profiles/r02_launch_list.csv)
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
from vtts_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
gen = vtts_b200.HiFiGAN().cuda().eval()
c = torch.randn(16, 80, 759, generator=torch.Generator().manual_seed(1)).cuda()
with torch.no_grad():
    gen(c)
    torch.cuda.synchronize()
    for n in [int(a) for a in sys.argv[1:]] or [2, 21, 31, 35]:
        _lib.check(lib.vtts_dbg_trace(100 + n, None, 0))
        gen(c)
        torch.cuda.synchronize()
        buf = np.zeros(64 * 16, dtype=np.int64)
        _lib.check(lib.vtts_dbg_trace(0, buf.ctypes.data, buf.size))
        t = buf.reshape(64, 16)
        t0 = t[0, 15]
        n_t = int((t[:, 3] != 0).sum())
        print(f"--- launch {n}: {n_t} tiles on CTA 0; cycles relative to kernel entry; previous kernel done (grid_dep_wait) at {t[0, 14] - t0}, "
              f"all roles done at {t[1, 15] - t0}")
        print("tile | mma:start accE_ok actF_ok issued | act:wait go | w:wait go | epi:wait accF_ok done")
        for i in range(0, min(n_t, 10)):
            r = t[i] - t0
            print(f"{i:4d} | {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d} | {r[6]:8d} {r[7]:8d} | {r[8]:8d} {r[9]:8d} {r[10]:8d}")
        live = [i for i in range(4, 40) if t[i, 3] and t[i + 1, 3]]
        if len(live) > 1:
            a, b = live[0], live[-1] + 1
            print(f"cycles per tile (MMA issue done), tiles {a}..{b}: {np.diff(t[a:b + 1, 3]).mean():.0f};  epilogue busy {np.mean(t[a:b, 10] - t[a:b, 9]):.0f};  "
                  f"mma wait act {np.mean(t[a:b, 2] - t[a:b, 1]):.0f}; mma wait accE {np.mean(t[a:b, 1] - t[a:b, 0]):.0f}; issue {np.mean(t[a:b, 3] - t[a:b, 2]):.0f}")

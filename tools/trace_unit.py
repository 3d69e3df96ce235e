"""Per-tile pipeline trace of fused ResidualBlock-unit launches inside one HiFi-GAN V1 forward (block 0 of each launch).

usage: python tools/trace_unit.py [launch ordinals ...]   (stage1 units 22..30, stage2 32..40, stage3 42..50; j*3+m order)
"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
from vtts_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
gen = vtts_b200.HiFiGAN().cuda().eval()
B, T = 16, 759
g = torch.Generator().manual_seed(1)
lens = torch.randint(300, T + 1, (B,), generator=g); lens[0] = T
c = torch.randn(B, 80, T, generator=g).cuda()
lens_d = lens.cuda()
with torch.no_grad():
    gen.forward_trimmed(c, lens_d)
    torch.cuda.synchronize()
    for n in [int(a) for a in sys.argv[1:]] or [42, 48, 28, 30]:
        _lib.check(lib.vtts_dbg_trace(100 + n, None, 0))
        gen.forward_trimmed(c, lens_d)
        torch.cuda.synchronize()
        buf = np.zeros(64 * 16, dtype=np.int64)
        _lib.check(lib.vtts_dbg_trace(0, buf.ctypes.data, buf.size))
        t = buf.reshape(64, 16)
        t0 = t[2, 0]
        print(f"--- launch {n}: cycles relative to tile-2 start")
        print("tile | mma: start  A_issued xt_full  B_empty  B_issued | EA: accA_ok done | EB: pre accB_ok done")
        for i in list(range(2, 6)) + list(range(40, 48)):
            r = t[i] - t0
            print(f"{i:4d} | {r[0]:9d} {r[1]:9d} {r[2]:8d} {r[3]:8d} {r[4]:9d} | {r[5]:8d} {r[6]:8d} | {r[7]:8d} {r[8]:8d} {r[9]:8d}")
        if t[3, 10] != 0:
            e = t[3:12]
            print(f"  conv1 issuer: waits {np.mean(e[:,10]-e[:,0]):.0f}  mma issue {np.mean(e[:,11]-e[:,10]):.0f}  commits {np.mean(e[:,1]-e[:,11]):.0f}  A period {np.diff(e[:,0]).mean():.0f} | conv2 issuer: B period {np.diff(e[:,4]).mean():.0f}")
        print("  A-issuer period by tile range:", " ".join(f"[{a}:{b}] {np.diff(t[a:b,0]).mean():.0f}" for a, b in ((3, 12), (12, 24), (24, 40), (40, 62))),
              "| EB done period:", " ".join(f"[{a}:{b}] {np.diff(t[a:b,9]).mean():.0f}" for a, b in ((3, 12), (12, 24), (24, 40), (40, 62))))
        e = t[3:12]
        print(f"period {np.diff(e[:,0]).mean():.0f} | A issue {np.mean(e[:,1]-e[:,0]):.0f}  wait xt {np.mean(e[:,2]-e[:,1]):.0f}  wait Bempty {np.mean(e[:,3]-e[:,2]):.0f}"
              f"  B issue {np.mean(e[:,4]-e[:,3]):.0f} | EA: accA_full after A-issued {np.mean(e[:,5]-e[:,1]):.0f}  EA busy {np.mean(e[:,6]-e[:,5]):.0f}"
              f" | EB: accB_full after B-issued {np.mean(e[:,8]-e[:,4]):.0f}  EB busy {np.mean(e[:,9]-e[:,8]):.0f}")

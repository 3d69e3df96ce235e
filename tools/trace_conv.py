"""Per-tile pipeline trace of one tcgen05 conv layer (block 0): where does each role wait?"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
from vtts_b200 import _lib
lib = _lib.load()
def run(B, C, L, k, d, with_res, want_x=True):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C, L, generator=g).cuda(); w = (torch.randn(C, C, k, generator=g) * 0.05).cuda()
    bias = torch.randn(C, generator=g).cuda(); res = torch.randn(B, C, L, generator=g).cuda() if with_res else None
    y = torch.empty(B, C, L, device="cuda"); ya = torch.empty(B, C, L, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def call():
        _lib.check(lib.vtts_dbg_conv1d_tc(x.data_ptr(), w.data_ptr(), bias.data_ptr(), 0 if res is None else res.data_ptr(),
                                          y.data_ptr() if want_x else 0, ya.data_ptr(), B, C, C, L, k, d, 0.1, 0.1, 1, st))
    call()
    _lib.check(lib.vtts_dbg_trace(1, None, 0))
    call()
    buf = np.zeros(64 * 16, dtype=np.int64)
    _lib.check(lib.vtts_dbg_trace(0, buf.ctypes.data, buf.size))
    t = buf.reshape(64, 16)
    t0 = t[0, 0]
    print(f"--- C={C} L={L} k={k} d={d} res={with_res} want_x={want_x}  (cycles relative to first MMA-thread stamp)")
    print("tile | mma:start accE_ok actF_ok issued | act:wait go | w:wait go | epi:wait accF_ok done")
    for i in list(range(4, 14)):
        r = t[i] - t0
        print(f"{i:4d} | {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d} | {r[6]:8d} {r[7]:8d} | {r[8]:8d} {r[9]:8d} {r[10]:8d}")
    e = t[4:16]
    steps = k * max(1, C // 64)
    print(f"MMA-thread per step ({steps} steps/tile): wait_w {np.mean(e[:,11])/steps:.0f}  fence+issue {np.mean(e[:,12])/steps:.0f}  commit {np.mean(e[:,13])/steps:.0f} cycles")
    per = np.diff(t[4:16, 3]).mean()
    print(f"steady-state cycles per tile (MMA issue done): {per:.0f};  epilogue busy {np.mean(t[4:16,10]-t[4:16,9]):.0f};  "
          f"mma wait act {np.mean(t[4:16,2]-t[4:16,1]):.0f}; mma wait accE {np.mean(t[4:16,1]-t[4:16,0]):.0f}; issue {np.mean(t[4:16,3]-t[4:16,2]):.0f}")
if __name__ == "__main__":
    run(16, 32, 194304, 3, 1, False, want_x=False)
    run(16, 32, 194304, 3, 1, True)
    run(16, 128, 48576, 11, 1, False, want_x=False)
    run(16, 128, 48576, 3, 1, True)

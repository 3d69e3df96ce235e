import sys, torch, ctypes
sys.path.insert(0, "viet-transformer-tts_b200")
import vtts_b200
from vtts_b200 import _lib
lib = _lib.load()
def run(B, T, D, masks=True, dmax=12, reps=20):
    g = torch.Generator().manual_seed(0)
    hs = torch.randn(B, T, D, generator=g).cuda()
    tl = torch.randint(T // 3, T + 1, (B,), generator=g); tl[0] = T
    ds = torch.randint(1, dmax, (B, T), generator=g); ds[torch.arange(T)[None] >= tl[:, None]] = 0
    ml = ds.sum(1); Tf = int(ml.max())
    hm = (torch.arange(Tf)[None] < ml[:, None]).to(torch.uint8).cuda().contiguous()
    dm = (torch.arange(T)[None] < tl[:, None]).to(torch.uint8).cuda().contiguous()
    dsd = ds.cuda(); out = torch.empty(B, Tf, D, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def call():
        _lib.check(lib.vtts_gauss_upsample(hs.data_ptr(), dsd.data_ptr(), hm.data_ptr() if masks else None, dm.data_ptr() if masks else None,
                                           out.data_ptr(), B, T, D, Tf, 0.1, st))
    for _ in range(3): call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): call()
    b.record(); torch.cuda.synchronize()
    print(f"B={B} T={T} D={D} masks={masks} Tf={Tf}: {a.elapsed_time(b)/reps*1e3:.1f} us")
run(16, 120, 256); run(16, 120, 256, masks=False); run(16, 120, 64); run(16, 30, 256); run(1, 120, 256); run(64, 120, 256); run(16, 120, 256, dmax=3)

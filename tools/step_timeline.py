"""Kernel-level timeline of one bench-shaped Synthesizer step (torch profiler / CUPTI): where does the step go?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
torch.manual_seed(1234)
gen = vtts_b200.HiFiGAN().cuda().eval()
synth = vtts_b200.Synthesizer(gen)
g = torch.Generator().manual_seed(0)
B, T, D = 16, 120, 256
tl = torch.randint(40, 121, (B,), generator=g); tl[0] = T
ds = torch.randint(1, 12, (B, T), generator=g)
ds[torch.arange(T)[None] >= tl[:, None]] = 0
hs = torch.randn(B, T, D, generator=g).pin_memory(); ds = ds.pin_memory()
for _ in range(3):
    synth(hs, ds)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        synth(hs, ds)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    k = e.name.split("(")[0][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{us / 3:10.1f} us/step  x{n // 3:3d}  {k}")
print(f"sum of GPU activity {tot / 3:.1f} us/step")
ts = sorted((e.time_range.start, e.time_range.end) for e in ev)
span = (ts[-1][1] - ts[0][0]) / 3
print(f"first-to-last GPU activity span {span:.1f} us/step (includes inter-step host time)")

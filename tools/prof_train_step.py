import sys, torch
sys.path.insert(0, "viet-transformer-tts_b200")
import vtts_b200
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(1234)
m = vtts_b200.HiFiGAN().cuda(); m.train_backend = "tc"
c = torch.randn(16, 80, 64, device="cuda")
def step():
    m.zero_grad(set_to_none=True); m(c).abs().mean().backward()
for _ in range(2): step()
torch.cuda.synchronize()
import time
t0=time.perf_counter(); y = m(c); torch.cuda.synchronize(); t1=time.perf_counter(); y.abs().mean().backward(); torch.cuda.synchronize(); t2=time.perf_counter()
print("forward %.1f ms backward %.1f ms (wall)" % ((t1-t0)*1e3, (t2-t1)*1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))

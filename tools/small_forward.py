"""Small HiFi-GAN V1 forward (all tensor-core kernels, with and without padding trim) for compute-sanitizer memcheck."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
torch.manual_seed(0)
m = vtts_b200.HiFiGAN().cuda().eval()
c = torch.randn(3, 80, 37, device="cuda")
lens = torch.tensor([37, 11, 23], device="cuda")
with torch.no_grad():
    y0 = m(c)
    y1 = m.forward_trimmed(c, lens)
torch.cuda.synchronize()
ok = all(torch.equal(y0[b, :, : int(lens[b]) * 256], y1[b, :, : int(lens[b]) * 256]) for b in range(3))
print("ok", tuple(y0.shape), "trim identical:", ok, "finite:", bool(torch.isfinite(y0).all()))

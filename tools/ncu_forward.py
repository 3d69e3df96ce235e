"""Small driver for ncu: HiFi-GAN V1 forward at the bench shape (B=16, T=759), twice."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 759
torch.manual_seed(1234)
m = vtts_b200.HiFiGAN(); m.precision = prec; m = m.cuda().eval()
c = torch.randn(B, 80, T, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = m(c)
torch.cuda.synchronize()
print("ok", tuple(y.shape), m.last_launch_count)

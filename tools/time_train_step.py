"""Forward + backward of HiFi-GAN V1 at the trainer's shape (batch 16 x 64-frame segments, hifigan_trainer.py:143-167):
train_backend "tc" (tcgen05 forward / dgrad, cuBLAS wgrad) vs "eager" (PyTorch / cuDNN, fp32 and TF32)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
torch.manual_seed(1234)
m = vtts_b200.HiFiGAN().cuda()
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 64)
c = torch.randn(B, 80, T, device="cuda")
def step():
    m.zero_grad(set_to_none=True)
    m(c).abs().mean().backward()
def ms(n=5):
    for _ in range(2): step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for backend, tf32 in (("tc", False), ("eager", False), ("eager", True)):
    m.train_backend = backend
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    print(f"B={B} T={T} backend={backend} tf32={tf32}: {ms():.1f} ms per forward+backward")

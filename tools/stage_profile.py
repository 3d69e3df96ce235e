"""Stage-boundary timing of the bench-shaped generator forward, trimmed and untrimmed (run with VTTS_STAGE_PROFILE=1)."""
import os, sys, torch
sys.path.insert(0, "viet-transformer-tts_b200"); sys.path.insert(0, ".")
import vtts_b200
from bench import make_workload
torch.manual_seed(1234)
gen = vtts_b200.HiFiGAN().cuda().eval(); gen.precision = "fp16"
lr = vtts_b200.LengthRegulator()
hs, ds = make_workload(seed=0, B=16)
with torch.no_grad():
    frames, mel_len = lr.forward_with_lengths(hs.cuda(), ds.cuda())
    mel = frames[..., :80].transpose(1, 2).contiguous()
    for _ in range(3):
        gen.forward_trimmed(mel, mel_len)
    torch.cuda.synchronize()
    print("--- trimmed", flush=True)
    for _ in range(2):
        gen.forward_trimmed(mel, mel_len)
    torch.cuda.synchronize()
    print("--- untrimmed")
    for _ in range(2):
        gen(mel)
    torch.cuda.synchronize()

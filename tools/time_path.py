"""CUDA-event timing of the vits2 path kernels at (16, 192, 120 -> ~760)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
g = torch.Generator().manual_seed(0)
B, D, T = 16, 192, 120
w = (torch.rand(B, 1, T, generator=g) * 12).ceil()
x_len = torch.randint(40, T + 1, (B,), generator=g); x_len[0] = T
x_mask = (torch.arange(T)[None] < x_len[:, None]).float().unsqueeze(1)
w = (w * x_mask).cuda()
y_len = w.sum([1, 2]).long()
y_mask = (torch.arange(int(y_len.max()), device="cuda")[None] < y_len[:, None]).float().unsqueeze(1)
attn_mask = (x_mask.cuda().unsqueeze(2) * y_mask.unsqueeze(-1))
m_p = torch.randn(B, D, T, generator=g).cuda()
def ms(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print("generate_path  %.1f us" % ms(lambda: vtts_b200.generate_path(w, attn_mask)))
print("expand_by_path %.1f us  (out %s)" % (ms(lambda: vtts_b200.expand_by_path(m_p, w, attn_mask)), tuple(vtts_b200.expand_by_path(m_p, w, attn_mask).shape)))

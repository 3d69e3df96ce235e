"""CPU emulation of the bf16 tensor-core numerics (operands rounded to bf16, fp32 accumulate,
fp32 residual stream, fp32 output conv) to separate scheme error from kernel bugs."""
import sys, os
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import restate, vtts_b200

def q(t, mode):
    if mode == "bf16": return t.bfloat16().float()
    if mode == "split2":  # hi + lo bf16 (two MMAs)
        hi = t.bfloat16().float(); return hi + (t - hi).bfloat16().float()
    return t

def fwd(sd, c, act_mode="bf16", w_mode="bf16", post_fp32=True, stage_modes=None):
    W = lambda p: restate._weight(sd, p); Bv = lambda p: sd.get(p + ".bias")
    def conv(x, p, am, wm, **kw): return F.conv1d(q(x, am), q(W(p), wm), Bv(p), **kw)
    x = conv(c, "input_conv", act_mode, w_mode, padding=3)
    scales = (8, 8, 2, 2); ks = (3, 7, 11); dil = (1, 3, 5)
    stages = [x]
    for i, s in enumerate(scales):
        am, wm = (stage_modes[i] if stage_modes else (act_mode, w_mode))
        x = F.conv_transpose1d(q(F.leaky_relu(x, 0.1), am), q(W(f"upsamples.{i}.1"), wm), Bv(f"upsamples.{i}.1"), stride=s, padding=s // 2)
        stages.append(x)
        cs = 0.0
        for j, k in enumerate(ks):
            y = x
            for m, d in enumerate(dil):
                p = f"blocks.{i*3+j}"
                xt = conv(F.leaky_relu(y, 0.1), f"{p}.convs1.{m}.1", am, wm, padding=(k - 1) // 2 * d, dilation=d)
                xt = conv(F.leaky_relu(xt, 0.1), f"{p}.convs2.{m}.1", am, wm, padding=(k - 1) // 2)
                y = xt + y
            cs = cs + y
        x = cs / 3
        stages.append(x)
    pm = ("fp32", "fp32") if post_fp32 else (act_mode, w_mode)
    y = torch.tanh(conv(F.leaky_relu(x, 0.01), "output_conv.1", pm[0], pm[1], padding=3))
    return y, stages

def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())

if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    torch.manual_seed(1234)
    m = vtts_b200.HiFiGAN()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    c = torch.randn(2, 80, T, generator=g)
    ref, rst = restate.hifigan_forward(sd, c, return_stages=True)
    for name, kw in [("all bf16 operands, fp32 post", {}),
                     ("act split2, w bf16", dict(act_mode="split2")),
                     ("act bf16, w split2", dict(w_mode="split2")),
                     ("both split2", dict(act_mode="split2", w_mode="split2")),
                     ("stage3 split2/split2 only", dict(stage_modes=[("bf16","bf16")]*3+[("split2","split2")])),
                     ("stage2+3 split2/split2", dict(stage_modes=[("bf16","bf16")]*2+[("split2","split2")]*2)),
                     ("stage3 act split2 only", dict(stage_modes=[("bf16","bf16")]*3+[("split2","bf16")])),
                     ]:
        y, st = fwd(sd, c, **kw)
        print(f"{name:32s} wav rel-L2 {rel(y, ref):.3e} max-abs {float((y-ref).abs().max()):.2e} | stages " +
              " ".join(f"{rel(a,b):.1e}" for a, b in zip(st, rst)))

import sys, torch
sys.path.insert(0, "viet-transformer-tts_b200")
import torch.nn.functional as F
import vtts_b200
import vtts_b200.training as T
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
torch.manual_seed(7)
m = vtts_b200.HiFiGAN(in_channels=80, channels=128, global_channels=16, upsample_scales=[4, 2], upsample_kernel_sizes=[8, 4],
                      resblock_kernel_sizes=[3, 7], resblock_dilations=[[1, 3], [1, 3]]).to(DEV).train()
g = torch.Generator().manual_seed(1)
c = torch.randn(3, 80, 24, generator=g).to(DEV).requires_grad_(True)
gc = torch.randn(3, 16, 1, generator=g).to(DEV)
w_out = torch.randn(3, 1, 192, generator=g).to(DEV)
def grads(backend):
    m.zero_grad(set_to_none=True); c.grad = None
    m.train_backend = backend
    y = m(c, gc); (y * w_out).sum().backward()
    return {n: p.grad.clone() for n, p in m.named_parameters()}
ref = grads("eager")
tc = grads("tc")
orig = T.conv1d_tc
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
# same graph, exact fp32 convs
T.conv1d_tc = lambda x, w, b, d, p="fp16": F.conv1d(x, w, b, padding=(w.shape[-1] - 1) // 2 * d, dilation=d)
ex = grads("tc")
# same graph, operands rounded to fp16 in forward only (backward exact)
class R(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return x.half().float()
    @staticmethod
    def backward(ctx, g): return g.half().float()      # round gradient too
T.conv1d_tc = lambda x, w, b, d, p="fp16": R.apply(F.conv1d(R.apply(x), R.apply(w), b, padding=(w.shape[-1] - 1) // 2 * d, dilation=d)) if False else F.conv1d(R.apply(x), R.apply(w), b, padding=(w.shape[-1] - 1) // 2 * d, dilation=d)
em = grads("tc")
T.conv1d_tc = orig
for n in ["blocks.3.convs1.0.1.weight_v", "blocks.1.convs1.1.1.weight_v", "input_conv.weight_v", "blocks.0.convs1.0.1.weight_v"]:
    print(n, "tc-vs-eager %.2e | exact-graph-vs-eager %.2e | emulated-rounding-vs-eager %.2e | tc-vs-emulated %.2e" % (rel(tc[n], ref[n]), rel(ex[n], ref[n]), rel(em[n], ref[n]), rel(tc[n], em[n])))

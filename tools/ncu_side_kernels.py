"""Small driver for ncu: the side kernels of the path at bandwidth-regime shapes.
usage: python tools/ncu_side_kernels.py lr|gauss|path
  lr     LengthRegulator (256, 330, 256) -> ~2200 frames      (lr_rowsum_kernel, lr_gather_kernel)
  gauss  GaussianUpsampling (16, 120, 256) -> ~760 frames      (gauss_upsample_kernel)
  path   vits2 generate_path + expansion (16, 192, 120) -> ~760 (path_generate_kernel, path_expand_kernel)
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
what = sys.argv[1] if len(sys.argv) > 1 else "lr"
g = torch.Generator().manual_seed(0)
dev = "cuda"
if what == "lr":
    B, T, D = 256, 330, 256
    hs = torch.randn(B, T, D, generator=g).to(dev)
    tl = torch.randint(T // 3, T + 1, (B,), generator=g); tl[0] = T
    ds = torch.randint(1, 13, (B, T), generator=g); ds[torch.arange(T)[None] >= tl[:, None]] = 0
    ds = ds.to(dev)
    lr = vtts_b200.LengthRegulator()
    for _ in range(2):
        out = lr(hs, ds)
    print("ok", tuple(out.shape))
elif what == "gauss":
    B, T, D = 16, 120, 256
    hs = torch.randn(B, T, D, generator=g).to(dev)
    tl = torch.randint(40, T + 1, (B,), generator=g); tl[0] = T
    ds = torch.randint(1, 12, (B, T), generator=g); ds[torch.arange(T)[None] >= tl[:, None]] = 0
    ml = ds.sum(1)
    hm = (torch.arange(int(ml.max()))[None] < ml[:, None]).to(dev)
    dm = (torch.arange(T)[None] < tl[:, None]).to(dev)
    gu = vtts_b200.GaussianUpsampling()
    with torch.no_grad():
        for _ in range(2):
            out = gu(hs, ds.to(dev), hm, dm)
    print("ok", tuple(out.shape))
else:
    B, D, T = 16, 192, 120
    w = (torch.rand(B, 1, T, generator=g) * 12).ceil()
    x_len = torch.randint(40, T + 1, (B,), generator=g); x_len[0] = T
    x_mask = (torch.arange(T)[None] < x_len[:, None]).float().unsqueeze(1)
    w = w * x_mask
    y_len = w.sum([1, 2]).long()
    y_mask = (torch.arange(int(y_len.max()))[None] < y_len[:, None]).float().unsqueeze(1)
    attn_mask = (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).to(dev)
    m_p = torch.randn(B, D, T, generator=g).to(dev)
    for _ in range(2):
        path = vtts_b200.generate_path(w.to(dev), attn_mask)
        out = vtts_b200.expand_by_path(m_p, w.to(dev), attn_mask)
    print("ok", tuple(out.shape), tuple(path.shape))
torch.cuda.synchronize()

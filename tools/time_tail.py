"""Where the acoustic tail's time goes at the C2 shape (B = 16 x 759 frames): CUDA-event timing of its pieces."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
from vtts_b200.acoustic import _operand
dev = "cuda"
cfg = {"decoder_head": 2, "conv_filter_size": 1024, "conv_kernel_size": [9, 1], "decoder_dropout": 0.2}
torch.manual_seed(1234)
tail = vtts_b200.AcousticTail(vtts_b200.Decoder(4, 256, 1000, cfg), torch.nn.Linear(256, 80),
                              vtts_b200.Postnet(80, {"embedding_dim": 512, "conv_layers": 5, "kernel_size": 5})).to(dev).eval()
B, T = 16, 759
frames = torch.randn(B, T, 256, device=dev)
mel_len = torch.randint(300, T + 1, (B,), device=dev); mel_len[0] = T
mask = torch.arange(T, device=dev)[None] >= mel_len[:, None]

def ms(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

with torch.no_grad():
    blk = tail.decoder.layer_stack[0]
    x = frames
    am = mask.unsqueeze(1).expand(-1, T, -1)
    print("whole tail            %.3f ms" % ms(lambda: tail(frames, mel_len)))
    print("decoder (4 blocks)    %.3f ms" % ms(lambda: tail.decoder(frames, mask)))
    print(" one block            %.3f ms" % ms(lambda: blk(x, mask=mask, slf_attn_mask=am, need_weights=False)))
    print("  attention (fused)   %.3f ms" % ms(lambda: blk.slf_attn(x, x, x, mask=am, need_weights=False)))
    print("  attention (ref ops) %.3f ms" % ms(lambda: blk.slf_attn(x, x, x, mask=am, need_weights=True)))
    print("  pos_ffn             %.3f ms" % ms(lambda: blk.pos_ffn(x)))
    c1, c2 = blk.pos_ffn._convs()
    a16 = _operand(x, "fp16", 256)
    print("   operand cast       %.3f ms" % ms(lambda: _operand(x, "fp16", 256)))
    print("   w_1 conv (k=9)     %.3f ms" % ms(lambda: c1.run(a16, "fp16", False, True, slope_out=0.0)))
    hid = c1.run(a16, "fp16", False, True, slope_out=0.0)[1]
    print("   w_2 conv (k=1)     %.3f ms" % ms(lambda: c2.run(hid, "fp16", True, False, res=x)))
    print("   layer_norm         %.3f ms" % ms(lambda: blk.pos_ffn.layer_norm(x)))
    outs = tail.feats_linear(x)
    print("feats_linear          %.3f ms" % ms(lambda: tail.feats_linear(x)))
    print("postnet               %.3f ms" % ms(lambda: tail.postnet(outs)))
    tc = tail.postnet._convs()
    a0 = _operand(outs, "fp16", tc[0].padded_channels(torch.device(dev)))
    print(" conv 80->512         %.3f ms" % ms(lambda: tc[0].run(a0, "fp16", False, True, act_tanh=True)))
    a1 = tc[0].run(a0, "fp16", False, True, act_tanh=True)[1]
    print(" conv 512->512        %.3f ms" % ms(lambda: tc[1].run(a1, "fp16", False, True, act_tanh=True)))
    print(" conv 512->80         %.3f ms" % ms(lambda: tc[4].run(a1, "fp16", True, False)))

"""Single-utterance latency of HiFiGAN.forward on (1, 80, 200): launch path and CUDA-graph replay."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
import vtts_b200
torch.manual_seed(0)
m = vtts_b200.HiFiGAN().cuda().eval()
c = torch.randn(1, 80, 200, device="cuda")
def lat(fn, n=30):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort(); return round(ts[len(ts) // 2], 3), round(ts[0], 3)
with torch.no_grad():
    for _ in range(5):
        m(c)
    torch.cuda.synchronize()
    a = lat(lambda: m(c))
    gf = m.graphed(c)
    for _ in range(3):
        gf(c)
    b = lat(lambda: gf(c))
    ts = []
    for _ in range(30):                      # host time of one launch-path call (no synchronisation inside)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); m(c); ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    print("host time of m(c) without sync: median ms", round(ts[len(ts) // 2], 3))
    import cProfile, pstats, io
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20):
        m(c)
    pr.disable(); torch.cuda.synchronize()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(14); print(st.getvalue()[:2500])
print("env", {k: v for k, v in os.environ.items() if k.startswith("VTTS_")}, "launch median/min", a, "graph median/min", b)

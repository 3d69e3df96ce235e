"""Per-step wall times of the pipelined Synthesizer.submit() loop (bench.py's e2e leg), one line per rank: finds host-side
stalls (allocator, pinned memory, other tenants of the box) that a 10-step average hides.  Run under torchrun or alone."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
sys.path.insert(0, ROOT)
import vtts_b200
from bench import make_workload
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(1234)
gen = vtts_b200.HiFiGAN().cuda().eval()
synth = vtts_b200.Synthesizer(gen)
hs, ds = make_workload(seed=0, B=16)
hs, ds = hs.pin_memory(), ds.pin_memory()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
with torch.no_grad():
    for _ in range(3):
        synth(hs, ds)
    for _ in range(2):
        synth.submit(hs, ds).result()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ts, sub = [], []
    pending = None
    t_prev = time.perf_counter()
    for k in range(steps):
        t0 = time.perf_counter()
        nxt = synth.submit(hs, ds)
        t1 = time.perf_counter()
        if pending is not None:
            pending.result()
        pending = nxt
        t2 = time.perf_counter()
        sub.append(1e3 * (t1 - t0)); ts.append(1e3 * (t2 - t_prev)); t_prev = t2
    pending.result()
print(f"rank {rank}: step ms " + " ".join(f"{t:.1f}" for t in ts))
print(f"rank {rank}: submit() host ms " + " ".join(f"{t:.1f}" for t in sub))
print(f"rank {rank}: allocator: " + str({k: v for k, v in torch.cuda.memory_stats().items() if k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "reserved_bytes.all.current")}))
if world > 1:
    dist.destroy_process_group()

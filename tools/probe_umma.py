"""Hardware probe: pin tcgen05 shared-memory descriptor semantics (row-shifted operand, SW128/SW64).

Writes gpurun_out/probe_umma.json.  Each case runs in this process; a trap poisons the context,
so the driver script runs variants in separate processes (see --variant).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))
from vtts_b200 import _lib  # noqa: E402


def run_case(lib, N, K, shift, variant, seed=0):
    g = torch.Generator().manual_seed(seed)
    rows_b = ((N + shift + 63) // 64) * 64
    a = torch.randn(128, K, generator=g).bfloat16().cuda()
    b = torch.randn(rows_b, K, generator=g).bfloat16().cuda()
    d = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.vtts_dbg_umma_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), 128, N, K, rows_b, shift, variant,
                                torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        return {"rc": rc, "err": lib.vtts_last_error().decode()}
    torch.cuda.synchronize()
    ref = a.float() @ b.float()[shift:shift + N].t()
    err = (d - ref).abs().max().item()
    return {"rc": 0, "max_err": err, "ref_scale": ref.abs().max().item(), "ok": bool(err < 1e-2 * max(1.0, ref.abs().max().item()))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, required=True)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe_umma.json"))
    args = ap.parse_args()
    lib = _lib.load()
    results = []
    ch = 32 if (args.variant & 2) else 64
    for N in (32, 128, 256):
        for kb in (1, 3):
            for shift in (0, 8, 1, 3, 13, 50):
                try:
                    r = run_case(lib, N, kb * ch, shift, args.variant)
                except Exception as e:  # CUDA error (trap) -> stop this process
                    r = {"rc": -99, "err": str(e)[:200]}
                    results.append({"variant": args.variant, "N": N, "kb": kb, "shift": shift, **r})
                    print(json.dumps(results[-1]), flush=True)
                    _dump(args.out, results)
                    return 1
                results.append({"variant": args.variant, "N": N, "kb": kb, "shift": shift, **r})
                print(json.dumps(results[-1]), flush=True)
    _dump(args.out, results)
    return 0


def _dump(path, results):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    prev = []
    if os.path.exists(path):
        try:
            prev = json.load(open(path))
        except Exception:
            prev = []
    json.dump(prev + results, open(path, "w"), indent=0)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- synthesized audio seconds per second (inverse RTF) of the hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|fp32]

Workload (BASELINE.json configs[1]): FastSpeech2 (4-layer, 256-hidden) LengthRegulator +
HiFi-GAN V1, batch 16 synthetic phoneme sequences per GPU: hidden states (16, 120, 256) fp32,
integer durations 1..11 inside each utterance's text length (40..120), 0 outside ->
LengthRegulator -> (acoustic decoder stand-in: first 80 features as the mel) -> HiFi-GAN V1
-> 22.05 kHz waveform.  Random-init weights drawn with torch.manual_seed(1234).

One "step" = one pass of the hot path over one batch.  `value` counts only VALID audio
(sum_b mel_len_b * 256 / 22050), not the padded tail of shorter utterances.

Under torchrun (N > 1) every rank runs its own batch (weak scaling, no data-path collective);
time = max over ranks, value = all ranks' audio / that time.

`--impl reference` times the reference's algorithm on the host CPU instead (rank 0 only):
the reference is pure Python and cannot travel to the GPU box, so the arm runs the CPU oracle
port (oracle/restate.py -- the same torch.nn.functional calls the reference modules make) with
all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "viet-transformer-tts_b200"))

SAMPLE_RATE = 22050
HOP = 256
FLOP_PER_FRAME_V1 = 614_105_088  # BASELINE.md section 3 (2*MAC, convs only, in_channels = 80)


def make_workload(seed: int, B: int = 16, Ttext: int = 120, D: int = 256):
    """SURVEY.md section 8d, C2 recipe (LR unit-bench variant: ds = randint(1,12) inside length)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    hs = torch.randn(B, Ttext, D, generator=g)
    text_len = torch.randint(40, Ttext + 1, (B,), generator=g)
    text_len[0] = Ttext
    ds = torch.randint(1, 12, (B, Ttext), generator=g)
    ds[torch.arange(Ttext)[None, :] >= text_len[:, None]] = 0
    return hs, ds


class ClockSampler:
    """Samples SM clock and throttle reasons while the timed region runs (pynvml)."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                try:
                    phys = int(visible.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_oracle_run(hs, ds, sd_folded, n_utts: int, repeats: int):
    """Time the CPU oracle port on the first `n_utts` utterances.  Returns (audio_s, seconds/run)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import restate

    hs_s, ds_s = hs[:n_utts].clone(), ds[:n_utts].clone()
    audio = float(ds_s.sum()) * HOP / SAMPLE_RATE
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            frames, _ = restate.lr_expand(hs_s, ds_s.clone())
            mel = frames[..., :80].transpose(1, 2).contiguous()
            restate.hifigan_forward(sd_folded, mel)
            times.append(time.perf_counter() - t0)
    return audio, times


def fold_state_dict(sd):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import restate

    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            p = k[: -len(".weight_g")]
            out[p + ".weight"] = restate.fold_weight_norm(v.detach().cpu(), sd[p + ".weight_v"].detach().cpu())
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v.detach().cpu()
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    import vtts_b200

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    gen = vtts_b200.HiFiGAN()
    sd = fold_state_dict(gen.state_dict())
    hs, ds = make_workload(0)
    n_utts = args.ref_utts
    for _ in range(args.warmup):
        cpu_oracle_run(hs, ds, sd, n_utts, 1)
    audio, times = cpu_oracle_run(hs, ds, sd, n_utts, args.steps)
    total = sum(times)
    value = audio * args.steps / total
    sample = f"first {n_utts} of 16 utterances per step ({audio:.2f} s audio), fp32, weight-norm pre-folded"
    line = {
        "impl": "reference", "metric": "synthesized audio sec/sec (inverse RTF)", "value": value,
        "unit": "audio_s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: LengthRegulator (16,120,256) + HiFi-GAN V1, 22.05 kHz, batch 16/GPU",
                   "reference_arm": "CPU oracle port of the reference modules, bounded sample"},
        "cpu_baseline": {"value": value, "unit": "audio_s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio_s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=None, choices=[None, "fp16", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--ref-utts", type=int, default=2, help="utterances per step for the CPU arms")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager (cuDNN) timing of the module tree")
    ap.add_argument("--no-trim", action="store_true", help="process the padded tail of short utterances too")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import vtts_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    precision = args.precision or vtts_b200.hifigan.DEFAULT_PRECISION
    torch.manual_seed(1234)  # config/train_config.yaml:1 -- identical weights on every rank
    gen = vtts_b200.HiFiGAN()
    gen.precision = precision
    gen = gen.to(dev).eval()
    lr = vtts_b200.LengthRegulator()
    synth = vtts_b200.Synthesizer(gen, lr, trim_padding=not args.no_trim)

    hs, ds = make_workload(seed=rank, B=args.batch)
    hs_pin, ds_pin = hs.pin_memory(), ds.pin_memory()
    hs_d, ds_d = hs.to(dev), ds.to(dev)
    valid_frames = int(ds.sum())
    audio_s = valid_frames * HOP / SAMPLE_RATE
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    trim = not args.no_trim and precision != "fp32"

    def run_gen(mel, mel_len):
        return gen.forward_trimmed(mel, mel_len) if trim else gen(mel)

    def step_device():
        frames, mel_len = lr.forward_with_lengths(hs_d, ds_d)
        mel = frames[..., :80].transpose(1, 2)
        wav = run_gen(mel, mel_len)
        return frames, wav

    def step_generator_only(mel):
        return gen(mel)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for _ in range(args.warmup):
            frames, wav = step_device()
            synth(hs_pin, ds_pin)
        torch.cuda.synchronize(dev)
        T_out = frames.shape[1]
        padded_frames = frames.shape[0] * T_out
        launches_per_step = 2 + gen.last_launch_count

        # ---- device-resident timing: K steps, L2 flushed between steps (outside the events) ----
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        gev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        with ClockSampler(local) as clocks:
            for k in range(args.steps):
                flush.fill_(k & 0xFF)
                ev[k][0].record()
                frames, mel_len = lr.forward_with_lengths(hs_d, ds_d)
                mel = frames[..., :80].transpose(1, 2)
                gev[k][0].record()
                wav = run_gen(mel, mel_len)
                gev[k][1].record()
                ev[k][1].record()
            barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        gen_ms = [a.elapsed_time(b) for a, b in gev]
        total_ms = sum(step_ms)

        # transparency: the same generator pass WITHOUT the padding trim (module-level forward, strict parity)
        uev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 5))]
        mel_full = frames[..., :80].transpose(1, 2)
        for a_, b_ in uev:
            flush.fill_(1)
            a_.record()
            gen(mel_full)
            b_.record()
        torch.cuda.synchronize(dev)
        untrimmed_ms = sum(a_.elapsed_time(b_) for a_, b_ in uev) / len(uev)

        # LengthRegulator alone (HBM-bandwidth class, launch-latency bound at these sizes): kernels only
        # (max_len given -> no host read), algorithmic bytes = xs + ds read, frames + mel_len written
        lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a_, b_ in lev:
            flush.fill_(2)
            a_.record()
            lr.forward_with_lengths(hs_d, ds_d, max_len=T_out)
            b_.record()
        torch.cuda.synchronize(dev)
        lr_us = sorted(1e3 * a_.elapsed_time(b_) for a_, b_ in lev)[len(lev) // 2]
        lr_bytes = hs_d.numel() * 4 + ds_d.numel() * 8 + frames.numel() * 4 + ds_d.shape[0] * 8
        # single-utterance latency at configs[0] (B=1, T=200 frames = 2.3 s of audio): blocking module call, wall clock
        mel1 = torch.randn(1, 80, 200, device=dev)
        for _ in range(3):
            gen(mel1)
        torch.cuda.synchronize(dev)
        lat = []
        for _ in range(10):
            t0 = time.perf_counter()
            gen(mel1)
            torch.cuda.synchronize(dev)
            lat.append(1e3 * (time.perf_counter() - t0))
        lat.sort()
        gf = gen.graphed(mel1)                    # the same forward replayed from a CUDA graph (no per-launch host cost)
        for _ in range(3):
            gf(mel1)
        torch.cuda.synchronize(dev)
        lat_g = []
        for _ in range(10):
            t0 = time.perf_counter()
            gf(mel1)
            torch.cuda.synchronize(dev)
            lat_g.append(1e3 * (time.perf_counter() - t0))
        lat_g.sort()

        # context (SURVEY 8d "the real bar"): the same nn.Module tree run eagerly by PyTorch/cuDNN on this GPU -- what the
        # reference's own modules do on a B200.  This is the shells' autograd/eager path (hifigan.py:_forward_eager),
        # outside every timed region above; rank 0 only.
        eager = None
        if rank == 0 and not args.no_eager_baseline:
            eager = {}
            old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            for name, tf32 in (("tf32", True), ("fp32", False)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                for _ in range(2):
                    gen._forward_eager(mel_full)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
                for a_, b_ in ev:
                    a_.record()
                    gen._forward_eager(mel_full)
                    b_.record()
                torch.cuda.synchronize(dev)
                ms = sorted(a_.elapsed_time(b_) for a_, b_ in ev)[1]
                eager[f"generator_ms_{name}"] = ms
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32
            eager["note"] = ("PyTorch eager (cuDNN) forward of the same module tree, padded batch, no trim; compare with "
                             "config.generator_ms_untrimmed")

        # ---- end-to-end through the public API with host buffers (H2D + D2H inside) -------------
        e2e_ms = []
        barrier()
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            wav_h, wav_len_h = synth(hs_pin, ds_pin)  # synchronises before returning
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
        barrier()
        e2e_sync_total = sum(e2e_ms)
        # pipelined form (Synthesizer.submit): same per-step host->device inputs and device->host waveform, but the
        # read-back of step k runs on a copy stream behind the kernels of step k+1.  No L2 flush inside this region
        # (it cannot be excluded from the clock here); the per-step working set is far larger than L2 anyway.
        for _ in range(2):
            synth.submit(hs_pin, ds_pin).result()
        torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        pending = None
        for k in range(args.steps):
            nxt = synth.submit(hs_pin, ds_pin)
            if pending is not None:
                wav_h, wav_len_h = pending.result()
            pending = nxt
        wav_h, wav_len_h = pending.result()
        e2e_total = 1e3 * (time.perf_counter() - t0)
        barrier()

    t = torch.tensor([total_ms, e2e_total, e2e_sync_total], dtype=torch.float64, device=dev)
    a = torch.tensor([audio_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
    total_ms_all, e2e_ms_all, e2e_sync_ms_all = t.tolist()
    audio_all = float(a.item())

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        gen_avg_ms = sum(gen_ms) / len(gen_ms)
        margin = gen.TRIM_MARGIN_FRAMES
        # algorithmic work of the trimmed step = the valid frames only (the few look-ahead frames per utterance that
        # the kernels still compute behind mel_len are overhead, not counted)
        needed_frames = valid_frames if trim else padded_frames
        flops = needed_frames * FLOP_PER_FRAME_V1
        achieved = flops / (gen_avg_ms * 1e-3) / 1e12
        # DRAM bytes of the same launch set: ncu launch list of the untrimmed forward at this shape (profiles/), scaled by
        # the fraction of frames the trimmed step computes (traffic is proportional to the tiles processed)
        traffic_bytes, traffic_note = None, "no ncu capture committed"
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
                tj = json.load(fh)
            traffic_bytes = tj["dram_bytes_per_forward"] * needed_frames / float(16 * 759)
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum over the 52 launches of one forward "
                            "(profiles/r01_launch_list.csv, B=16 x 759 frames untrimmed), scaled by frames computed / 12144")
        except Exception:
            pass
        value = audio_all * args.steps / (total_ms_all * 1e-3)
        line = {
            "metric": "synthesized audio sec/sec (inverse RTF)", "value": value, "unit": "audio_s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_all / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp16": "fp16", "bf16": "bf16", "fp32": "f32"}[precision], "data": "synthetic",
            "config": {
                "workload": "configs[1]: LengthRegulator (16,120,256) + HiFi-GAN V1, 22.05 kHz, batch 16/GPU",
                "precision": precision, "batch_per_gpu": args.batch, "padded_mel_frames": padded_frames,
                "valid_mel_frames": valid_frames, "l2": "256 MiB flush buffer written between timed steps; "
                "per-step activation working set also exceeds the 126 MB L2",
                "value_counts": "valid (unpadded) audio only",
                "padding_trim": (f"on: generator tiles beyond mel_len + per-layer receptive-field margin (<= {margin} frames) "
                                 "skipped (valid samples bit-identical, tests/test_tc_gpu.py)" if trim else "off"),
                "frames_needed_for_roofline": needed_frames,
                "generator_ms_untrimmed": untrimmed_ms,
                "untrimmed_generator_tflops": padded_frames * FLOP_PER_FRAME_V1 / (untrimmed_ms * 1e-3) / 1e12,
                "torch_eager_gpu": eager,
                "length_regulator": {"us": lr_us, "algorithmic_bytes": lr_bytes, "GBps": lr_bytes / (lr_us * 1e-6) / 1e9,
                                     "note": "kernels only (rowsum + gather), output length supplied; launch-latency bound"},
                "latency_ms_configs0_b1_t200": {"median": lat[len(lat) // 2], "min": lat[0],
                                                "cuda_graph_median": lat_g[len(lat_g) // 2],
                                                "note": "HiFiGAN.forward on (1,80,200), blocking, wall clock; cuda_graph = "
                                                        "HiFiGAN.graphed() replay of the same launches"},
            },
            "clocks": clocks.summary(),
            "e2e": {"value": audio_all * args.steps / (e2e_ms_all * 1e-3), "unit": "audio_s/s",
                    "h2d_bytes_per_step": hs.numel() * 4 + ds.numel() * 8,
                    "d2h_bytes_per_step": int(wav.numel()) * 4 + ds.shape[0] * 8,
                    "ms_per_step": e2e_ms_all / args.steps,
                    "mode": "Synthesizer.submit(): read-back of step k overlaps the kernels of step k+1; every step's "
                            "inputs and waveform cross PCIe inside the timed region",
                    "sync_value": audio_all * args.steps / (e2e_sync_ms_all * 1e-3),
                    "sync_note": "Synthesizer.__call__(): one blocking call per step (H2D, kernels, D2H, sync)"},
            "gpu_launches": launches_per_step * args.steps * world,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                         "frac": achieved / tc_peak, "traffic": traffic_bytes,
                         "kernel": "generator conv kernels (all launches of one forward)",
                         "flops_per_launch_set": flops, "ms": gen_avg_ms, "peak_source": peak_src,
                         "traffic_source": traffic_note},
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            sd = fold_state_dict(gen.state_dict())
            cpu_oracle_run(hs, ds, sd, 1, 1)  # warm-up
            caudio, ctimes = cpu_oracle_run(hs, ds, sd, args.ref_utts, 3)
            ctimes.sort()
            line["cpu_baseline"] = {
                "value": caudio / ctimes[len(ctimes) // 2], "unit": "audio_s/s", "cores": cores, "kind": "port",
                "sample": f"first {args.ref_utts} of {args.batch} utterances ({caudio:.2f} s audio), median of 3, "
                          "CPU oracle port (torch fp32, weight-norm pre-folded)",
            }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

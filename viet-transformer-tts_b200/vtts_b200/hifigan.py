"""Drop-in ``HiFiGAN`` / ``ResidualBlock`` whose synthesis forward runs in libvtts_b200.so.

API mirror of the reference's ESPnet-style generator:
  * ``HiFiGAN``        -- models/gan_tts/hifigan/generator.py:16-213
  * ``ResidualBlock``  -- models/gan_tts/hifigan/layers.py:16-98

The shells keep the reference's constructor signature, attribute names, sub-module tree and
therefore its ``state_dict()`` keys (``input_conv.weight_g`` ... 234 tensors for V1), so
checkpoints written by the reference trainers load unchanged.  The sub-modules are ordinary
``nn.Conv1d`` / ``nn.ConvTranspose1d`` objects used as *parameter containers*: the synthesis
path never calls them, it hands their parameters to the C ABI, which folds the weight norm and
packs them once per parameter version.

Policy for autograd (hard part 6): when gradients are required (training,
``hifigan_trainer.py:143``) the forward runs the sub-modules through PyTorch autograd exactly
like the reference does -- backward kernels are a later row of the scope table.  Everything
else (``torch.no_grad()`` / ``inference()`` / parameters frozen) runs the CUDA kernels, and
raises if the library or a CUDA device is missing; there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib

# "fp16": tcgen05 with fp16 operands (default: meets rel-L2 <= 1e-3), "bf16": same kernels with bf16
# operands (rel-L2 ~3e-3 on random-init V1, as the CPU emulation in tools/emulate_bf16.py predicts),
# "fp32": CUDA-core reference path.
DEFAULT_PRECISION = os.environ.get("VTTS_B200_PRECISION", "fp16")


def _act(name: str, params: Dict[str, Any]) -> nn.Module:
    return getattr(nn, name)(**params)


class ResidualBlock(nn.Module):
    """Three (dilated conv -> plain conv) units with identity skips (layers.py:16-98)."""

    def __init__(
        self,
        kernel_size: int = 3,
        channels: int = 512,
        dilations: List[int] = [1, 3, 5],
        bias: bool = True,
        use_additional_convs: bool = True,
        nonlinear_activation: str = "LeakyReLU",
        nonlinear_activation_params: Dict[str, Any] = {"negative_slope": 0.1},
    ):
        super().__init__()
        assert kernel_size % 2 == 1, "Kernel size must be odd number."
        self.use_additional_convs = use_additional_convs
        self.kernel_size = kernel_size
        self.dilations = list(dilations)
        half = (kernel_size - 1) // 2

        def unit(dil: int) -> nn.Sequential:
            return nn.Sequential(
                _act(nonlinear_activation, nonlinear_activation_params),
                nn.Conv1d(channels, channels, kernel_size, 1, dilation=dil, bias=bias, padding=half * dil),
            )

        # construction order (convs1[i] then convs2[i]) matches the reference so that a given
        # torch seed draws identical parameters
        self.convs1 = nn.ModuleList()
        if use_additional_convs:
            self.convs2 = nn.ModuleList()
        for d in self.dilations:
            self.convs1.append(unit(d))
            if use_additional_convs:
                self.convs2.append(unit(1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Autograd/eager form (B, C, T) -> (B, C, T); the fused CUDA path lives in HiFiGAN."""
        for i in range(len(self.convs1)):
            y = self.convs1[i](x)
            if self.use_additional_convs:
                y = self.convs2[i](y)
            x = y + x
        return x


class _GeneratorBase(nn.Module):
    """Shared machinery: handle lifetime, weight upload, workspace, kernel forward."""

    precision: str = DEFAULT_PRECISION
    train_backend: str = "eager"      # autograd path: "eager" = the module tree through PyTorch (reference behaviour),
                                      # "tc" = every conv forward / dgrad on the tcgen05 kernel, wgrad on cuBLAS (training.py)

    def _init_runtime(self) -> None:
        self._handles: Dict[int, int] = {}          # device index -> VttsGen*
        self._uploaded: Dict[int, tuple] = {}       # device index -> parameter version signature
        self._workspace: Dict[int, torch.Tensor] = {}
        self.last_launch_count = 0

    # subclasses provide -------------------------------------------------------------------
    def _gen_config(self) -> _lib.VttsGenConfig:  # pragma: no cover - abstract
        raise NotImplementedError

    def _layer_modules(self) -> List[nn.Module]:  # pragma: no cover - abstract
        raise NotImplementedError

    def _forward_eager(self, c, g=None):  # pragma: no cover - abstract
        raise NotImplementedError

    # ----------------------------------------------------------------------------------------
    def __getstate__(self):
        # device handles / workspaces are per-object runtime state: never copied or pickled
        state = self.__dict__.copy()
        for k in ("_handles", "_uploaded", "_workspace"):
            state[k] = {}
        state.pop("_layer_cache", None)
        return state

    def __del__(self):
        try:
            lib = _lib.load()
            for h in getattr(self, "_handles", {}).values():
                lib.vtts_gen_destroy(h)
        except Exception:
            pass

    _PARAM_NAMES = ("weight_g", "weight_v", "weight", "bias")

    def _signature(self) -> tuple:
        """Version signature of every tensor the kernels consume (re-upload / re-capture trigger).

        Reads the parameters straight from the layer modules' ``_parameters`` dicts (weight-norm removal and
        parameter replacement change the entries; in-place updates bump ``_version``; ``.to()`` changes ``data_ptr``)
        instead of walking ``self.parameters()``: 10x cheaper, and this runs on every forward.
        """
        mods = self.__dict__.get("_layer_cache")
        if mods is None:                      # the layer modules are fixed after construction; indexing ModuleLists is slow
            mods = self._layer_modules()
            self.__dict__["_layer_cache"] = mods
        sig = [("epoch", self.__dict__.get("_epoch", 0), 0)]
        for m in mods:
            ps = m._parameters
            for n in self._PARAM_NAMES:
                t = ps.get(n)
                if t is not None:
                    sig.append((n, t.data_ptr(), t._version))
        return tuple(sig)

    def invalidate(self) -> None:
        """Force a re-upload (and CUDA-graph re-capture) on the next forward.

        The automatic trigger (:meth:`_signature`) sees parameter replacement and in-place autograd-visible updates
        (optimizer steps, ``load_state_dict``).  Writes through ``.data`` (``p.data.copy_(ema)``) do not bump the version
        counter: call this after them.  ``reset_parameters`` / ``remove_weight_norm`` / ``apply_weight_norm`` /
        ``load_state_dict`` call it themselves.
        """
        self._uploaded.clear()
        self.__dict__["_epoch"] = self.__dict__.get("_epoch", 0) + 1

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        if "_uploaded" in self.__dict__:
            self.invalidate()

    def _handle(self, dev: torch.device) -> int:
        lib = _lib.load()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            cfg = self._gen_config()
            out = ctypes.c_void_p()
            _lib.check(lib.vtts_gen_create(ctypes.byref(cfg), ctypes.byref(out)))
            h = out.value
            self._handles[idx] = h
            mods = self._layer_modules()
            n = _lib.check(lib.vtts_gen_num_layers(h))
            if n != len(mods):
                raise RuntimeError(f"vtts_b200: handle has {n} layers, module tree has {len(mods)}")
        sig = self._signature()
        if self._uploaded.get(idx) != sig:
            self._upload(h, dev)
            self._uploaded[idx] = sig
        return h

    def _upload(self, h: int, dev: torch.device) -> None:
        """Hand every layer's parameters to the library (weight-norm folded on the device)."""
        lib = _lib.load()
        stream = _lib.current_stream(dev)
        info = _lib.VttsLayerInfo()
        keep = []
        for i, m in enumerate(self._layer_modules()):
            _lib.check(lib.vtts_gen_layer_info(h, i, ctypes.byref(info)))
            if hasattr(m, "weight_g"):
                v, g = m.weight_v, m.weight_g
            else:
                v, g = m.weight, None
            want = (info.cout, info.cin, info.ksize) if info.kind == 0 else (info.cin, info.cout, info.ksize)
            if tuple(v.shape) != want:
                raise RuntimeError(f"vtts_b200: layer {i} weight shape {tuple(v.shape)} != {want}")
            v = v.detach().to(dev, torch.float32).contiguous()
            g = None if g is None else g.detach().to(dev, torch.float32).contiguous()
            b = None if m.bias is None else m.bias.detach().to(dev, torch.float32).contiguous()
            keep += [v, g, b]
            _lib.check(lib.vtts_gen_load_layer(h, i, _lib.ptr(v), _lib.ptr(g), _lib.ptr(b), stream))
        torch.cuda.current_stream(dev).synchronize()  # staging copies may be freed now
        del keep

    def _needs_autograd(self, *tensors) -> bool:
        if not torch.is_grad_enabled():
            return False
        if any(t is not None and t.requires_grad for t in tensors):
            return True
        return any(p.requires_grad for p in self.parameters())

    # Look-ahead of the V1 generator in mel frames, rounded up (input conv: 12-13 frames).  Accounting only (bench.py
    # counts lengths + this many frames as computed); the kernels derive their own per-layer margins from the config.
    TRIM_MARGIN_FRAMES = 16

    def forward_trimmed(self, c: torch.Tensor, lengths: torch.Tensor, g: Optional[torch.Tensor] = None,
                        margin_frames: Optional[int] = None) -> torch.Tensor:
        """Extension: like ``forward`` for a padded batch whose rows have ``lengths`` valid frames.

        Work that cannot reach the first ``lengths[b] * upsample_factor`` samples of row b is skipped: every layer
        computes ``lengths[b]`` frames plus the look-ahead its successors need (derived from the module's own kernel
        sizes, dilations and scales; ``margin_frames`` only adds extra frames on top).  Those samples are bit-identical
        to ``forward(c, g)``, later samples are zero or undefined.  Synthesis only (no autograd).
        """
        if self.precision == "fp32":
            return self._run_kernels(c, g)
        lens = lengths.detach().to(c.device, torch.int64).contiguous()
        if lens.numel() != c.shape[0]:
            raise ValueError("lengths must have one entry per batch row")
        m = -1 if margin_frames is None else int(margin_frames)
        return self._run_kernels(c, g, trim=(lens, m))

    def _run_kernels(self, c: torch.Tensor, g: Optional[torch.Tensor], dump_stage: int = -1, trim=None):
        lib = _lib.load()
        if not c.is_cuda:
            raise RuntimeError(
                f"vtts_b200.{type(self).__name__}: input is on {c.device}; the synthesis path only runs its CUDA "
                "kernels (no CPU fallback). Move the module and inputs to a B200."
            )
        if c.dim() != 3:
            raise ValueError(f"expected (B, in_channels, T), got {tuple(c.shape)}")
        cfg = self._gen_config()
        if c.shape[1] != cfg.in_channels:
            raise ValueError(f"expected {cfg.in_channels} input channels, got {c.shape[1]}")
        if self.precision not in _lib.PRECISION:
            raise ValueError(f"precision must be one of {list(_lib.PRECISION)}, got {self.precision!r}")
        prec = _lib.PRECISION[self.precision]
        dev = c.device
        B, _, T = c.shape
        with torch.cuda.device(dev):
            h = self._handle(dev)
            stream = _lib.current_stream(dev)
            x = c.detach().to(torch.float32).contiguous()
            gg = None
            if g is not None:
                if cfg.global_channels <= 0:
                    raise RuntimeError("g given but the module has no global conditioning conv")
                gg = g.detach().to(dev, torch.float32).reshape(B, cfg.global_channels).contiguous()
            need = ctypes.c_size_t()
            _lib.check(lib.vtts_gen_workspace_bytes(h, B, T, prec, ctypes.byref(need)))
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            ws = self._workspace.get(idx)
            if ws is None or ws.numel() < need.value:
                ws = None
                self._workspace.pop(idx, None)
                ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
                self._workspace[idx] = ws
            L = self._stage_shape(2 * cfg.num_upsamples, T)[1]
            wav = torch.empty((B, cfg.out_channels, L), dtype=torch.float32, device=dev)
            dump = None
            if dump_stage >= 0:
                ch, Ld = self._stage_shape(dump_stage, T)
                dump = torch.empty((B, ch, Ld), dtype=torch.float32, device=dev)
            if trim is not None:
                _lib.check(lib.vtts_gen_set_valid_lengths(h, trim[0].data_ptr(), trim[1]))
            try:
                _lib.check(lib.vtts_gen_forward(h, x.data_ptr(), _lib.ptr(gg), wav.data_ptr(), B, T, ws.data_ptr(),
                                                ws.numel(), prec, dump_stage, _lib.ptr(dump), stream))
            finally:
                if trim is not None:
                    lib.vtts_gen_set_valid_lengths(h, None, 0)
            self.last_launch_count = lib.vtts_gen_last_launch_count(h)
        if c.dtype != torch.float32:
            wav = wav.to(c.dtype)
        return (wav, dump) if dump_stage >= 0 else wav

    def _stage_shape(self, stage: int, T: int):
        cfg = self._gen_config()
        ch, L = cfg.channels, T
        if stage == 0:
            return ch, L
        for i in range(cfg.num_upsamples):
            s, k = cfg.upsample_scales[i], cfg.upsample_kernel_sizes[i]
            L = (L - 1) * s - 2 * cfg.upsample_paddings[i] + k + cfg.upsample_output_paddings[i]
            ch //= 2
            if stage in (2 * i + 1, 2 * i + 2):
                return ch, L
        raise ValueError(f"no stage {stage}")

    def graphed(self, c: torch.Tensor, g: Optional[torch.Tensor] = None,
                lengths: Optional[torch.Tensor] = None) -> "GraphedForward":
        """Extension for low-latency serving: capture the ~50 launches of one forward at this input shape into a CUDA
        graph.  ``c`` / ``g`` / ``lengths`` are examples that fix shapes and dtypes; see :class:`GraphedForward`."""
        return GraphedForward(self, c, g, lengths)

    FP16_MAX = 65504.0

    def activation_range(self, c: torch.Tensor, g: Optional[torch.Tensor] = None) -> dict:
        """Diagnostic for the 16-bit paths: run ``c`` through the fp32 kernels and report ``max |x|`` of every tensor the
        16-bit paths round to 16-bit operands (the input and every conv layer's output after its residual add).

        Returns ``{"layers": [max_abs per layer, library order], "input": max |c|, "max": overall max,
        "fp16_headroom": 65504 / max}``.  fp16 operands (``precision = "fp16"``, the default) clip silently at 65504
        (``cvt.rn.satfinite``): a checkpoint / input distribution with ``fp16_headroom`` comfortably above 1 cannot
        saturate; below 1 use ``precision = "bf16"`` or ``"fp32"``.  Synthesis-time check, not part of the hot path.
        """
        lib = _lib.load()
        dev = c.device
        with torch.cuda.device(dev):
            h = self._handle(dev)
            n = _lib.check(lib.vtts_gen_num_layers(h))
            probe = torch.zeros(n + 1, dtype=torch.float32, device=dev)
            saved = self.precision
            _lib.check(lib.vtts_gen_set_range_probe(h, probe.data_ptr()))
            try:
                self.precision = "fp32"
                self._run_kernels(c, g)
            finally:
                self.precision = saved
                lib.vtts_gen_set_range_probe(h, None)
            vals = probe.cpu().tolist()
        mx = max(vals)
        return {"layers": vals[:n], "input": vals[n], "max": mx, "fp16_headroom": self.FP16_MAX / mx if mx > 0 else float("inf")}

    def debug_stage(self, c: torch.Tensor, stage: int, g: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Test hook: intermediate tensor (B, C, L) fp32 of the kernel path (see vtts_gen_forward)."""
        return self._run_kernels(c, g, dump_stage=stage)[1]


class GraphedForward:
    """One generator forward at a fixed (B, T) captured into a CUDA graph (``module.graphed(example_c, ...)``).

    Replaying the graph removes the host cost of the ~50 kernel launches (tensor-map encoding, ctypes calls), which
    dominates single-utterance latency.  ``__call__`` copies the inputs into the captured buffers, replays and returns
    the captured output tensor -- valid until the next call.  The capture is redone automatically when a parameter of the
    module changes (same version signature that triggers the weight re-upload).  Synthesis only.
    """

    def __init__(self, module: "_GeneratorBase", c: torch.Tensor, g: Optional[torch.Tensor] = None,
                 lengths: Optional[torch.Tensor] = None):
        if not c.is_cuda:
            raise RuntimeError("vtts_b200.GraphedForward: inputs must be CUDA tensors (no CPU fallback)")
        self._m = module
        self._c = c.detach().clone()
        self._g = None if g is None else g.detach().clone()
        self._len = None if lengths is None else lengths.detach().to(c.device, torch.int64).clone()
        self._graph = None
        self._sig = None
        self._out = None

    def _run(self):
        if self._len is not None:
            return self._m.forward_trimmed(self._c, self._len, self._g)
        return self._m._run_kernels(self._c, self._g)

    def _capture(self):
        dev = self._c.device
        with torch.no_grad(), torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._run()                      # warm-up outside capture: weight upload, workspace allocation
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._out = self._run()
        self._graph, self._sig = graph, self._m._signature()

    @torch.no_grad()
    def __call__(self, c: torch.Tensor, g: Optional[torch.Tensor] = None,
                 lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        if tuple(c.shape) != tuple(self._c.shape):
            raise ValueError(f"GraphedForward was captured for {tuple(self._c.shape)}, got {tuple(c.shape)}")
        if (g is None) != (self._g is None) or (lengths is None) != (self._len is None):
            raise ValueError("GraphedForward: g / lengths must be given exactly as at capture time")
        self._c.copy_(c, non_blocking=True)
        if g is not None:
            self._g.copy_(g.reshape(self._g.shape), non_blocking=True)
        if lengths is not None:
            self._len.copy_(lengths, non_blocking=True)
        if self._graph is None or self._sig != self._m._signature():
            self._capture()
        self._graph.replay()
        return self._out


class HiFiGAN(_GeneratorBase):
    """HiFi-GAN generator (generator.py:16-213): input_conv, upsamples, blocks, output_conv."""

    def __init__(
        self,
        in_channels: int = 80,
        out_channels: int = 1,
        channels: int = 512,
        global_channels: int = -1,
        kernel_size: int = 7,
        upsample_scales: List[int] = [8, 8, 2, 2],
        upsample_kernel_sizes: List[int] = [16, 16, 4, 4],
        resblock_kernel_sizes: List[int] = [3, 7, 11],
        resblock_dilations: List[List[int]] = [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        use_additional_convs: bool = True,
        bias: bool = True,
        nonlinear_activation: str = "LeakyReLU",
        nonlinear_activation_params: Dict[str, Any] = {"negative_slope": 0.1},
        use_weight_norm: bool = True,
    ):
        super().__init__()
        assert kernel_size % 2 == 1, "Kernel size must be odd number."
        assert len(upsample_scales) == len(upsample_kernel_sizes)
        assert len(resblock_dilations) == len(resblock_kernel_sizes)
        if nonlinear_activation != "LeakyReLU":
            raise NotImplementedError("vtts_b200 kernels fuse LeakyReLU only (the reference never uses another)")
        self._cfg_args = dict(
            in_channels=in_channels, out_channels=out_channels, channels=channels, kernel_size=kernel_size,
            upsample_scales=list(upsample_scales), upsample_kernel_sizes=list(upsample_kernel_sizes),
            resblock_kernel_sizes=list(resblock_kernel_sizes),
            resblock_dilations=[list(d) for d in resblock_dilations],
            use_additional_convs=use_additional_convs, slope=float(nonlinear_activation_params.get("negative_slope", 0.01)),
        )
        self.upsample_factor = int(np.prod(upsample_scales) * out_channels)
        self.num_upsamples = len(upsample_kernel_sizes)
        self.num_blocks = len(resblock_kernel_sizes)
        self.global_channels = global_channels

        self.input_conv = nn.Conv1d(in_channels, channels, kernel_size, 1, padding=(kernel_size - 1) // 2)
        self.upsamples = nn.ModuleList()
        self.blocks = nn.ModuleList()
        width = channels
        for scale, ksize in zip(upsample_scales, upsample_kernel_sizes):
            assert ksize == 2 * scale
            self.upsamples.append(nn.Sequential(
                _act(nonlinear_activation, nonlinear_activation_params),
                nn.ConvTranspose1d(width, width // 2, ksize, scale, padding=scale // 2 + scale % 2,
                                   output_padding=scale % 2),
            ))
            width //= 2
            for rk, rd in zip(resblock_kernel_sizes, resblock_dilations):
                self.blocks.append(ResidualBlock(
                    kernel_size=rk, channels=width, dilations=rd, bias=bias,
                    use_additional_convs=use_additional_convs, nonlinear_activation=nonlinear_activation,
                    nonlinear_activation_params=nonlinear_activation_params,
                ))
        # the last activation deliberately uses nn.LeakyReLU()'s default slope 0.01 (generator.py:111)
        self.output_conv = nn.Sequential(
            nn.LeakyReLU(),
            nn.Conv1d(width, out_channels, kernel_size, 1, padding=(kernel_size - 1) // 2),
            nn.Tanh(),
        )
        if global_channels > 0:
            self.global_conv = nn.Conv1d(global_channels, channels, 1)
        if use_weight_norm:
            self.apply_weight_norm()
        self.reset_parameters()
        self._init_runtime()

    # -- reference API -----------------------------------------------------------------------
    def forward(self, c: torch.Tensor, g: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, in_channels, T) [, (B, global_channels, 1)] -> (B, out_channels, T * upsample_factor)."""
        if self._needs_autograd(c, g):
            if self.train_backend == "tc":       # generator backward on the kernels (training.py); "eager": PyTorch / cuDNN
                from .training import hifigan_forward_tc
                return hifigan_forward_tc(self, c, g)
            return self._forward_eager(c, g)
        return self._run_kernels(c, g)

    def inference(self, c: torch.Tensor, g: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(T, in_channels) -> (T * upsample_factor, out_channels)   [generator.py:197-213]."""
        if g is not None:
            g = g.unsqueeze(0)
        with torch.no_grad():
            y = self.forward(c.transpose(1, 0).unsqueeze(0), g=g)
        return y.squeeze(0).transpose(1, 0)

    def reset_parameters(self):
        """normal_(0, 0.01) on every conv ``weight`` (generator.py:158-171).

        Under weight norm ``weight`` is recomputed from weight_g / weight_v, so -- exactly as in
        the reference -- this does not change the effective initialisation.
        """
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
                m.weight.data.normal_(0.0, 0.01)
        if "_uploaded" in self.__dict__:
            self.invalidate()           # .data writes do not bump the version counter the upload cache keys on

    def remove_weight_norm(self):
        for m in self.modules():
            try:
                nn.utils.remove_weight_norm(m)
            except ValueError:
                pass
        self.__dict__.pop("_layer_cache", None)
        if "_uploaded" in self.__dict__:
            self.invalidate()

    def apply_weight_norm(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
                nn.utils.weight_norm(m)
        self.__dict__.pop("_layer_cache", None)
        if "_uploaded" in self.__dict__:
            self.invalidate()

    # -- plumbing ----------------------------------------------------------------------------
    def _forward_eager(self, c, g=None):
        c = self.input_conv(c)
        if g is not None:
            c = c + self.global_conv(g)
        for i in range(self.num_upsamples):
            c = self.upsamples[i](c)
            cs = 0.0
            for j in range(self.num_blocks):
                cs += self.blocks[i * self.num_blocks + j](c)
            c = cs / self.num_blocks
        return self.output_conv(c)

    def _layer_modules(self) -> List[nn.Module]:
        mods: List[nn.Module] = [self.input_conv]
        for i in range(self.num_upsamples):
            mods.append(self.upsamples[i][1])
            for j in range(self.num_blocks):
                blk = self.blocks[i * self.num_blocks + j]
                for m in range(len(blk.convs1)):
                    mods.append(blk.convs1[m][1])
                    if blk.use_additional_convs:
                        mods.append(blk.convs2[m][1])
        mods.append(self.output_conv[1])
        if self.global_channels > 0:
            mods.append(self.global_conv)
        return mods

    def _gen_config(self) -> _lib.VttsGenConfig:
        a = self._cfg_args
        cfg = _lib.VttsGenConfig()
        cfg.in_channels, cfg.out_channels, cfg.channels = a["in_channels"], a["out_channels"], a["channels"]
        cfg.global_channels = self.global_channels if self.global_channels > 0 else 0
        cfg.kernel_size = a["kernel_size"]
        cfg.num_upsamples = len(a["upsample_scales"])
        for i, (s, k) in enumerate(zip(a["upsample_scales"], a["upsample_kernel_sizes"])):
            cfg.upsample_scales[i], cfg.upsample_kernel_sizes[i] = s, k
            cfg.upsample_paddings[i], cfg.upsample_output_paddings[i] = s // 2 + s % 2, s % 2
        cfg.num_blocks = len(a["resblock_kernel_sizes"])
        for j, (k, dil) in enumerate(zip(a["resblock_kernel_sizes"], a["resblock_dilations"])):
            cfg.resblock_kernel_sizes[j] = k
            cfg.num_dilations[j] = len(dil)
            for m, d in enumerate(dil):
                cfg.resblock_dilations[j][m] = d
        cfg.use_additional_convs = 1 if a["use_additional_convs"] else 0
        cfg.lrelu_slope = a["slope"]
        cfg.final_lrelu_slope = 0.01
        return cfg

"""Fused AdamW for the training configuration (config/sod.yml:56-76): `optim_wrapper.optimizer = AdamW(lr=5e-4,
weight_decay=0.1)` with the per-prefix lr multipliers of `paramwise_cfg.custom_keys`, as ONE kernel launch per step
over flat parameter / gradient / moment buffers (csrc/optim_ops.cu).

The reference builds `torch.optim.AdamW` through mmengine's `DefaultOptimWrapperConstructor`; the update rule and
the custom-key matching (longest matching key wins; `lr_mult` / `decay_mult`) are restated here.  Parameters are
re-homed into one flat fp32 buffer (every `p.data` becomes a view, values preserved) laid out exactly like the flat
gradient buffer of `twig/graphs.py::GraphedTrainStep`, so a training step is: graph replay -> one all-reduce of the
flat gradients -> one optimizer launch.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import flat
from .ops import capi
from .ops.capi import call, ptr, stream

__all__ = ["FusedAdamW", "paramwise_options", "SOD_CUSTOM_KEYS"]

# config/sod.yml:62-76 (identical in config/cod.yml)
SOD_CUSTOM_KEYS: Dict[str, Dict[str, float]] = {
    "hitnet.backbone": {"lr_mult": 0.2},
    "hitnet.backbone.prompt_encoder.encoder2.downsample_layers": {"lr_mult": 0.02},
    "hitnet.backbone.prompt_encoder.encoder2.stages.0": {"lr_mult": 0.02},
    "hitnet.backbone.prompt_encoder.encoder2.stages.1": {"lr_mult": 0.02},
    "hitnet.backbone.prompt_encoder.encoder2.stages.2": {"lr_mult": 0.02},
    "hitnet.backbone.prompt_encoder.encoder2.stages.3": {"lr_mult": 0.02},
}
SLICE = 4096


def paramwise_options(name: str, lr: float, weight_decay: float,
                      custom_keys: Optional[Dict[str, Dict[str, float]]]) -> Tuple[float, float]:
    """(lr, weight_decay) of one parameter under mmengine's custom_keys rule: keys are tried longest first (ties
    alphabetically) and the first key that is a substring of the full parameter name applies."""
    if custom_keys:
        for key in sorted(sorted(custom_keys.keys()), key=len, reverse=True):
            if key in name:
                opt = custom_keys[key]
                return lr * opt.get("lr_mult", 1.0), weight_decay * opt.get("decay_mult", 1.0)
    return lr, weight_decay


class FusedAdamW:
    """opt = FusedAdamW(named_params, lr, weight_decay=..., custom_keys=...); opt.step(lr_scale=...).

    ORDER MATTERS with captured steps: the constructor re-homes every `p.data` into `self.flat_param`, so it must run
    BEFORE `GraphedTrainStep(..., flat_grad=opt.flat_grad)` / `GraphedPredict` capture -- a graph captured earlier
    would keep reading the old, freed parameter storage.  Both Graphed* classes record the parameter pointers at
    capture and raise if they changed; this constructor refuses parameters a live graph has captured."""

    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], lr: float = 5e-4,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.1,
                 custom_keys: Optional[Dict[str, Dict[str, float]]] = None,
                 flat_grad: Optional[torch.Tensor] = None):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        assert named, "no trainable parameters"
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW runs on CUDA parameters only (no CPU fallback)")
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        from . import graphs
        if graphs.captured_by_live_graph(self.params):
            raise RuntimeError("FusedAdamW: a CUDA graph has already captured these parameters; create the optimizer "
                               "first, then GraphedTrainStep(..., flat_grad=opt.flat_grad)")
        offs, total = flat.flat_offsets(self.params)
        self.offsets = offs
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        if flat_grad is None:
            flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
            flat.bind_views(self.params, flat_grad, "grad")
        else:   # the caller's flat gradient buffer must use the same layout (twig/flat.py)
            assert flat_grad.dtype == torch.float32 and flat_grad.is_cuda
            assert flat.views_match(self.params, flat_grad), ".grad views do not follow the twig/flat.py layout of flat_grad"
        self.flat_grad = flat_grad
        self.betas, self.eps, self.t = betas, eps, 0
        self.options: List[Tuple[float, float]] = []
        entries = []
        for p in self.params:
            assert p.dtype == torch.float32, "fp32 master parameters expected"
        flat.bind_views(self.params, self.flat_param, "data")
        for name, p, off in zip(self.names, self.params, offs):
            n = p.numel()
            plr, pwd = paramwise_options(name, lr, weight_decay, custom_keys)
            self.options.append((plr, pwd))
            for s in range(0, n, SLICE):
                entries.append(struct.pack("<qiffi", off + s, min(SLICE, n - s), plr, pwd, 0))
        assert capi.load().dgtd_adamw_slice_bytes() == struct.calcsize("<qiffi")
        self.nslices = len(entries)
        raw = torch.frombuffer(bytearray(b"".join(entries)), dtype=torch.uint8)
        self.table = raw.to(dev)

    def zero_grad(self) -> None:
        self.flat_grad.zero_()

    def step(self, grad_scale: float = 1.0, lr_scale: float = 1.0) -> None:
        """`lr_scale` multiplies every parameter's lr for this step: the factor of the run's scheduler
        (config/sod.yml `param_scheduler`: CosineAnnealingLR -> lr_t / lr_0), so a schedule needs no table rewrite."""
        self.t += 1
        call("dgtd_adamw_step", ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.table), self.nslices, float(self.betas[0]), float(self.betas[1]), float(self.eps), self.t,
             float(grad_scale), float(lr_scale), stream())
        # the kernel wrote the parameters behind autograd's back: bump the version counters so that cached
        # re-packed / down-cast shadows of the inference path (texture_diffuser._Packed) are refreshed
        torch.autograd.graph.increment_version(self.params)

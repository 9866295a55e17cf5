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
    """opt = FusedAdamW(named_params, lr, weight_decay=..., custom_keys=..., flat_grad=step.flat_grad); opt.step()."""

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
        total = sum(p.numel() for p in self.params)
        self.flat_param = torch.empty(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        if flat_grad is None:
            flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
            bind_grads = True
        else:
            assert flat_grad.numel() == total and flat_grad.dtype == torch.float32 and flat_grad.is_cuda
            bind_grads = False
        self.flat_grad = flat_grad
        self.betas, self.eps, self.t = betas, eps, 0
        self.options: List[Tuple[float, float]] = []
        entries = []
        off = 0
        with torch.no_grad():
            for name, p in zip(self.names, self.params):
                n = p.numel()
                assert p.dtype == torch.float32, "fp32 master parameters expected"
                self.flat_param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + n].view_as(p)
                if bind_grads:
                    p.grad = self.flat_grad[off:off + n].view_as(p)
                else:   # the caller's flat gradient buffer must use the same order and offsets
                    assert p.grad is not None and p.grad.data_ptr() == self.flat_grad.data_ptr() + 4 * off, \
                        f"{name}: .grad is not the view at offset {off} of flat_grad"
                plr, pwd = paramwise_options(name, lr, weight_decay, custom_keys)
                self.options.append((plr, pwd))
                for s in range(0, n, SLICE):
                    entries.append(struct.pack("<qiffi", off + s, min(SLICE, n - s), plr, pwd, 0))
                off += n
        assert capi.load().dgtd_adamw_slice_bytes() == struct.calcsize("<qiffi")
        self.nslices = len(entries)
        raw = torch.frombuffer(bytearray(b"".join(entries)), dtype=torch.uint8)
        self.table = raw.to(dev)

    def zero_grad(self) -> None:
        self.flat_grad.zero_()

    def step(self, grad_scale: float = 1.0) -> None:
        self.t += 1
        call("dgtd_adamw_step", ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.table), self.nslices, float(self.betas[0]), float(self.betas[1]), float(self.eps), self.t,
             float(grad_scale), stream())
        # the kernel wrote the parameters behind autograd's back: bump the version counters so that cached
        # re-packed / down-cast shadows of the inference path (texture_diffuser._Packed) are refreshed
        torch.autograd.graph.increment_version(self.params)

"""MAE, S-measure, F-measure and E-measure (twig/metric/{MAE,Smeasure,Fmeasure,Emeasure}.py:18-36: the four
evaluators of config/cod.yml:123-128) computed by csrc/metric_ops.cu.

The reference moves prediction and label to the host and loops over images in numpy (pysodmetrics 1.3.1); here
a batch is four kernel launches on the tensors `cod.forward(mode='predict')` returns, and 16 bytes per image come
back.  The wrappers keep the reference's protocol, including its quirk that every batch records the evaluator's
RUNNING mean over all images seen so far and `compute_metrics` averages those records (Smeasure.py:30-36).
mmengine's `BaseMetric` (collection across ranks) is runner plumbing and out of scope.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from ..ops import capi
from ..ops.capi import call, check_cuda, ptr, stream

__all__ = ["sod_metrics", "MAE", "Smeasure", "Fmeasure", "Emeasure"]


def sod_metrics(pred: torch.Tensor, gt: torch.Tensor, curves: bool = False):
    """pred, gt (B,1,H,W) fp32 in [0,1] on the GPU -> (B, 2) float64 = per-image (MAE, S-measure); with
    `curves=True` also (B, 2, 256) float64 = per-image changeable F-measure and E-measure curves."""
    check_cuda(pred, gt)
    assert pred.dtype == torch.float32 and gt.dtype == torch.float32, "fp32 maps expected (cod.py:217)"
    assert pred.shape == gt.shape and pred.dim() == 4 and pred.shape[1] == 1, (tuple(pred.shape), tuple(gt.shape))
    B, _, H, W = pred.shape
    ws = torch.empty(int(capi.load().dgtd_sod_metrics_ws_bytes(B, H, W)), device=pred.device, dtype=torch.uint8)
    out = torch.empty(B, 2, device=pred.device, dtype=torch.float64)
    cur = torch.empty(B, 2, 256, device=pred.device, dtype=torch.float64) if curves else None
    call("dgtd_sod_metrics_fwd", ptr(pred), ptr(gt), ptr(ws), ptr(out), ptr(cur), B, H, W, stream())
    return (out, cur) if curves else out


class _Running:
    default_prefix = 'COD'
    _column = 0
    _key = ''
    _name = ''

    def __init__(self, collect_device: str = 'cpu', prefix: Optional[str] = None, data_range: Optional[float] = 1.0):
        self.collect_device = collect_device
        self.prefix = prefix or self.default_prefix
        self.results: List[dict] = []
        self._sum = 0.0
        self._count = 0

    def process(self, data_batch, data_samples: Tuple[torch.Tensor, torch.Tensor]) -> None:
        pred, gt = data_samples
        vals = sod_metrics(pred.detach().float().contiguous(), gt.detach().float().contiguous())[:, self._column]
        for v in vals.cpu().tolist():           # the evaluator's list of per-image values, as a running sum
            self._sum += v
            self._count += 1
        self.results.append({self._key: self._sum / self._count})

    def compute_metrics(self, results: list) -> dict:
        return {self._name: sum(r[self._key] for r in results) / len(results)}

    def evaluate(self) -> dict:
        return self.compute_metrics(self.results)


class _RunningCurve:
    """Fmeasure.py / Emeasure.py: the evaluator keeps every image's 256-threshold curve; a batch records the
    maximum over thresholds of the MEAN curve over all images seen so far."""
    default_prefix = 'COD'
    _row = 0
    _key = ''
    _name = ''

    def __init__(self, collect_device: str = 'cpu', prefix: Optional[str] = None, data_range: Optional[float] = 1.0):
        self.collect_device = collect_device
        self.prefix = prefix or self.default_prefix
        self.results: List[dict] = []
        self._sum: Optional[torch.Tensor] = None      # running sum of curves, kept on the GPU
        self._count = 0

    def process(self, data_batch, data_samples: Tuple[torch.Tensor, torch.Tensor]) -> None:
        pred, gt = data_samples
        _, cur = sod_metrics(pred.detach().float().contiguous(), gt.detach().float().contiguous(), curves=True)
        batch_sum = cur[:, self._row].sum(0)
        self._sum = batch_sum if self._sum is None else self._sum + batch_sum
        self._count += cur.shape[0]
        self.results.append({self._key: float((self._sum / self._count).max())})

    def compute_metrics(self, results: list) -> dict:
        return {self._name: sum(r[self._key] for r in results) / len(results)}

    def evaluate(self) -> dict:
        return self.compute_metrics(self.results)


class Fmeasure(_RunningCurve):
    """twig/metric/Fmeasure.py."""
    _row, _key, _name = 0, 'fm', 'Fmeasure'


class Emeasure(_RunningCurve):
    """twig/metric/Emeasure.py."""
    _row, _key, _name = 1, 'em', 'Emeasure'


class MAE(_Running):
    """twig/metric/MAE.py."""
    _column, _key, _name = 0, 'mae', 'MAE'


class Smeasure(_Running):
    """twig/metric/Smeasure.py."""
    _column, _key, _name = 1, 'sm', 'Smeasure'

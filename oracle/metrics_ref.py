"""CPU restatement of the evaluation metrics of the reference's test path: MAE and S-measure.

TEST INFRASTRUCTURE ONLY: imported by tests/ and bench.py's CPU legs; never by the product path.

The arithmetic lives in a third-party dependency that is NOT under /root/reference and NOT installed here:
**pysodmetrics 1.3.1** (`requirements.txt:110`, imported as `py_sod_metrics`).  This file restates its published
algorithm (Fan et al., "Structure-measure", ICCV 2017, alpha = 0.5; MAE = mean |pred - gt|) in plain numpy
float64 and anchors on the reference's own call sites:

  twig/metric/Smeasure.py:18-36   pred, gt (B,1,H,W) in [0,1] -> `(x * 255).astype(np.uint8)` (truncation) ->
                                  evaluator.step per image -> the batch records the evaluator's RUNNING mean
                                  over all images seen so far; compute_metrics = mean of those records
  twig/metric/MAE.py:18-36        the same wrapper around the MAE evaluator
  twig/metric/Fmeasure.py:18-36   `get_results()["fm"]["curve"].max()`: max over the 256 thresholds of the mean
  twig/metric/Emeasure.py:18-36   changeable F / E curve over all images seen so far, recorded per batch

**Parity unpinned**: the reference holds no golden vectors for the metrics and the library cannot be run here;
tests/test_oracle_metrics.py pins this restatement on hand-computable cases only (empty / full ground truth,
perfect and inverted predictions, a 2x2 case worked out by hand).

py_sod_metrics semantics followed (`_prepare_data`, `MAE.step`, `Smeasure.cal_sm/object/s_object/region/
centroid/divide_with_xy/ssim`): gt = gt > 128; pred = pred / 255, min-max normalised unless constant;
centroid = round-half-even of the mean foreground index, + 1; unbiased variances (N - 1); eps = np.spacing(1).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

EPS = np.spacing(1)


def prepare(pred_u8: np.ndarray, gt_u8: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    gt = gt_u8 > 128
    pred = pred_u8.astype(np.float64) / 255.0
    if pred.max() != pred.min():
        pred = (pred - pred.min()) / (pred.max() - pred.min())
    return pred, gt


def mae_one(pred_u8: np.ndarray, gt_u8: np.ndarray) -> float:
    pred, gt = prepare(pred_u8, gt_u8)
    return float(np.mean(np.abs(pred - gt)))


def _s_object(values: np.ndarray) -> float:
    x = np.mean(values)
    sigma = np.std(values, ddof=1)
    return float(2 * x / (x * x + 1 + sigma + EPS))


def _object(pred: np.ndarray, gt: np.ndarray) -> float:
    u = np.mean(gt)
    return float(u * _s_object(pred[gt]) + (1 - u) * _s_object((1 - pred)[~gt]))


def _centroid(gt: np.ndarray) -> Tuple[int, int]:
    h, w = gt.shape
    if np.count_nonzero(gt) == 0:
        x, y = np.round(w / 2), np.round(h / 2)
    else:
        y, x = np.argwhere(gt).mean(axis=0).round()
    return int(x) + 1, int(y) + 1


def _ssim(pred: np.ndarray, gt: np.ndarray) -> float:
    n = pred.size
    x, y = np.mean(pred), np.mean(gt)
    sx = np.sum((pred - x) ** 2) / (n - 1)
    sy = np.sum((gt - y) ** 2) / (n - 1)
    sxy = np.sum((pred - x) * (gt - y)) / (n - 1)
    alpha = 4 * x * y * sxy
    beta = (x * x + y * y) * (sx + sy)
    if alpha != 0:
        return float(alpha / (beta + EPS))
    return 1.0 if beta == 0 else 0.0


def _region(pred: np.ndarray, gt: np.ndarray) -> float:
    h, w = gt.shape
    x, y = _centroid(gt)
    area = h * w
    g = gt.astype(np.float64)
    w1 = x * y / area
    w2 = y * (w - x) / area
    w3 = (h - y) * x / area
    w4 = 1 - w1 - w2 - w3
    return float(w1 * _ssim(pred[:y, :x], g[:y, :x]) + w2 * _ssim(pred[:y, x:], g[:y, x:]) +
                 w3 * _ssim(pred[y:, :x], g[y:, :x]) + w4 * _ssim(pred[y:, x:], g[y:, x:]))


def smeasure_one(pred_u8: np.ndarray, gt_u8: np.ndarray, alpha: float = 0.5) -> float:
    pred, gt = prepare(pred_u8, gt_u8)
    y = np.mean(gt)
    if y == 0:
        return float(1 - np.mean(pred))
    if y == 1:
        return float(np.mean(pred))
    return max(0.0, alpha * _object(pred, gt) + (1 - alpha) * _region(pred, gt))


BETA = 0.3


def _cum_hists(pred: np.ndarray, gt: np.ndarray):
    """`cal_pr` / `cal_em_with_cumsumhistogram`: the normalised prediction is re-quantised to uint8 and the
    foreground / background counts at the 256 thresholds are flipped cumulative histograms
    (entry i = number of pixels with value >= 255 - i)."""
    q = (pred * 255).astype(np.uint8)
    bins = np.linspace(0, 256, 257)
    fg_hist, _ = np.histogram(q[gt], bins=bins)
    bg_hist, _ = np.histogram(q[~gt], bins=bins)
    return np.cumsum(np.flip(fg_hist)), np.cumsum(np.flip(bg_hist))


def fmeasure_curve_one(pred_u8: np.ndarray, gt_u8: np.ndarray, beta: float = BETA) -> np.ndarray:
    """pysodmetrics `Fmeasure.cal_pr` -> changeable F-measure at the 256 thresholds (beta^2 = 0.3)."""
    pred, gt = prepare(pred_u8, gt_u8)
    fg_w, bg_w = _cum_hists(pred, gt)
    tps = fg_w
    ps = fg_w + bg_w
    ps = np.where(ps == 0, 1, ps)
    t = max(np.count_nonzero(gt), 1)
    precisions = tps / ps
    recalls = tps / t
    numerator = (1 + beta) * precisions * recalls
    denominator = np.where(numerator == 0, 1, beta * precisions + recalls)
    return numerator / denominator


def emeasure_curve_one(pred_u8: np.ndarray, gt_u8: np.ndarray) -> np.ndarray:
    """pysodmetrics `Emeasure.cal_em_with_cumsumhistogram`: enhanced-alignment measure at the 256 thresholds."""
    pred, gt = prepare(pred_u8, gt_u8)
    size = gt.shape[0] * gt.shape[1]
    gt_fg = np.count_nonzero(gt)
    fg_fg, fg_bg = _cum_hists(pred, gt)
    pred_fg = fg_fg + fg_bg
    pred_bg = size - pred_fg
    if gt_fg == 0:
        total = pred_bg.astype(np.float64)
    elif gt_fg == size:
        total = pred_fg.astype(np.float64)
    else:
        bg_fg = gt_fg - fg_fg
        bg_bg = pred_bg - bg_fg
        mean_pred = pred_fg / size
        mean_gt = gt_fg / size
        combos = [(1 - mean_pred, 1 - mean_gt), (1 - mean_pred, 0 - mean_gt),
                  (0 - mean_pred, 1 - mean_gt), (0 - mean_pred, 0 - mean_gt)]
        total = np.zeros(256, np.float64)
        parts = np.empty((4, 256), np.float64)
        for i, (numel, (a, b)) in enumerate(zip([fg_fg, fg_bg, bg_fg, bg_bg], combos)):
            align = 2 * (a * b) / (a ** 2 + b ** 2 + EPS)
            parts[i] = (align + 1) ** 2 / 4 * numel
        total = parts.sum(axis=0)
    return total / (size - 1 + EPS)


class RunningCurveMetric:
    """Fmeasure.py:18-36 / Emeasure.py:18-36: per batch record max over thresholds of the MEAN curve over all
    images seen so far (`get_results()[..]["curve"].max()`); compute_metrics = mean of the records."""

    def __init__(self, fn):
        self.fn = fn
        self.curves: List[np.ndarray] = []
        self.results: List[float] = []

    def process(self, pred, gt) -> None:
        for p, g in zip(quantise(pred), quantise(gt)):
            self.curves.append(self.fn(p, g))
        self.results.append(float(np.mean(np.array(self.curves, dtype=np.float64), axis=0).max()))

    def compute_metrics(self) -> float:
        return sum(self.results) / len(self.results)


def quantise(t) -> np.ndarray:
    """Smeasure.py:25-26: `(x * 255).astype(np.uint8)` on the float32 (B,1,H,W) batch, channel squeezed."""
    a = np.asarray(t, dtype=np.float32)
    return (a[:, 0] * np.float32(255)).astype(np.uint8)


class RunningMetric:
    """The reference wrappers (Smeasure.py:18-36 / MAE.py:18-36): per batch append the running mean."""

    def __init__(self, fn):
        self.fn = fn
        self.values: List[float] = []
        self.results: List[float] = []

    def process(self, pred, gt) -> None:
        for p, g in zip(quantise(pred), quantise(gt)):
            self.values.append(self.fn(p, g))
        self.results.append(float(np.mean(np.array(self.values, dtype=np.float64))))

    def compute_metrics(self) -> float:
        return sum(self.results) / len(self.results)

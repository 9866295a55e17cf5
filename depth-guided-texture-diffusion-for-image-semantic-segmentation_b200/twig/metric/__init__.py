"""Evaluation metrics of the reference's test path on the GPU: mirrors of twig/metric/MAE.py
and twig/metric/Smeasure.py (same class names, `process` / `compute_metrics` protocol and result keys)."""
from .sod import MAE, Emeasure, Fmeasure, Smeasure, sod_metrics  # noqa: F401

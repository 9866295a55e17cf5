"""MAE / S-measure kernels (csrc/metric_ops.cu) against the numpy restatement of pysodmetrics
(oracle/metrics_ref.py), and the reference wrappers' protocol (twig/metric/*.py)."""
import warnings

import numpy as np
import pytest
import torch

import common
from oracle import metrics_ref as M

pytestmark = pytest.mark.gpu


def _blobs(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.zeros(B, 1, H, W)
    for b in range(B):
        for _ in range(2):
            y0, x0 = int(torch.randint(0, H // 2, (1,), generator=g)), int(torch.randint(0, W // 2, (1,), generator=g))
            hh, ww = int(torch.randint(2, H // 2, (1,), generator=g)), int(torch.randint(2, W // 2, (1,), generator=g))
            gt[b, 0, y0:y0 + hh, x0:x0 + ww] = 1.0
    pred = torch.sigmoid(4 * (gt - 0.5) + 2.0 * torch.randn(B, 1, H, W, generator=g))
    return pred, gt


def _oracle(pred, gt):
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for p, g in zip(M.quantise(pred.numpy()), M.quantise(gt.numpy())):
            out.append((M.mae_one(p, g), M.smeasure_one(p, g)))
    return np.array(out)


@pytest.mark.parametrize("shape", [(3, 64, 48), (2, 384, 384), (5, 37, 53), (1, 9, 1000)])
def test_metrics_match_the_oracle(shape):
    common.package()
    from dgtd_b200.twig.metric import sod_metrics
    B, H, W = shape
    pred, gt = _blobs(B, H, W, seed=H + W)
    got = sod_metrics(pred.cuda(), gt.cuda()).cpu().numpy()
    ref = _oracle(pred, gt)
    assert np.abs(got - ref).max() <= 1e-12, (got, ref)


def test_edge_cases():
    """Empty / full ground truth, constant prediction, perfect and inverted prediction, a foreground confined to the
    last column (degenerate quadrant: the library's NaN -> 0), soft labels around the 128 cut."""
    common.package()
    from dgtd_b200.twig.metric import sod_metrics
    H, W = 24, 40
    g = torch.Generator().manual_seed(0)
    rnd = torch.rand(1, 1, H, W, generator=g)
    blob = torch.zeros(1, 1, H, W)
    blob[..., 6:18, 10:30] = 1.0
    last_col = torch.zeros(1, 1, H, W)
    last_col[..., :, W - 1] = 1.0
    soft = torch.rand(1, 1, H, W, generator=g)
    cases = [(rnd, torch.zeros(1, 1, H, W)), (rnd, torch.ones(1, 1, H, W)), (torch.full((1, 1, H, W), 0.2), blob),
             (blob.clone(), blob), (1 - blob, blob), (rnd, last_col), (rnd, soft)]
    pred = torch.cat([c[0] for c in cases])
    gt = torch.cat([c[1] for c in cases])
    got = sod_metrics(pred.cuda(), gt.cuda()).cpu().numpy()
    ref = _oracle(pred, gt)
    assert np.abs(got - ref).max() <= 1e-12, (got, ref)
    assert got[3, 0] == 0.0 and abs(got[3, 1] - 1.0) < 1e-12 and got[4, 0] == 1.0 and got[5, 1] == 0.0


def test_wrappers_follow_the_reference_protocol():
    common.package()
    from dgtd_b200.twig.metric import MAE, Smeasure
    mae, sm = MAE(), Smeasure()
    omae, osm = M.RunningMetric(M.mae_one), M.RunningMetric(M.smeasure_one)
    for seed in range(3):
        pred, gt = _blobs(2, 32, 32, seed)
        for m in (mae, sm):
            m.process(None, (pred.cuda(), gt.cuda()))
        omae.process(pred.numpy(), gt.numpy())
        osm.process(pred.numpy(), gt.numpy())
    assert abs(mae.evaluate()["MAE"] - omae.compute_metrics()) < 1e-12
    assert abs(sm.evaluate()["Smeasure"] - osm.compute_metrics()) < 1e-12
    assert [set(r) for r in mae.results] == [{"mae"}] * 3 and [set(r) for r in sm.results] == [{"sm"}] * 3


def test_bit_stable_across_batch_composition():
    """Integer moments: an image's metrics do not depend on which batch it is evaluated in."""
    common.package()
    from dgtd_b200.twig.metric import sod_metrics
    pred, gt = _blobs(4, 96, 80, seed=3)
    a = sod_metrics(pred.cuda(), gt.cuda()).cpu()
    b = torch.cat([sod_metrics(pred[i:i + 1].cuda(), gt[i:i + 1].cuda()).cpu() for i in range(4)])
    assert torch.equal(a, b)


def _oracle_curves(pred, gt):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.array([[M.fmeasure_curve_one(p, g), M.emeasure_curve_one(p, g)]
                         for p, g in zip(M.quantise(pred.numpy()), M.quantise(gt.numpy()))])


@pytest.mark.parametrize("shape", [(3, 64, 48), (2, 384, 384), (4, 37, 53)])
def test_f_and_e_measure_curves_match_the_oracle(shape):
    common.package()
    from dgtd_b200.twig.metric import sod_metrics
    B, H, W = shape
    pred, gt = _blobs(B, H, W, seed=H * W)
    vals, cur = sod_metrics(pred.cuda(), gt.cuda(), curves=True)
    ref = _oracle_curves(pred, gt)
    assert cur.shape == (B, 2, 256)
    assert np.abs(cur.cpu().numpy() - ref).max() <= 1e-12
    assert np.abs(vals.cpu().numpy() - _oracle(pred, gt)).max() <= 1e-12


def test_curve_edge_cases_and_requantisation():
    """Empty / full ground truth, constant prediction, perfect / inverted prediction, and narrow prediction ranges
    (where the float64 re-quantisation `(p * 255).astype(uint8)` lands on or next to integers)."""
    common.package()
    from dgtd_b200.twig.metric import sod_metrics
    H, W = 24, 40
    g = torch.Generator().manual_seed(0)
    rnd = torch.rand(1, 1, H, W, generator=g)
    blob = torch.zeros(1, 1, H, W)
    blob[..., 6:18, 10:30] = 1.0
    cases = [(rnd, torch.zeros(1, 1, H, W)), (rnd, torch.ones(1, 1, H, W)), (torch.full((1, 1, H, W), 0.2), blob),
             (blob.clone(), blob), (1 - blob, blob)]
    for lo, hi in ((0.1, 0.9), (0.3, 0.31), (0.0, 0.4), (0.52, 1.0), (0.2, 0.2 + 3 / 255), (0.11, 0.77)):
        cases.append((lo + (hi - lo) * torch.rand(1, 1, H, W, generator=g), blob))
    pred = torch.cat([c[0] for c in cases])
    gt = torch.cat([c[1] for c in cases])
    _, cur = sod_metrics(pred.cuda(), gt.cuda(), curves=True)
    assert np.abs(cur.cpu().numpy() - _oracle_curves(pred, gt)).max() <= 1e-12


def test_curve_wrappers_follow_the_reference_protocol():
    common.package()
    from dgtd_b200.twig.metric import Emeasure, Fmeasure
    fm, em = Fmeasure(), Emeasure()
    ofm, oem = M.RunningCurveMetric(M.fmeasure_curve_one), M.RunningCurveMetric(M.emeasure_curve_one)
    for seed in range(3):
        pred, gt = _blobs(2, 32, 32, seed)
        for m in (fm, em):
            m.process(None, (pred.cuda(), gt.cuda()))
        ofm.process(pred.numpy(), gt.numpy())
        oem.process(pred.numpy(), gt.numpy())
    assert abs(fm.evaluate()["Fmeasure"] - ofm.compute_metrics()) < 1e-12
    assert abs(em.evaluate()["Emeasure"] - oem.compute_metrics()) < 1e-12

"""the numpy restatement of pysodmetrics' MAE / S-measure (oracle/metrics_ref.py) on
hand-computable cases.  The library is not installed and the reference holds no golden vectors for the metrics
(parity unpinned); these cases pin the published definitions."""
import warnings

import numpy as np

from oracle import metrics_ref as M


def test_perfect_prediction():
    gt = np.zeros((16, 20), np.uint8)
    gt[4:12, 5:15] = 255
    assert M.mae_one(gt.copy(), gt) == 0.0
    assert abs(M.smeasure_one(gt.copy(), gt) - 1.0) < 1e-12


def test_inverted_prediction():
    gt = np.zeros((16, 20), np.uint8)
    gt[4:12, 5:15] = 255
    assert M.mae_one(255 - gt, gt) == 1.0
    # object term: fg values are all 0 and bg values all 0 -> 0; region term: anti-correlated quadrants -> negative,
    # the sum is clamped at 0
    assert M.smeasure_one(255 - gt, gt) == 0.0


def test_empty_and_full_ground_truth():
    rng = np.random.default_rng(0)
    pred = rng.integers(0, 256, (12, 12)).astype(np.uint8)
    norm = (pred / 255.0 - pred.min() / 255.0) / (pred.max() / 255.0 - pred.min() / 255.0)
    assert abs(M.smeasure_one(pred, np.zeros_like(pred)) - (1 - norm.mean())) < 1e-12
    assert abs(M.smeasure_one(pred, np.full_like(pred, 255)) - norm.mean()) < 1e-12
    assert abs(M.mae_one(pred, np.zeros_like(pred)) - norm.mean()) < 1e-12


def test_constant_prediction_is_not_normalised():
    gt = np.zeros((8, 8), np.uint8)
    gt[:, :4] = 255
    pred = np.full((8, 8), 51, np.uint8)                  # 0.2 everywhere
    assert abs(M.mae_one(pred, gt) - (0.5 * 0.2 + 0.5 * 0.8)) < 1e-12


def test_hand_worked_two_by_two():
    """gt = [[1,0],[0,0]], pred (after min-max) = [[1, .5],[0, 0]] -> centroid (x, y) = (1, 1); every quadrant is a
    single pixel, unbiased variances are 0/0: the library's NaN propagates and `max(0, nan)` yields 0."""
    gt = np.array([[255, 0], [0, 0]], np.uint8)
    pred = np.array([[255, 128], [1, 1]], np.uint8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert M.smeasure_one(pred, gt) == 0.0
    p = (pred / 255.0 - 1 / 255.0) / (254 / 255.0)
    assert abs(M.mae_one(pred, gt) - (abs(p[0, 0] - 1) + p[0, 1] + p[1, 0] + p[1, 1]) / 4) < 1e-12


def test_object_and_region_terms_by_formula():
    """4x4 case small enough to evaluate the published formulas term by term."""
    gt = np.zeros((4, 4), np.uint8)
    gt[1:3, 1:3] = 255
    pred = np.array([[0, 0, 0, 0], [0, 255, 204, 0], [0, 204, 255, 51], [0, 0, 0, 0]], np.uint8)
    p = pred / 255.0
    g = gt > 128
    fg, bg = p[g], (1 - p)[~g]
    so = lambda v: 2 * v.mean() / (v.mean() ** 2 + 1 + v.std(ddof=1) + np.spacing(1))  # noqa: E731
    obj = g.mean() * so(fg) + (1 - g.mean()) * so(bg)
    # centroid of the 2x2 block: rows {1,2} -> mean 1.5 -> round-half-even 2 -> y = 3; same for x
    x = y = 3

    def ssim(a, b):
        n = a.size
        ma, mb = a.mean(), b.mean()
        sa, sb = ((a - ma) ** 2).sum() / (n - 1), ((b - mb) ** 2).sum() / (n - 1)
        sab = ((a - ma) * (b - mb)).sum() / (n - 1)
        al, be = 4 * ma * mb * sab, (ma ** 2 + mb ** 2) * (sa + sb)
        return al / (be + np.spacing(1)) if al != 0 else (1.0 if be == 0 else 0.0)
    gf = g.astype(float)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        reg = (9 / 16) * ssim(p[:y, :x], gf[:y, :x]) + (3 / 16) * ssim(p[:y, x:], gf[:y, x:]) + \
              (3 / 16) * ssim(p[y:, :x], gf[y:, :x]) + (1 - 15 / 16) * ssim(p[y:, x:], gf[y:, x:])
        want = max(0.0, 0.5 * obj + 0.5 * reg)
        got = M.smeasure_one(pred, gt)
    assert (np.isnan(want) and got == 0.0) or abs(got - want) < 1e-12


def test_running_mean_quirk_of_the_wrappers():
    rng = np.random.default_rng(1)
    m = M.RunningMetric(M.mae_one)
    batches = [(rng.random((2, 1, 8, 8), dtype=np.float32), (rng.random((2, 1, 8, 8)) > 0.5).astype(np.float32))
               for _ in range(3)]
    per_image = []
    for p, g in batches:
        m.process(p, g)
        per_image += [M.mae_one(a, b) for a, b in zip(M.quantise(p), M.quantise(g))]
    running = [np.mean(per_image[:2]), np.mean(per_image[:4]), np.mean(per_image[:6])]
    assert abs(m.compute_metrics() - np.mean(running)) < 1e-15


def test_fmeasure_and_emeasure_curves_on_hand_cases():
    gt = np.zeros((10, 10), np.uint8)
    gt[2:6, 3:8] = 255                                       # 20 foreground pixels
    f = M.fmeasure_curve_one(gt.copy(), gt)
    e = M.emeasure_curve_one(gt.copy(), gt)
    # perfect binary prediction: every threshold 1..255 reproduces gt -> precision = recall = 1 -> F = 1; the
    # threshold 0 (last entry: value >= 0) marks everything foreground -> precision 0.2
    assert np.allclose(f[:255], 1.0) and abs(f[255] - 1.3 * 0.2 / (0.3 * 0.2 + 1.0)) < 1e-12
    # E-measure of a perfectly aligned map: every pixel scores 1 -> size / (size - 1)
    assert np.allclose(e[:255], 100 / 99.0)
    # inverted prediction: no true positives at thresholds > 0 -> F = 0
    fi = M.fmeasure_curve_one(255 - gt, gt)
    assert np.allclose(fi[:255], 0.0)
    # empty ground truth: E counts the predicted background
    rng = np.random.default_rng(2)
    pred = rng.integers(0, 256, (10, 10)).astype(np.uint8)
    e0 = M.emeasure_curve_one(pred, np.zeros_like(pred))
    p, _ = M.prepare(pred, np.zeros_like(pred))
    q = (p * 255).astype(np.uint8)
    want = np.array([(q < 255 - i).sum() for i in range(256)]) / (100 - 1 + M.EPS)
    assert np.allclose(e0, want)
    # a threshold-by-threshold brute force of both curves on random data
    gtr = (rng.random((12, 9)) > 0.6).astype(np.uint8) * 255
    pr = rng.integers(0, 256, (12, 9)).astype(np.uint8)
    p, g = M.prepare(pr, gtr)
    q = (p * 255).astype(np.uint8)
    fc, ec = M.fmeasure_curve_one(pr, gtr), M.emeasure_curve_one(pr, gtr)
    for i in (0, 17, 128, 200, 255):
        b = q >= 255 - i
        tp = (b & g).sum()
        prec, rec = tp / max(b.sum(), 1), tp / max(g.sum(), 1)
        num = 1.3 * prec * rec
        assert abs(fc[i] - (num / (0.3 * prec + rec) if num else 0.0)) < 1e-12
        a, c = b - b.mean(), g - g.mean()
        align = 2 * a * c / (a * a + c * c + M.EPS)
        assert abs(ec[i] - ((align + 1) ** 2 / 4).sum() / (g.size - 1 + M.EPS)) < 1e-9

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

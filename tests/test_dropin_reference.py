"""True drop-in test (VERDICT r1 "missing" 6): the reference's OWN `PyramidVisionTransformerImpr.forward_features`
(cod.py:1455-1509) is executed UNMODIFIED -- reference patch embeds, PVT blocks, norms, the interpolate / flatten /
permute injection lines, all in eager PyTorch -- with this repo's hot-path classes substituted for the names that
`PyramidVisionTransformerImpr.__init__` resolves (cod.py:1394-1396), exactly the edit INTEGRATION.md section 2(a)
describes.  The result must equal the golden fixture recorded from the pure reference (tests/golden/pvt_128.npz).

The reference file is the staged copy under oracle/_ref (oracle/make_ref.py); skipped when it is absent."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import ref_loader as R

HOT_PATH_CLASSES = ("LayerNorm", "ShapePropWeightRegressor", "convnext_Block", "ShapePropEncoder", "MessagePassing",
                    "ShapePropDecoder", "prompt_encoder", "prompt_decoder")


def _patched_backbone(m, TD):
    saved = {n: getattr(m, n) for n in HOT_PATH_CLASSES}
    try:
        for n in HOT_PATH_CLASSES:
            setattr(m, n, getattr(TD, n))
        torch.manual_seed(0)
        net = m.pvt_v2_b2()
    finally:
        for n, v in saved.items():
            setattr(m, n, v)
    return net


@pytest.mark.skipif(not R.reference_available(), reason="reference file not staged (oracle/make_ref.py)")
def test_patched_reference_keeps_the_reference_state_dict():
    """CPU part: construction through the reference's own __init__ gives the reference's key set / shapes."""
    m = R.load_reference()
    TD = common.package()
    net = _patched_backbone(m, TD)
    assert type(net.prompt_encoder) is TD.prompt_encoder and type(net.prompt_decoder[0]) is TD.prompt_decoder
    torch.manual_seed(0)
    ref = m.pvt_v2_b2()
    a, b = net.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    # same seed -> same values: the mirror constructors consume the RNG exactly like the reference's
    assert all(torch.equal(a[k], b[k]) for k in a)
    net.load_state_dict(b)                                   # a reference checkpoint loads unchanged


@pytest.mark.gpu
@pytest.mark.skipif(not R.reference_available(), reason="reference file not staged (oracle/make_ref.py)")
def test_reference_forward_features_runs_on_the_repo_classes(monkeypatch):
    m = R.load_reference()
    TD = common.package()
    g = np.load(os.path.join(common.GOLDEN, "pvt_128.npz"))
    S, B = int(g["S"]), int(g["B"])
    # the reference's own eager convolutions (patch embeds, spatial-reduction convs) would otherwise run as TF32 in
    # cuDNN (torch default `cudnn.allow_tf32 = True`: 1.8e-4 on stage 1, measured) -- the fixture is a float64 run
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    net = _patched_backbone(m, TD).eval()
    common.fill_params_(net, seed=0)
    net = net.cuda()
    TD.set_precision(net, "fp32")
    image, depth = common.synthetic_inputs(B, S, seed=7)
    from dgtd_b200.twig.ops import capi
    n0 = capi.launch_count()
    with torch.no_grad():
        e1, outs = net.forward_features(image.cuda(), depth.cuda())      # the reference's code, line for line
    assert capi.launch_count() > n0, "the hot path did not run on libdgtd_ops.so"
    worst = 0.0
    for s, o in enumerate(outs):
        ref = torch.from_numpy(g[f"out{s}"]).double()
        got = o[:, ::4, ::2, ::2].double().cpu()
        err = float((got - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        assert err <= 1e-4, (s, err)
        assert np.allclose(common.moments(o.cpu()), g[f"out{s}_moments"], rtol=1e-4, atol=1e-7)
    print(f"drop-in: reference forward_features on the repo's hot-path classes, worst rel err vs pure reference {worst:.3e}")
    # and the pure reference in eager fp32 on the same GPU (what the fixture was recorded from, there in float64)
    torch.manual_seed(0)
    pure = m.pvt_v2_b2().eval()
    pure.prompt_encoder.message_passing.img_size = S
    common.fill_params_(pure, seed=0)
    pure = pure.cuda()
    with torch.no_grad():
        _, pouts = pure.forward_features(image.cuda(), depth.cuda())
    for a, b in zip(outs, pouts):
        assert common.rel_err(a, b) <= 1e-4

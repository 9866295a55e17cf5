"""SURVEY.md 8f-2 in training: the Hitnet decoder under autograd (cod.py:355-506, 685-807; train-mode BatchNorm,
cod.py:362) and the whole `cod.forward(mode='loss')` step (cod.py:118-146) on the CUDA path vs float64 autograd of the
CPU oracle (oracle/hitnet_ref.py with train=True, oracle/loss_ref.py) and vs the fixture generated from the UNMODIFIED
reference in train() (tests/golden/make_golden_hitnet_train.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import common
from oracle import hitnet_ref as H
from oracle import loss_ref as L

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel2(a, b):
    """L2-relative error: the metric of the bf16-operand checks (a max-norm over a small tensor is all rounding noise)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2)


@pytest.fixture(scope="module")
def HT():
    common.package()
    from dgtd_b200.twig.ops.functions import hitnet_train_func
    return hitnet_train_func


# ---- primitives -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,C", [(234, 32), (1, 32), (5000, 96), (513, 8)])
def test_batchnorm_train_forward_backward(HT, M, C):
    """Batch statistics, running update (momentum 0.1, unbiased variance) and the three gradients vs nn.BatchNorm in
    float64; ragged row counts (one partial CTA, several CTAs, a single row)."""
    g = torch.Generator().manual_seed(3)
    y = torch.randn(M, C, generator=g) * 2.0 + 3.0          # |mean| > sigma: the cancellation-prone case
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    go = torch.randn(M, C, generator=g)
    bn = torch.nn.BatchNorm1d(C).double().train()
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
        bn.running_mean.normal_(generator=g)
        bn.running_var.uniform_(0.5, 2.0, generator=g)
    rm, rv = bn.running_mean.clone().float().cuda(), bn.running_var.clone().float().cuda()
    yd = y.double().requires_grad_(True)
    if M > 1:
        want = bn(yd)
        ry, rg, rb = torch.autograd.grad((want * go.double()).sum(), [yd, bn.weight, bn.bias])
    out, mean, rstd = HT.bn_fwd(y.cuda(), gamma.cuda(), beta.cuda(), rm, rv, 0.1, 1e-5, True)
    dx, dg, db = HT.bn_bwd(go.cuda(), y.cuda(), mean, rstd, gamma.cuda(), True)
    if M > 1:
        assert rel(out, want) <= 2e-6
        assert rel(rm, bn.running_mean) <= 1e-6 and rel(rv, bn.running_var) <= 1e-6
        assert rel(dx, ry) <= 2e-5 and rel(dg, rg) <= 1e-5 and rel(db, rb) <= 1e-5, (rel(dx, ry), rel(dg, rg), rel(db, rb))
    else:
        assert torch.isfinite(out).all() and torch.isfinite(dx).all()
    # fixed statistics (eval-mode BatchNorm inside a graph)
    bn.eval()
    want = bn(yd)
    ry, rg, rb = torch.autograd.grad((want * go.double()).sum(), [yd, bn.weight, bn.bias])
    rm2, rv2 = bn.running_mean.float().cuda(), bn.running_var.float().cuda()
    out, mean, rstd = HT.bn_fwd(y.cuda(), gamma.cuda(), beta.cuda(), rm2, rv2, 0.1, 1e-5, False)
    dx, dg, db = HT.bn_bwd(go.cuda(), y.cuda(), mean, rstd, gamma.cuda(), False)
    assert rel(out, want) <= 2e-6 and rel(dx, ry) <= 1e-6 and rel(dg, rg) <= 1e-5 and rel(db, rb) <= 1e-5
    assert torch.equal(rm2.cpu(), bn.running_mean.float())       # untouched


def test_prelu_shared_slope(HT):
    g = torch.Generator().manual_seed(4)
    u = torch.randn(3, 7, 9, 32, generator=g)
    go = torch.randn(3, 7, 9, 32, generator=g)
    a = torch.tensor([0.2])
    ud, ad = u.double().requires_grad_(True), a.double().requires_grad_(True)
    want = F.prelu(ud, ad)
    ru, ra = torch.autograd.grad((want * go.double()).sum(), [ud, ad])
    v = HT.prelu_fwd(u.cuda(), a.cuda())
    du, da = HT.prelu_bwd(u.cuda(), go.cuda(), a.cuda())
    assert rel(v, want) == 0.0 or rel(v, want) <= 1e-7
    assert rel(du, ru) <= 1e-7 and rel(da, ra) <= 1e-5, (rel(du, ru), rel(da, ra))


@pytest.mark.parametrize("h,w,oh,ow", [(4, 4, 16, 16), (6, 5, 12, 10), (12, 13, 6, 6), (1, 3, 4, 12), (5, 7, 1, 1)])
def test_resize_align_corners_adjoint(HT, h, w, oh, ow):
    """nn.Upsample(align_corners=True) x2 / x4 / x0.5 (cod.py:709,733,737), a one-row map and a 1x1 target."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, h, w, 8, generator=g)
    go = torch.randn(2, oh, ow, 8, generator=g)
    xd = x.double().requires_grad_(True)
    want = nhwc(H.resize_bilinear(nchw(xd), oh, ow, True))
    rx, = torch.autograd.grad((want * go.double()).sum(), [xd])
    xc = x.cuda().requires_grad_(True)
    got = HT.ResizeLdFn.apply(xc, (oh, ow), True)
    (got * go.cuda()).sum().backward()
    assert rel(got, want) <= 1e-6 and rel(xc.grad, rx) <= 1e-6, (rel(got, want), rel(xc.grad, rx))


def test_resize_half_pixel_adjoint(HT):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 5, 6, 4, generator=g)
    go = torch.randn(1, 11, 9, 4, generator=g)
    xd = x.double().requires_grad_(True)
    want = nhwc(H.resize_bilinear(nchw(xd), 11, 9, False))
    rx, = torch.autograd.grad((want * go.double()).sum(), [xd])
    xc = x.cuda().requires_grad_(True)
    got = HT.ResizeLdFn.apply(xc, (11, 9), False)
    (got * go.cuda()).sum().backward()
    assert rel(got, want) <= 1e-6 and rel(xc.grad, rx) <= 1e-6


def _mode(name):
    from dgtd_b200.twig.ops.capi import BF16, F32
    return BF16 if name == "bf16" else F32


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("Cin,Cout,k,s,p,hw,train", [(64, 32, 1, 1, 0, (9, 13), True), (96, 32, 3, 1, 1, (12, 12), True),
                                                      (64, 32, 8, 4, 2, (16, 24), True), (320, 32, 1, 1, 0, (8, 8), False),
                                                      (96, 32, 3, 1, 1, (5, 7), False)])
def test_conv_bn_function(HT, Cin, Cout, k, s, p, hw, train, prec):
    """BasicConv2d under autograd for the three conv geometries of the decoder (1x1, 3x3 pad 1, 8x8 stride 4 pad 2),
    batch and running statistics; bf16 operands on tcgen05 where the map is large enough."""
    g = torch.Generator().manual_seed(7)
    B = 4
    x = torch.randn(B, hw[0], hw[1], Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    gamma, beta = 1 + 0.1 * torch.randn(Cout, generator=g), 0.1 * torch.randn(Cout, generator=g)
    rm, rv = 0.1 * torch.randn(Cout, generator=g), 1 + 0.3 * torch.rand(Cout, generator=g)
    params = {"conv.weight": w.double().requires_grad_(True), "bn.weight": gamma.double().requires_grad_(True),
              "bn.bias": beta.double().requires_grad_(True), "bn.running_mean": rm.double(), "bn.running_var": rv.double()}
    xd = x.double().requires_grad_(True)
    want = nhwc(H.basic_conv(nchw(xd), params, stride=s, padding=p, train=train))
    go = torch.randn(want.shape, generator=g)
    ref = torch.autograd.grad((want * go.double()).sum(), [xd, params["conv.weight"], params["bn.weight"], params["bn.bias"]])
    leaves = [t.cuda().requires_grad_(True) for t in (x, w, gamma, beta)]
    rmc, rvc = rm.cuda(), rv.cuda()
    got = HT.ConvBnFn.apply(*leaves, rmc if not train else rmc.clone(), rvc if not train else rvc.clone(),
                            (k, s, p, _mode(prec), 1e-5, 0.1, train))
    (got * go.cuda()).sum().backward()
    tol, err = (1e-4, rel) if prec == "fp32" else (5e-2, rel2)
    errs = [err(got, want)] + [err(t.grad, r) for t, r in zip(leaves, ref)]
    assert max(errs) <= tol, errs


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("C,hw", [(32, (6, 6)), (64, (24, 20)), (96, (12, 16))])
def test_cab_function(HT, C, hw, prec):
    g = torch.Generator().manual_seed(8)
    B = 2
    x = torch.randn(B, hw[0], hw[1], C, generator=g)
    p = {"body.0.weight": torch.randn(C, C, 3, 3, generator=g) / (9 * C) ** 0.5, "body.1.weight": torch.tensor([0.2]),
         "body.2.weight": torch.randn(C, C, 3, 3, generator=g) / (9 * C) ** 0.5,
         "CA.conv_du.0.weight": torch.randn(C // 4, C, 1, 1, generator=g) / C ** 0.5,
         "CA.conv_du.2.weight": torch.randn(C, C // 4, 1, 1, generator=g) / (C // 4) ** 0.5}
    pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xd = x.double().requires_grad_(True)
    want = nhwc(H.cab(nchw(xd), pd))
    go = torch.randn(want.shape, generator=g)
    ref = torch.autograd.grad((want * go.double()).sum(), [xd] + list(pd.values()))
    leaves = [t.cuda().requires_grad_(True) for t in [x] + list(p.values())]
    got = HT.CabFn.apply(*leaves, _mode(prec))
    (got * go.cuda()).sum().backward()
    tol, err = (1e-4, rel) if prec == "fp32" else (5e-2, rel2)
    errs = [err(got, want)] + [err(t.grad, r) for t, r in zip(leaves, ref)]
    assert max(errs) <= tol, errs


def test_sam_and_head_functions(HT):
    g = torch.Generator().manual_seed(9)
    B, C = 3, 32
    xh, xl = torch.randn(B, 6, 7, C, generator=g), torch.randn(B, 6, 7, C, generator=g)
    p = {"fc.0.weight": torch.randn(2, C, generator=g) / C ** 0.5, "fc.2.weight": torch.randn(C, 2, generator=g),
         "fc_wight.0.weight": torch.randn(2, C, generator=g) / C ** 0.5, "fc_wight.2.weight": torch.randn(1, 2, generator=g)}
    hw_, hb_ = torch.randn(1, C, 1, 1, generator=g) / C ** 0.5, torch.randn(1, generator=g)
    pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xhd, xld = xh.double().requires_grad_(True), xl.double().requires_grad_(True)
    hwd, hbd = hw_.double().requires_grad_(True), hb_.double().requires_grad_(True)
    want = H.conv1x1_bias(H.sam(nchw(xhd), nchw(xld), pd), hwd, hbd)
    go = torch.randn(want.shape, generator=g)
    ref = torch.autograd.grad((want * go.double()).sum(), [xhd, xld] + list(pd.values()) + [hwd, hbd])
    leaves = [t.cuda().requires_grad_(True) for t in [xh, xl] + list(p.values()) + [hw_, hb_]]
    s = HT.SamFn.apply(*leaves[:6])
    got = HT.Head1Fn.apply(s, leaves[6], leaves[7])
    (got * go.cuda()).sum().backward()
    errs = [rel(got, want)] + [rel(t.grad, r) for t, r in zip(leaves, ref)]
    assert max(errs) <= 1e-4, errs


# ---- the whole model ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def net():
    common.package()
    from dgtd_b200.twig.model import hitnet
    m = hitnet.cod(win_size=22, filter_ratio=0.9, using_sam=True, using_depth=True, finetune=True, binary_thresh=0.2)
    common.hitnet_fixture_params_(m.hitnet, seed=0)
    common.perturb_regressor_(m.hitnet.backbone.prompt_encoder)
    return m.cuda()


def _case():
    fx = np.load(os.path.join(common.GOLDEN, "hitnet_train_128.npz"))
    S, B = int(fx["S"]), int(fx["B"])
    image, depth = common.synthetic_inputs(B, S, seed=7)
    _, label = common.loss_inputs(B, S, S, seed=11)
    return fx, image, depth, label


def _step(net, image, depth, label):
    for p in net.parameters():
        p.grad = None
    out = net(None, image.cuda(), label.cuda(), depth.cuda(), mode="loss")
    out["loss"].backward()
    # detached: a live loss keeps the autograd graph (and its AccumulateGrad nodes, bound to this stream) alive, which
    # would invalidate a later CUDA-graph capture of the same parameters
    return out["loss"].detach(), {k: p.grad for k, p in net.hitnet.named_parameters()}


def _train_mode(net):
    net.train()
    net.hitnet.backbone.eval()          # DropPath off (the fixture / oracle have none); the decoder's BatchNorms train


def test_full_model_loss_and_gradients_match_reference_fixture(net):
    """`cod.forward(mode='loss')` + backward at 128^2, B = 2, decoder in train(): the loss, all 845 gradients (norm and
    8 samples each) and the BatchNorm running statistics vs the fixture of the UNMODIFIED reference (float64), fp32
    tolerance 1e-4; the reference's grad-less parameters get no (or an exactly removable, ~0) gradient."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    fx, image, depth, label = _case()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    set_precision(net, "fp32")
    _train_mode(net)
    loss, got = _step(net, image, depth, label)
    # the SSIM term has no gradient path; the fixture's loss is the deep-supervision sum alone
    from dgtd_b200.twig.model import losses
    with torch.no_grad():
        emb1 = net.hitnet.backbone._forward_features_nhwc(image.cuda(), depth.cuda())[0]
    ssim = float(losses.ssim_constant(emb1, image.cuda()))
    assert abs(float(loss) - ssim - float(fx["loss"])) <= 2e-5 * abs(float(fx["loss"])), (float(loss), ssim, float(fx["loss"]))
    scale = max(float(fx[k][0]) for k in fx.files if k.startswith("g/"))
    worst, n = ("", 0.0), 0
    for k in fx.files:
        if not k.startswith("g/"):
            continue
        name = k[2:]
        want = torch.from_numpy(fx[k])
        gr = got[name].detach().double().cpu().flatten()
        step = max(1, gr.numel() // 8)
        have = torch.cat([gr.norm().reshape(1), gr[::step][:8]])
        e = float((have - want).abs().max() / want.abs().max())
        n += 1
        if e > worst[1]:
            worst = (name, e)
    print("checked", n, "gradients against the reference fixture; worst", worst)
    assert n >= 840 and worst[1] <= 1e-4, worst
    for name in [str(s) for s in fx["unused"]]:
        gr = got.get(name)
        assert gr is None or float(gr.abs().max()) <= 1e-6 * scale, name
    sd = net.hitnet.state_dict()
    for name in ("Translayer2_1.bn.running_mean", "Translayer2_1.bn.running_var", "conv4.bn.running_mean",
                 "conv4.bn.running_var"):
        assert rel(sd[name], torch.from_numpy(fx["after/" + name])) <= 1e-5, name
    assert int(sd["conv4.bn.num_batches_tracked"]) == 4 and int(sd["compress_out.bn.num_batches_tracked"]) == 3
    net.load_state_dict(state)
    set_precision(net, None)
    net.eval()


def test_full_model_gradients_match_oracle_everywhere(net):
    """Same step against float64 autograd of the oracle on EVERY element of every gradient (the fixture holds 9
    numbers per tensor)."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    _, image, depth, label = _case()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    set_precision(net, "fp32")
    _train_mode(net)
    sd = {k: v.detach().double().cpu().requires_grad_("running" not in k) for k, v in net.hitnet.state_dict().items()
          if v.dtype.is_floating_point}
    _, P1, P2 = H.hitnet_forward(image.double(), depth.double(), sd, train=True)
    oloss = L.deep_supervision_loss(P1, P2, label.double())
    names = [k for k, v in sd.items() if v.requires_grad]
    ref = dict(zip(names, torch.autograd.grad(oloss, [sd[k] for k in names], allow_unused=True)))
    _, got = _step(net, image, depth, label)
    slope_names = [k for k in names if k.endswith("body.1.weight")]
    ref[slope_names[0]] = sum(ref[k] for k in slope_names)       # ONE shared nn.PReLU() (cod.py:686)
    scale = max(float(r.abs().max()) for r in ref.values() if r is not None)
    worst, n = ("", 0.0), 0
    for k, gr in got.items():
        r = ref.get(k)
        if r is None or float(r.abs().max()) < 1e-13:
            assert gr is None or float(gr.abs().max()) <= 1e-6 * scale, k
            continue
        assert gr is not None, k
        n += 1
        e = rel(gr, r)
        if e > worst[1]:
            worst = (k, e)
    print("checked", n, "full gradients; worst", worst)
    assert n >= 840 and worst[1] <= 1e-4, worst
    net.load_state_dict(state)
    set_precision(net, None)
    net.eval()


def test_full_model_bf16_gradients_track_fp32(net):
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    _, image, depth, label = _case()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    _train_mode(net)
    set_precision(net, "fp32")
    l32, exact = _step(net, image, depth, label)
    exact = {k: v.clone() for k, v in exact.items() if v is not None}
    net.load_state_dict(state)
    set_precision(net, "bf16")
    l16, got = _step(net, image, depth, label)
    assert abs(float(l16) - float(l32)) <= 3e-2 * abs(float(l32)), (float(l16), float(l32))
    bad = []
    for k, r in exact.items():
        if r.numel() < 4096:
            continue
        gq, rr = got[k].double().flatten(), r.double().flatten()
        cos = float((gq @ rr) / (gq.norm() * rr.norm()).clamp_min(1e-300))
        ratio = float(gq.norm() / rr.norm().clamp_min(1e-300))
        if cos < 0.98 or abs(ratio - 1.0) > 0.08:
            bad.append((k, cos, ratio))
    assert not bad, bad[:8]
    net.load_state_dict(state)
    set_precision(net, None)
    net.eval()


def test_train_mode_without_grad_uses_batch_statistics(net):
    """train() under no_grad (what `model.train(); with torch.no_grad(): model(...)` does in the reference): batch
    statistics, running buffers updated; eval() afterwards is the fused inference path again."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    _, image, depth, label = _case()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    set_precision(net, "fp32")
    _train_mode(net)
    sd = {k: v.detach().double().cpu() for k, v in net.hitnet.state_dict().items() if v.dtype.is_floating_point}
    with torch.no_grad():
        _, P1, P2 = net.hitnet(image.cuda(), depth.cuda())
        _, rP1, rP2 = H.hitnet_forward(image.double(), depth.double(), sd, train=True)
    assert rel(P2, rP2) <= 1e-4 and rel(P1[-1], rP1[-1]) <= 1e-4
    assert not torch.equal(net.hitnet.state_dict()["conv4.bn.running_mean"], state["hitnet.conv4.bn.running_mean"])
    net.load_state_dict(state)
    net.eval()
    with torch.no_grad():
        _, P1e, P2e = net.hitnet(image.cuda(), depth.cuda())
        sd = {k: v.detach().double().cpu() for k, v in net.hitnet.state_dict().items() if v.dtype.is_floating_point}
        _, eP1, eP2 = H.hitnet_forward(image.double(), depth.double(), sd, train=False)
    assert rel(P2e, eP2) <= 1e-4
    set_precision(net, None)


def test_graphed_full_model_step_replays_the_eager_step(net):
    """twig/graphs.py::GraphedModelTrainStep: the captured `cod.forward(mode='loss')` + backward leaves the eager
    step's gradients in the flat buffer, replays follow new inputs, and the BatchNorm buffers advance inside the graph.
    bf16 mode has no atomics anywhere (tcgen05 / mma.sync / TMA kernels with fixed-order reductions): the replay is
    bit-identical to the eager step.  (fp32 mode keeps fp32 atomics in the exact attention / depthwise-3x3 backward
    kernels, so only bf16 mode is held to bit equality.)"""
    from dgtd_b200.twig import graphs
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    _, image, depth, label = _case()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    set_precision(net, "bf16")
    _train_mode(net)
    loss, eager = _step(net, image, depth, label)
    eager = {k: v.clone() for k, v in eager.items() if v is not None}
    for p in net.parameters():
        p.grad = None
    step = graphs.GraphedModelTrainStep(net, image.cuda(), depth.cuda(), label.cuda(), precision="bf16", warmup=2)
    tracked = int(net.hitnet.conv4.bn.num_batches_tracked)
    got_loss = step()
    assert int(net.hitnet.conv4.bn.num_batches_tracked) == tracked + 4
    assert abs(float(got_loss) - float(loss)) <= 1e-6 * abs(float(loss)), (float(got_loss), float(loss))
    named = dict(net.hitnet.named_parameters())
    differ = []
    for k, r in eager.items():
        assert rel(named[k].grad, r) <= 1e-5, (k, rel(named[k].grad, r))
        if not torch.equal(named[k].grad, r):
            differ.append(k)
    assert not differ, (len(differ), differ[:6])
    image2, depth2 = common.synthetic_inputs(image.shape[0], image.shape[-1], seed=8)
    loss2 = float(step(image2.cuda(), depth2.cuda(), label.cuda()))
    assert loss2 != float(loss) and loss2 == loss2
    step.close()
    del step
    for p in net.parameters():
        p.grad = None
    net.load_state_dict(state)
    set_precision(net, None)
    net.eval()

"""the Hitnet iterative decoder + the `cod` predict head on the CUDA path, against the golden
fixture from the unmodified reference (tests/golden/hitnet_128.npz) and the float64 oracle.  This is where the
north-star mask criterion is checked: binarised masks identical to the reference's in fp32 mode, Hamming
distance reported (and bounded by the reference's own bf16-autocast distance) in bf16 mode."""
import math
import os

import numpy as np
import pytest
import torch

import common
from oracle import hitnet_ref as H

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def net():
    common.package()
    from dgtd_b200.twig.model import hitnet
    m = hitnet.Hitnet().eval()
    common.hitnet_fixture_params_(m, seed=0)
    return m.cuda()


@pytest.fixture(scope="module", params=["hitnet_128.npz", "hitnet_352.npz"])
def golden(request):
    """128^2 x 2 images, and BASELINE configs[0] (COD forward, batch 1, 352 x 352)."""
    return np.load(os.path.join(common.GOLDEN, request.param))


def test_state_dict_keys(net):
    sd = net.state_dict()
    assert len(sd) == 879                       # == the reference's Hitnet (checked key by key at authoring time)
    # the ONE shared PReLU slope of the default constructor argument (cod.py:686)
    slopes = {sd[k].data_ptr() for k in sd if k.endswith("body.1.weight")}
    assert len(slopes) == 1


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_matches_reference_golden(net, golden, precision):
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    S, B = int(golden["S"]), int(golden["B"])
    image, depth = common.synthetic_inputs(B, S, seed=7)
    set_precision(net, precision)
    try:
        emb1, P1, P2 = net(image.cuda(), depth.cuda())
    finally:
        set_precision(net, None)
    assert len(P1) == 4 and all(tuple(p.shape) == (B, 1, S, S) for p in P1) and tuple(P2.shape) == (B, 1, S, S)
    # fp32: 1e-4; bf16: no worse than 2x the reference's own bf16-autocast error
    tol = 1e-4 if precision == "fp32" else max(2.0 * float(golden["ref_bf16_relerr"]), 2e-2)
    sub = int(golden["sub"])
    for i, p in enumerate(P1):
        e = rel(p[:, :, ::sub, ::sub].cpu(), torch.from_numpy(golden[f"P1_{i}"]))
        assert e <= tol, (precision, i, e, tol)
    e = rel(P2[:, :, ::sub, ::sub].cpu(), torch.from_numpy(golden["P2"]))
    assert e <= tol, (precision, "P2", e, tol)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_predict_masks_match_reference(net, golden, precision):
    """cod.py:148-149 + binarisation at 0.5 and at binary_thresh = 0.2 (cod.yml:47)."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    S, B = int(golden["S"]), int(golden["B"])
    image, depth = common.synthetic_inputs(B, S, seed=7)
    set_precision(net, precision)
    try:
        _, logits = net.predict_logits(image.cuda(), depth.cuda(), (S, S))
    finally:
        set_precision(net, None)
    ref = torch.from_numpy(golden["logits"])
    got = logits.cpu().double()
    err = rel(got, ref)
    report = {"precision": precision, "logit_rel_err": err}
    for thr, key in ((0.5, "mask50"), (0.2, "mask20")):
        cut = math.log(thr / (1 - thr))
        ref_mask = np.unpackbits(golden[key])[:ref.numel()].reshape(ref.shape).astype(bool)
        assert np.array_equal(ref_mask, (torch.sigmoid(ref) > thr).numpy())
        got_mask = (got > cut).numpy()
        ham = int((got_mask != ref_mask).sum())
        report[f"hamming_{key}"] = ham
        if precision == "fp32":
            assert err <= 1e-4, err
            # identical masks; a bit may only differ where the reference logit is within the fp32 error of the cut
            band = (ref - cut).abs().numpy() <= 1e-4 * float(ref.abs().max())
            assert int(((got_mask != ref_mask) & ~band).sum()) == 0
            assert ham <= int(band.sum())
        else:
            ref_ham = int(golden[f"ref_bf16_hamming_{key}"])
            assert ham <= max(2 * ref_ham, int(0.02 * ref.numel())), (key, ham, ref_ham)
    print("mask parity:", report)


def _sd64(m):
    return {k: v.detach().double().cpu() for k, v in m.state_dict().items() if v.dtype.is_floating_point}


def test_decoder_modules_against_the_oracle(net):
    """CAB / BasicConv2d (1x1, 3x3, 8x8 stride 4) / SAM forwards (reference signatures, NCHW in and out) vs the
    float64 restatement, non-square maps, batch > 1 (the gates are per image)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 64, 20, 12, generator=g)
    got = net.decoder_level1[0](x.cuda())
    ref = H.cab(x.double(), _sd64(net.decoder_level1[0]))
    assert rel(got.cpu(), ref) <= 1e-5
    x = torch.randn(2, 96, 9, 14, generator=g)
    got = net.decoder_level2[1](x.cuda())
    assert rel(got.cpu(), H.cab(x.double(), _sd64(net.decoder_level2[1]))) <= 1e-5
    x = torch.randn(2, 64, 24, 16, generator=g)
    got = net.compress_out(x.cuda())
    ref = H.basic_conv(x.double(), _sd64(net.compress_out), stride=4, padding=2)
    assert got.shape == ref.shape and rel(got.cpu(), ref) <= 1e-5
    x = torch.randn(2, 96, 10, 7, generator=g)
    assert rel(net.conv4(x.cuda()).cpu(), H.basic_conv(x.double(), _sd64(net.conv4), padding=1)) <= 1e-5
    x = torch.randn(2, 512, 5, 4, generator=g)
    assert rel(net.Translayer4_1(x.cuda()).cpu(), H.basic_conv(x.double(), _sd64(net.Translayer4_1))) <= 1e-5
    a, b = torch.randn(3, 32, 11, 6, generator=g), torch.randn(3, 32, 11, 6, generator=g)
    assert rel(net.SAM(a.cuda(), b.cuda()).cpu(), H.sam(a.double(), b.double(), _sd64(net.SAM))) <= 1e-5


@pytest.mark.parametrize("align", [True, False])
def test_resize_both_conventions(align):
    common.package()
    from dgtd_b200.twig.ops.functions import hitnet_func as HF
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 7, 5, 8, generator=g)                       # NHWC
    for size in ((14, 10), (3, 2), (28, 20), (7, 5)):
        got = HF.resize_ld(x.cuda(), size, align).cpu()
        ref = H.resize_bilinear(x.permute(0, 3, 1, 2).double(), size[0], size[1], align).permute(0, 2, 3, 1)
        assert rel(got, ref) <= 1e-6, (align, size)
    # writing into a channel slice of a wider tensor (the torch.cat operand)
    wide = torch.zeros(2, 14, 10, 24, device="cuda")
    HF.resize_ld(x.cuda(), (14, 10), align, out=wide[..., 8:16])
    ref = H.resize_bilinear(x.permute(0, 3, 1, 2).double(), 14, 10, align).permute(0, 2, 3, 1)
    assert rel(wide[..., 8:16].cpu(), ref) <= 1e-6 and float(wide[..., :8].abs().max()) == 0 and float(wide[..., 16:].abs().max()) == 0


def test_decode_against_the_oracle_on_random_features(net):
    """The decoder alone (4 feedback iterations) on random backbone maps of a non-square image."""
    g = torch.Generator().manual_seed(9)
    B, h, w = 2, 8, 6                                              # stride-32 grid
    feats = [torch.randn(B, c, h * s, w * s, generator=g) for c, s in ((64, 8), (128, 4), (320, 2), (512, 1))]
    p = _sd64(net)
    ref_preds, ref_p2 = H.decode([f.double() for f in feats], p)
    nhwc = [f.permute(0, 2, 3, 1).contiguous().cuda() for f in feats]
    preds, p2, _ = net.decode(nhwc)
    for a, b in zip(list(preds) + [p2], list(ref_preds) + [ref_p2]):
        assert a.shape == b.shape and rel(a.cpu(), b) <= 1e-4, rel(a.cpu(), b)


def test_train_mode_uses_batch_statistics_like_nn_batchnorm(net):
    """BasicConv2d in train() (cod.py:362): batch statistics + running update, checked against nn.Conv2d + nn.BatchNorm2d
    in float64 on the same weights (the whole decoder in train(): tests/test_gpu_hitnet_train.py)."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    set_precision(net, "fp32")
    x = torch.randn(3, 96, 6, 5, generator=torch.Generator().manual_seed(2))
    m = net.conv4
    state = {k: v.clone() for k, v in m.state_dict().items()}
    ref_conv = torch.nn.Conv2d(96, 32, 3, padding=1, bias=False).double()
    ref_bn = torch.nn.BatchNorm2d(32).double().train()
    ref_conv.load_state_dict({"weight": state["conv.weight"].double().cpu()})
    ref_bn.load_state_dict({k[3:]: v.double().cpu() if v.dtype.is_floating_point else v.cpu() for k, v in state.items()
                            if k.startswith("bn.")})
    want = ref_bn(ref_conv(x.double()))
    m.train()
    try:
        with torch.no_grad():
            got = m(x.cuda())
        assert rel(got.cpu(), want.detach()) <= 1e-5
        assert rel(m.bn.running_mean.cpu(), ref_bn.running_mean) <= 1e-5
        assert rel(m.bn.running_var.cpu(), ref_bn.running_var) <= 1e-5
        assert int(m.bn.num_batches_tracked) == int(ref_bn.num_batches_tracked)
    finally:
        m.load_state_dict(state)
        net.eval()
        set_precision(net, None)


def test_cod_modes():
    golden = np.load(os.path.join(common.GOLDEN, "hitnet_128.npz"))
    common.package()
    from dgtd_b200.twig.model import hitnet
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    m = hitnet.cod(win_size=22, filter_ratio=0.9, using_sam=True, using_depth=True, finetune=True, binary_thresh=0.2,
                   pretrain_sam=None, head=None).eval()
    common.hitnet_fixture_params_(m.hitnet, seed=0)
    m = m.cuda()
    set_precision(m, "fp32")
    S, B = int(golden["S"]), int(golden["B"])
    image, depth = common.synthetic_inputs(B, S, seed=7)
    label = (torch.rand(B, 1, S, S, generator=torch.Generator().manual_seed(1)) > 0.5).float().cuda()
    depth_list = [d for d in depth.cuda()]                        # the reference passes a list of (1,S,S) maps
    out = m(None, image.cuda(), label, depth_list, mode="tensor")
    assert rel(out.cpu(), torch.from_numpy(golden["logits"])) <= 1e-4
    prob, lab = m(None, image.cuda(), label, depth_list, mode="predict")
    assert lab is label and rel(prob.cpu(), torch.sigmoid(torch.from_numpy(golden["logits"]))) <= 1e-4
    loss = m(None, image.cuda(), label, depth_list, mode="loss")["loss"]
    # the whole training loss of cod.py:135-146 (deep supervision + the SSIM constant) from the oracle
    from oracle import loss_ref as L
    p = _sd64(m.hitnet)
    e1, P1, P2 = H.hitnet_forward(image.double(), depth.double(), p)
    ref = L.total_loss(e1, P1, P2, image.double(), label.cpu().double())
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    with pytest.raises(NotImplementedError):
        m(None, image.cuda(), label, depth_list, mode="nope")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_smeasure_and_mae_deltas_against_the_reference_prediction(net, golden, precision):
    """North-star report: S-measure / MAE of the new path's prediction next to those of the reference's prediction
    (golden logits) on the same synthetic label, through the GPU metrics and through the oracle metrics."""
    import warnings
    from dgtd_b200.twig.metric import sod_metrics
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    from dgtd_b200.twig.ops.functions import hitnet_func as HF
    from oracle import metrics_ref as M
    S, B = int(golden["S"]), int(golden["B"])
    image, depth = common.synthetic_inputs(B, S, seed=7)
    label = torch.zeros(B, 1, S, S)
    label[:, :, S // 4: 3 * S // 4, S // 8: S // 2] = 1.0
    set_precision(net, precision)
    try:
        _, logits = net.predict_logits(image.cuda(), depth.cuda(), (S, S))
    finally:
        set_precision(net, None)
    ours = sod_metrics(HF.sigmoid(logits), label.cuda()).cpu().numpy()
    ref_prob = torch.sigmoid(torch.from_numpy(golden["logits"]).float())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = np.array([(M.mae_one(p, g), M.smeasure_one(p, g))
                        for p, g in zip(M.quantise(ref_prob.numpy()), M.quantise(label.numpy()))])
    d_mae, d_sm = np.abs(ours[:, 0] - ref[:, 0]).max(), np.abs(ours[:, 1] - ref[:, 1]).max()
    print(f"metric deltas [{precision}]: |dMAE| = {d_mae:.3e}, |dS| = {d_sm:.3e}  (reference MAE {ref[:, 0].mean():.4f}, "
          f"S {ref[:, 1].mean():.4f})")
    tol = 1e-4 if precision == "fp32" else 1e-2
    assert d_mae <= tol and d_sm <= tol, (d_mae, d_sm)


def test_bf16_tensor_core_decoder_convs(net):
    """bf16 mode of the decoder: im2col(bf16) + tcgen05 GEMM with folded BatchNorm; against the float64 oracle on the
    same parameters.  Stated tolerance 2e-2 relative (bf16 operands, fp32 accumulation, K up to 4096)."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    g = torch.Generator().manual_seed(13)
    set_precision(net, "bf16")
    try:
        x = torch.randn(3, 96, 12, 20, generator=g)
        e = rel(net.decoder_level2[0](x.cuda()).cpu(), H.cab(x.double(), _sd64(net.decoder_level2[0])))
        assert e <= 2e-2, e
        x = torch.randn(2, 64, 24, 16, generator=g)
        e = rel(net.compress_out(x.cuda()).cpu(), H.basic_conv(x.double(), _sd64(net.compress_out), stride=4, padding=2))
        assert e <= 2e-2, e
        x = torch.randn(2, 320, 6, 5, generator=g)
        e = rel(net.Translayer3_1(x.cuda()).cpu(), H.basic_conv(x.double(), _sd64(net.Translayer3_1)))
        assert e <= 2e-2, e
        feats = [torch.randn(2, c, 8 * s, 6 * s, generator=g) for c, s in ((64, 8), (128, 4), (320, 2), (512, 1))]
        ref_preds, ref_p2 = H.decode([f.double() for f in feats], _sd64(net))
        preds, p2, _ = net.decode([f.permute(0, 2, 3, 1).contiguous().cuda() for f in feats])
        for a, b in zip(list(preds) + [p2], list(ref_preds) + [ref_p2]):
            assert rel(a.cpu(), b) <= 3e-2, rel(a.cpu(), b)
    finally:
        set_precision(net, None)


@pytest.mark.parametrize("shape", [(2, 64, 24, 16, 64), (1, 96, 11, 23, 96), (3, 32, 12, 12, 32), (2, 96, 48, 48, 32),
                                   (1, 64, 7, 5, 64)])
def test_implicit_conv3x3_tcgen05(shape):
    """dgtd_conv3x3_tc_fwd (shifted TMA boxes, zero fill = padding, masked partial patches) against a float64
    conv on the same bf16-rounded operands: <= 1e-4 (only the fp32 accumulation order differs); also into a
    channel slice of a wider output."""
    common.package()
    from dgtd_b200.twig.model.hitnet import _tap_major_padded_bf16
    from dgtd_b200.twig.ops.functions import hitnet_func as HF
    B, Cin, h, w, Cout = shape
    g = torch.Generator().manual_seed(h * w + Cin)
    x = torch.randn(B, Cin, h, w, generator=g)
    wt = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    shift = torch.randn(Cout, generator=g)
    slope = torch.tensor([0.2])
    xq = torch.where(x >= 0, x, 0.2 * x).to(torch.bfloat16).double()
    wq = wt.to(torch.bfloat16).double()
    ref = torch.nn.functional.conv2d(xq, wq, shift.double(), padding=1).permute(0, 2, 3, 1)
    wide = torch.full((B, h, w, Cout + 32), 7.0, device="cuda")
    HF.conv3_tc(x.permute(0, 2, 3, 1).contiguous().cuda(), _tap_major_padded_bf16(wt.cuda()), shift=shift.cuda(),
                prelu_in=slope.cuda(), out=wide[..., 16:16 + Cout])
    assert rel(wide[..., 16:16 + Cout].cpu(), ref) <= 1e-4
    assert float((wide[..., :16] - 7.0).abs().max()) == 0 and float((wide[..., 16 + Cout:] - 7.0).abs().max()) == 0


def test_graphed_predict_is_bit_identical_to_the_eager_forward():
    """twig/graphs.py::GraphedPredict: the captured predict step replayed on new inputs equals the eager call bit
    for bit (every kernel takes the caller's stream, nothing allocates behind torch's back)."""
    common.package()
    from dgtd_b200.twig import graphs
    from dgtd_b200.twig.model import hitnet
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    m = hitnet.cod(binary_thresh=0.2).eval()
    common.hitnet_fixture_params_(m.hitnet, seed=0)
    m = m.cuda()
    for precision in ("fp32", "bf16"):
        set_precision(m, precision)
        image, depth = common.synthetic_inputs(2, 128, seed=21)
        run = graphs.GraphedPredict(m, image.cuda(), depth.cuda())
        image2, depth2 = common.synthetic_inputs(2, 128, seed=22)
        got = run(image2.cuda(), depth2.cuda()).clone()
        _, want = m.hitnet.predict_logits(image2.cuda(), depth2.cuda(), (128, 128))
        assert torch.equal(got, want), precision
        got1 = run(image.cuda(), depth.cuda()).clone()
        _, want1 = m.hitnet.predict_logits(image.cuda(), depth.cuda(), (128, 128))
        assert torch.equal(got1, want1) and not torch.equal(got1, got)


def test_host_pipeline_with_the_full_model_and_metrics():
    """twig/pipeline.py with a `forward` callable: pinned image + depth + label in, per-image (MAE, S-measure) out,
    equal to the direct calls batch by batch."""
    common.package()
    from dgtd_b200.twig.metric import sod_metrics
    from dgtd_b200.twig.model import hitnet
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    from dgtd_b200.twig.pipeline import HostPipeline
    m = hitnet.cod(binary_thresh=0.2).eval()
    common.hitnet_fixture_params_(m.hitnet, seed=0)
    m = m.cuda()
    set_precision(m, "bf16")
    batches = []
    for seed in range(3):
        image, depth = common.synthetic_inputs(2, 128, seed=30 + seed)
        label = (torch.rand(2, 1, 128, 128, generator=torch.Generator().manual_seed(seed)) > 0.6).float()
        batches.append(tuple(t.pin_memory() for t in (image, depth, label)))

    def fwd(im, dp, lb):
        prob, _ = m(None, im, lb, dp, mode="predict")
        return (sod_metrics(prob, lb),)
    outs = [torch.empty(2, 2, dtype=torch.float64).pin_memory() for _ in range(2)]
    pipe = HostPipeline(None, None, device=torch.device("cuda:0"), forward=fwd)
    got = []
    for i, host, done in pipe.run(iter(batches), lambda v: v, outs):
        done.synchronize()
        got.append(host.clone())
    assert len(got) == 3
    for (im, dp, lb), g in zip(batches, got):
        want = fwd(im.cuda(), dp.cuda(), lb.cuda())[0].cpu()
        assert torch.equal(g, want)


# ---- randomized shape sweeps (hypothesis), oracle = float64 torch on the same inputs ---------------------------
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

_SET = dict(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


@settings(**_SET)
@given(B=st.integers(1, 3), cin=st.sampled_from([4, 8, 32, 96]), cout=st.sampled_from([4, 32, 64, 96]),
       h=st.integers(3, 19), w=st.integers(3, 21), cfg=st.sampled_from([(1, 1, 0), (3, 1, 1), (8, 4, 2), (3, 2, 1)]),
       pre=st.booleans(), res=st.booleans(), seed=st.integers(0, 1000))
def test_sweep_conv_affine(B, cin, cout, h, w, cfg, pre, res, seed):
    """fp32 implicit GEMM with folded-BN / PReLU / residual epilogue, strided output slice."""
    common.package()
    from dgtd_b200.twig.ops.functions import hitnet_func as HF
    k, s, p = cfg
    if h + 2 * p < k or w + 2 * p < k:
        return
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (k * cin ** 0.5)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    slope = torch.tensor([0.25])
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), None, stride=s, padding=p)
    ref = ref * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
    if pre:
        ref = torch.where(ref >= 0, ref, 0.25 * ref)
    oh, ow = ref.shape[2], ref.shape[3]
    r = torch.randn(B, oh, ow, cout, generator=g)
    if res:
        ref = ref + r.permute(0, 3, 1, 2).double()
    wide = torch.zeros(B, oh, ow, cout + 8, device="cuda")
    HF.conv_affine(x.permute(0, 2, 3, 1).contiguous().cuda(), wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().cuda(),
                   (oh, ow), k, s, -p, scale=scale.cuda(), shift=shift.cuda(), prelu=slope.cuda() if pre else None,
                   residual=r.cuda() if res else None, out=wide[..., 4:4 + cout])
    assert rel(wide[..., 4:4 + cout].cpu().permute(0, 3, 1, 2), ref) <= 1e-5
    assert float(wide[..., :4].abs().max()) == 0 and float(wide[..., 4 + cout:].abs().max()) == 0


@settings(**_SET)
@given(B=st.integers(1, 4), C=st.sampled_from([4, 32, 64, 96]), h=st.integers(1, 40), w=st.integers(1, 40),
       red=st.sampled_from([2, 4, 16]), seed=st.integers(0, 1000))
def test_sweep_gates_and_gated_sum(B, C, h, w, red, seed):
    """channel sums -> gate MLP -> res * gate * scalar + x, against float64."""
    common.package()
    from dgtd_b200.twig.ops.functions import hitnet_func as HF
    g = torch.Generator().manual_seed(seed)
    cr = max(1, C // red)
    a, b_ = torch.randn(B, h, w, C, generator=g), torch.randn(B, h, w, C, generator=g)
    w1, w2 = torch.randn(cr, C, generator=g), torch.randn(C, cr, generator=g)
    w3 = torch.randn(1, cr, generator=g)
    part, hw = HF.channel_sums(a.cuda())
    gate = HF.channel_gate(part, hw, w1.cuda(), w2.cuda())
    scal = HF.channel_gate(part, hw, w1.cuda(), w3.cuda())
    mean = a.double().mean(dim=(1, 2))
    hid = torch.relu(mean @ w1.double().t())
    gref, sref = torch.sigmoid(hid @ w2.double().t()), torch.sigmoid(hid @ w3.double().t())
    assert rel(gate.cpu(), gref) <= 1e-5 and rel(scal.cpu(), sref) <= 1e-5
    out = HF.gated_sum(a.cuda(), ga=gate, sa=scal.reshape(-1), b=b_.cuda())
    ref = a.double() * gref[:, None, None, :] * sref[:, None, None, :] + b_.double()
    assert rel(out.cpu(), ref) <= 1e-5


@settings(**_SET)
@given(B=st.integers(1, 3), h=st.integers(1, 50), w=st.integers(2, 70), thr=st.floats(0.05, 0.95),
       seed=st.integers(0, 1000))
def test_sweep_metrics(B, h, w, thr, seed):
    """All four evaluators on random maps of random size against the numpy restatement."""
    import warnings
    from oracle import metrics_ref as M
    common.package()
    from dgtd_b200.twig.metric import sod_metrics
    if h < 2:
        h = 2
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand(B, 1, h, w, generator=g)
    gt = (torch.rand(B, 1, h, w, generator=g) > thr).float()
    vals, cur = sod_metrics(pred.cuda(), gt.cuda(), curves=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, (p, q) in enumerate(zip(M.quantise(pred.numpy()), M.quantise(gt.numpy()))):
            assert abs(float(vals[i, 0]) - M.mae_one(p, q)) <= 1e-12
            assert abs(float(vals[i, 1]) - M.smeasure_one(p, q)) <= 1e-12
            assert np.abs(cur[i, 0].cpu().numpy() - M.fmeasure_curve_one(p, q)).max() <= 1e-12
            assert np.abs(cur[i, 1].cpu().numpy() - M.emeasure_curve_one(p, q)).max() <= 1e-12


def test_highres_768_full_model_modes_agree(net):
    """BASELINE configs[4] geometry (768 x 768: 192 / 96 / 48 / 24 token grids, 576 reduced keys): the predict path
    runs end to end and the bf16 tensor-core mode stays within the bf16 tolerance of the exact fp32 mode."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    image, depth = common.synthetic_inputs(1, 768, seed=5)
    out = {}
    for precision in ("fp32", "bf16"):
        set_precision(net, precision)
        try:
            _, out[precision] = net.predict_logits(image.cuda(), depth.cuda(), (768, 768))
        finally:
            set_precision(net, None)
    assert tuple(out["fp32"].shape) == (1, 1, 768, 768) and torch.isfinite(out["fp32"]).all()
    assert rel(out["bf16"].cpu(), out["fp32"].cpu()) <= 6e-2

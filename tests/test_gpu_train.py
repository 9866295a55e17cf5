"""Training path (config/sod.yml semantics): gradients of every parameter of the hot path through
the CUDA autograd Functions vs float64 autograd of the CPU oracle on the same seeded inputs.
Small image (96x96 -> trunk 24/12/6/3) so the oracle backward finishes in seconds."""
import pytest
import torch

import common
from oracle import texture_diffuser_ref as O

pytestmark = pytest.mark.gpu


def _oracle_grads(enc, dec, image, depth, gout_e3, gout_tok, dtype=torch.float64, tokens_in_loss=True):
    pe = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    pd = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    e1, e3, toks = O.texture_prompts(image.to(dtype), depth.to(dtype), pe, pd)
    loss = (e3 * gout_e3.to(dtype)).sum()
    if tokens_in_loss:
        for s in range(4):
            for i, t in enumerate(toks[s]):
                loss = loss + (t * gout_tok[s][i].to(dtype)).sum()
    names = [k for k in pe if not k.startswith("adaptor")]
    ge = torch.autograd.grad(loss, [pe[k] for k in names] + list(pd.values()), allow_unused=True)
    out = {"enc." + k: g for k, g in zip(names, ge[:len(names)])}
    out.update({"dec." + k: g for k, g in zip(pd, ge[len(names):])})
    return (e1, e3, toks), out


def _path_inputs(S, B, seed):
    image, depth = common.synthetic_inputs(B, S, seed=seed)
    grids = common.pvt_token_grids((S, S))
    g = torch.Generator().manual_seed(7)
    gout_e3 = torch.randn(B, 24, S // 4, S // 4, generator=g) * 1e-2
    gout_tok = [[torch.randn(B, grids[s][0] * grids[s][1], e, generator=g) * 1e-2 for _ in range(n)]
                for s, (e, n) in enumerate(zip(common.PVT_EMBED_DIMS, common.PVT_DEPTHS))]
    return image, depth, gout_e3, gout_tok


def _our_grads(TD, enc, dec, image, depth, gout_e3, gout_tok, tokens_in_loss=True):
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    e1, e3, toks = TD.texture_prompts_train(enc, dec, image.cuda(), depth.cuda())
    loss = (e3 * gout_e3.cuda()).sum()
    if tokens_in_loss:
        for s in range(4):
            for i, t in enumerate(toks[s]):
                loss = loss + (t * gout_tok[s][i].cuda()).sum()
    loss.backward()
    return (e1, e3, toks), {pre + k: p.grad for pre, mod in (("enc.", enc), ("dec.", dec)) for k, p in mod.named_parameters()}


def test_full_path_gradients_match_oracle():
    """96^2 (trunk 24/12/6/3): every one of the 452 parameter gradients at the north-star fp32 tolerance 1e-4."""
    TD = common.package()
    S, B = 96, 2
    enc, dec = TD.build_texture_diffuser(seed=0)
    common.perturb_regressor_(enc)
    image, depth, gout_e3, gout_tok = _path_inputs(S, B, 3)
    (r1, r3, rtoks), ref = _oracle_grads(enc, dec, image, depth, gout_e3, gout_tok)
    enc, dec = enc.cuda().eval(), dec.cuda().eval()      # eval: DropPath off (masks are tested separately)
    (e1, e3, toks), got = _our_grads(TD, enc, dec, image, depth, gout_e3, gout_tok)
    assert common.rel_err(e3, r3) <= 1e-4 and common.rel_err(e1, r1) <= 1e-4
    for s in range(4):
        for i, t in enumerate(toks[s]):
            assert common.rel_err(t, rtoks[s][i]) <= 1e-4
    worst = ("", 0.0)
    n_checked = 0
    for k, gr in got.items():
        r = ref.get(k)
        if r is None:                                   # adaptor.* never receives a gradient (cod.py:1251)
            assert gr is None or float(gr.abs().max()) == 0.0, k
            continue
        assert gr is not None, k
        err = common.rel_err(gr, r)
        n_checked += 1
        if err > worst[1]:
            worst = (k, err)
    print("checked", n_checked, "gradients; worst", worst)
    assert n_checked == 358 - 2 + 96
    assert worst[1] <= 1e-4, worst     # north-star fp32 tolerance; fp32 accumulation through 36 blocks vs float64


def _quadratic_loss(e3, toks):
    """0.5 * sum of squares of embedding3 and of every token tensor, each normalised by its element count: a COHERENT
    upstream gradient (g = t / numel), so that the pixel sums of the weight gradients do not cancel."""
    loss = 0.5 * (e3 * e3).mean()
    for row in toks:
        for t in row:
            loss = loss + 0.5 * (t * t).mean()
    return loss


def test_gradients_at_config_size_384():
    """BASELINE configs[2] geometry (384^2: trunk 96/48/24/12, decoder bank with the folded stride-2/4/8 convs), B = 2:
    all 452 parameter gradients <= 1e-4 (north-star fp32 tolerance) of float64 autograd of the oracle.

    The loss is quadratic in the outputs (`_quadratic_loss`), not a random projection: with RANDOM upstream gradients
    the 18 432-pixel sums of the decoder weight gradients cancel to ~1 % of their terms, and the real embedding3
    (|max| 99) puts a few of the 7 million ReLU pre-activations within fp32 rounding of zero -- one flipped ReLU mask
    then moves a gradient by ~1e-3 of its max.  That is conditioning, not kernel error: the FLOAT64 oracle's own
    gradients move by up to 1.3e-2 under a 1e-6 relative perturbation of embedding3 (tools/debug_bank2.py), and the
    oracle run in float32 on the CPU is 4e-4 (median) / 5e-3 (worst) away from its float64 self on that loss.  The
    random-projection form stays in `test_full_path_gradients_match_oracle` (96^2), where it is well conditioned;
    trunk-only and decoder-bank-only checks at this size are (a) below and `test_decoder_bank_at_config_geometry`."""
    TD = common.package()
    S, B = 384, 2
    enc, dec = TD.build_texture_diffuser(seed=0)
    common.perturb_regressor_(enc)
    image, depth, gout_e3, gout_tok = _path_inputs(S, B, 3)
    # float64 oracle: (a) random projection of embedding3 only (smooth: LayerNorm / GELU), (b) quadratic loss on everything
    pe = {k: v.detach().double().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    pd = {k: v.detach().double().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    r1, r3, rtoks = O.texture_prompts(image.double(), depth.double(), pe, pd)
    names = [k for k in pe if not k.startswith("adaptor")]
    ga = torch.autograd.grad((r3 * gout_e3.double()).sum(), [pe[k] for k in names], retain_graph=True)
    ref_a = {"enc." + k: g for k, g in zip(names, ga)}
    gb = torch.autograd.grad(_quadratic_loss(r3, rtoks), [pe[k] for k in names] + list(pd.values()))
    ref_b = {"enc." + k: g for k, g in zip(names, gb[:len(names)])}
    ref_b.update({"dec." + k: g for k, g in zip(pd, gb[len(names):])})

    enc, dec = enc.cuda().eval(), dec.cuda().eval()
    params = {pre + k: p for pre, mod in (("enc.", enc), ("dec.", dec)) for k, p in mod.named_parameters()}

    def run(loss_fn):
        for p in params.values():
            p.grad = None
        e1, e3, toks = TD.texture_prompts_train(enc, dec, image.cuda(), depth.cuda())
        loss_fn(e3, toks).backward()
        return e1, e3, toks

    e1, e3, toks = run(lambda e3, toks: (e3 * gout_e3.cuda()).sum())
    assert common.rel_err(e3, r3) <= 1e-4 and common.rel_err(e1, r1) <= 1e-4
    for s in range(4):
        for i, t in enumerate(toks[s]):
            assert common.rel_err(t, rtoks[s][i]) <= 1e-4
    errs = {k: common.rel_err(params[k].grad, r) for k, r in ref_a.items()}
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"384^2 (a) embedding3 projection: {len(errs)} gradients, worst {worst}")
    assert len(errs) == 356 and worst[1] <= 1e-4, worst

    run(_quadratic_loss)
    errs = {}
    for k, r in ref_b.items():
        assert params[k].grad is not None and torch.isfinite(params[k].grad).all(), k
        errs[k] = common.rel_err(params[k].grad, r)
    v = sorted(errs.values())
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"384^2 (b) quadratic loss: {len(errs)} gradients, median {v[len(v) // 2]:.2e} p95 {v[int(len(v) * 0.95)]:.2e} worst {worst}")
    assert len(errs) == 452 and worst[1] <= 1e-4, worst


def test_decoder_bank_at_config_geometry():
    """The 16 ShapePropDecoders + folded injection as one Function at the configs[2] geometry (B = 2, 96 x 96 map,
    token grids 96/48/24/12), fp32 mode, on a unit-variance input: all 96 parameter gradients and the input
    gradient <= 1e-4 of float64 autograd of plain conv / relu / bilinear (cod.py:1217-1221, 1471)."""
    import torch.nn.functional as F
    TD = common.package()
    from dgtd_b200.twig.ops import capi
    from dgtd_b200.twig.ops.functions import decoder_bank as DB
    _, dec = TD.build_texture_diffuser(seed=0)
    dec = dec.cuda()
    B, h = 2, 96
    g = torch.Generator().manual_seed(h)
    emb = torch.randn(B, h, h, 24, generator=g).cuda().requires_grad_(True)
    grids = [(h, h), (h // 2, h // 2), (h // 4, h // 4), (h // 8, h // 8)]
    cfg = {"stages": [(len(dec[s].decoder), grids[s]) for s in range(4)], "mode": capi.F32}
    params = [t for s in range(4) for d in dec[s].decoder for t in (d.decoder[0].weight, d.decoder[0].bias, d.decoder[2].weight,
                                                                      d.decoder[2].bias, d.decoder[4].weight, d.decoder[4].bias)]
    outs = DB.DecoderBankFn.apply(emb, cfg, *params)
    gouts = [torch.randn(o.shape, generator=g).cuda() * 1e-2 for o in outs]
    grads = torch.autograd.grad(sum((o * go).sum() for o, go in zip(outs, gouts)), [emb] + params)
    emb64 = emb.detach().double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    p64 = [p.detach().double().cpu().requires_grad_(True) for p in params]
    routs, i = [], 0
    for s in range(4):
        for _ in range(len(dec[s].decoder)):
            w1, b1, w2, b2, w3, b3 = p64[6 * i:6 * i + 6]
            y = F.conv2d(F.relu(F.conv2d(F.relu(F.conv2d(emb64, w1, b1, padding=1)), w2, b2, padding=1)), w3, b3, padding=1)
            if grids[s] != (h, h):
                y = F.interpolate(y, size=grids[s], mode="bilinear")
            routs.append(y.flatten(2).permute(0, 2, 1))
            i += 1
    assert max(common.rel_err(a, b) for a, b in zip(outs, routs)) <= 1e-5
    rg = torch.autograd.grad(sum((o * go.double().cpu()).sum() for o, go in zip(routs, gouts)), [emb64] + p64)
    assert common.rel_err(grads[0].permute(0, 3, 1, 2), rg[0]) <= 1e-4
    worst = max(common.rel_err(a, b) for a, b in zip(grads[1:], rg[1:]))
    print("decoder bank at 96 x 96, B = 2: worst parameter-gradient error", worst)
    assert worst <= 1e-4


def test_module_forwards_build_graphs_like_the_reference_modules():
    """prompt_encoder(image, cues) / prompt_decoder[s](embedding3) called the reference's way."""
    TD = common.package()
    enc, dec = TD.build_texture_diffuser(seed=0)
    enc, dec = enc.cuda().train(), dec.cuda().train()
    image, depth = common.synthetic_inputs(1, 96, seed=4)
    x, emb3 = enc(image.cuda(), depth.cuda())
    assert emb3.requires_grad and emb3.shape == (1, 24, 24, 24) and not x.requires_grad
    prompts = dec[1](emb3)
    assert len(prompts) == 4 and prompts[0].shape == (1, 128, 24, 24)
    sum(p.sum() for p in prompts).backward()
    assert enc.encoder2.stages[2][5].pwconv1.weight.grad is not None
    assert enc.encoder1.weight.grad is not None and enc.propagation_weight_regressor.reg.weight.grad is not None
    assert dec[1].decoder[3].decoder[4].weight.grad is not None
    assert dec[0].decoder[0].decoder[0].weight.grad is None and enc.adaptor.weight.grad is None


def test_drop_path_mask_scales_the_branch():
    """Stochastic depth (cod.py:1102,1116): with keep = 0 the block is the identity, with keep = s
    the branch is scaled by s (timm semantics: mask / keep_prob per sample)."""
    TD = common.package()
    from dgtd_b200.twig.ops.functions import train_func as TF
    blk = TD.convnext_Block(128, drop_path=0.5, layer_scale_init_value=1.0).cuda()
    x = torch.randn(3, 6, 5, 128, device="cuda")
    args = (blk.dwconv.weight, blk.dwconv.bias, blk.norm.weight, blk.norm.bias, blk.pwconv1.weight,
            blk.pwconv1.bias, blk.pwconv2.weight, blk.pwconv2.bias, blk.gamma)
    with torch.no_grad():
        full = TF.ConvNextBlockFn.apply(x, *args, None, 1e-6)
        keep = torch.tensor([0.0, 2.0, 1.0], device="cuda")
        y = TF.ConvNextBlockFn.apply(x, *args, keep, 1e-6)
    assert torch.equal(y[0], x[0])
    assert float((y[1] - (x[1] + 2.0 * (full[1] - x[1]))).abs().max()) < 1e-4
    assert float((y[2] - full[2]).abs().max()) < 1e-5


def test_bf16_tensor_core_training_gradients():
    """Trunk GEMMs (forward, dgrad, split-K wgrad) on tcgen05 with bf16 operands: gradients of all
    parameters vs float64 autograd of the oracle.  Stated tolerance (max|d|/max|ref| per tensor):
    median <= 8e-2, 95th percentile <= 1.2e-1, worst <= 2e-1, i.e. the level of the reference module
    itself under bf16 autocast (see the comment at the assertion)."""
    TD = common.package()
    S, B = 192, 2
    enc, dec = TD.build_texture_diffuser(seed=0)
    common.perturb_regressor_(enc)
    image, depth = common.synthetic_inputs(B, S, seed=5)
    grids = common.pvt_token_grids((S, S))
    g = torch.Generator().manual_seed(8)
    gout_e3 = torch.randn(B, 24, S // 4, S // 4, generator=g) * 1e-2
    gout_tok = [[torch.randn(B, grids[s][0] * grids[s][1], e, generator=g) * 1e-2 for _ in range(n)]
                for s, (e, n) in enumerate(zip(common.PVT_EMBED_DIMS, common.PVT_DEPTHS))]
    (_, r3, rtoks), ref = _oracle_grads(enc, dec, image, depth, gout_e3, gout_tok)
    enc, dec = enc.cuda().eval(), dec.cuda().eval()
    e1, e3, toks = TD.texture_prompts_train(enc, dec, image.cuda(), depth.cuda(), precision="bf16")
    assert common.rel_err(e3, r3) <= 3e-2
    loss = (e3 * gout_e3.cuda()).sum()
    for s in range(4):
        for i, t in enumerate(toks[s]):
            loss = loss + (t * gout_tok[s][i].cuda()).sum()
    loss.backward()
    errs = {}
    for prefix, mod in (("enc.", enc), ("dec.", dec)):
        for k, p in mod.named_parameters():
            r = ref.get(prefix + k)
            if r is None:
                continue
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
            errs[prefix + k] = common.rel_err(p.grad, r)
    v = sorted(errs.values())
    worst = max(errs.items(), key=lambda kv: kv[1])
    med, p95 = v[len(v) // 2], v[int(len(v) * 0.95)]
    print(f"bf16 training: gradient error median {med:.2e}, p95 {p95:.2e}, worst {worst}")
    # bf16 operand rounding accumulates through the 36 residual blocks (gamma = 1 at init).  The
    # reference itself under torch.autocast(bfloat16), same inputs (measured in the authoring container,
    # CPU): median 5.9e-2, p95 9.2e-2, max 1.25e-1.  Bound = about 1.3x those figures.
    assert med <= 8e-2 and p95 <= 1.2e-1 and worst[1] <= 2e-1, (med, p95, worst)


def test_bf16_training_survives_gradscaler_loss_scales():
    """The reference trains under fp16 AMP + GradScaler (config/sod.yml:57, AmpOptimWrapper): the loss -- hence
    every upstream gradient -- arrives multiplied by the loss scale (65536 initially).  This library maps autocast
    to its bf16 tensor-core path, whose fp32-accumulated wgrad / dgrad GEMMs have fp32's exponent range: gradients
    of the 2^16-scaled loss must be finite and equal 2^16 x the unscaled ones (a power-of-two scale commutes with
    every rounding; tolerance covers bf16 re-rounding of intermediates that cross a binade differently: none
    expected, 1e-3 allowed), so that GradScaler.unscale_ recovers them and never sees an inf."""
    TD = common.package()
    S, B = 96, 2
    enc, dec = TD.build_texture_diffuser(seed=0)
    common.perturb_regressor_(enc)
    enc, dec = enc.cuda().eval(), dec.cuda().eval()
    image, depth = common.synthetic_inputs(B, S, seed=9)
    grids = common.pvt_token_grids((S, S))
    g = torch.Generator().manual_seed(10)
    gout_e3 = (torch.randn(B, 24, S // 4, S // 4, generator=g) * 1e-2).cuda()
    gout_tok = [[(torch.randn(B, grids[s][0] * grids[s][1], e, generator=g) * 1e-2).cuda() for _ in range(n)]
                for s, (e, n) in enumerate(zip(common.PVT_EMBED_DIMS, common.PVT_DEPTHS))]
    params = [(k, p) for mod in (enc, dec) for k, p in mod.named_parameters() if p.requires_grad]

    def grads(scale):
        for _, p in params:
            p.grad = None
        e1, e3, toks = TD.texture_prompts_train(enc, dec, image.cuda(), depth.cuda(), precision="bf16")
        loss = (e3 * gout_e3).sum()
        for s in range(4):
            for i, t in enumerate(toks[s]):
                loss = loss + (t * gout_tok[s][i]).sum()
        (loss * scale).backward()
        return {k: (None if p.grad is None else p.grad.clone()) for k, p in params}

    base, scaled = grads(1.0), grads(65536.0)
    worst = ("", 0.0)
    for k, gb in base.items():
        gs = scaled[k]
        if gb is None:
            assert gs is None or float(gs.abs().max()) == 0.0, k
            continue
        assert torch.isfinite(gs).all(), k
        err = common.rel_err(gs / 65536.0, gb)
        if err > worst[1]:
            worst = (k, err)
    print("loss scale 65536: worst deviation from 65536 x unscaled gradients", worst)
    assert worst[1] <= 1e-3, worst


# ---- decoder bank: data-movement halves of the conv gradients ---------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ks,stride,off,h,oh", [(3, 1, -1, 12, 12), (4, 2, -1, 12, 6), (4, 4, 0, 16, 4), (4, 8, 2, 16, 2)])
def test_col2im_is_the_adjoint_of_the_conv_gather(ks, stride, off, h, oh):
    """<im2col(x), dcol> == <x, col2im(dcol)> with im2col written as plain torch indexing; bit-level
    determinism (gather form, no atomics)."""
    from dgtd_b200.twig.ops.functions import decoder_bank as DB
    g = torch.Generator().manual_seed(3)
    B, w, ow, C = 2, h, oh, 24
    x = torch.randn(B, h, w, C, generator=g, dtype=torch.float64)
    dcol = torch.randn(B * oh * ow, ks * ks * C, generator=g, dtype=torch.float64)
    # reference im2col (zero outside the map)
    xp = torch.zeros(B, h + 16, w + 16, C, dtype=torch.float64)
    xp[:, 8:8 + h, 8:8 + w] = x
    cols = []
    for ty in range(ks):
        for tx in range(ks):
            ys = torch.arange(oh) * stride + off + ty + 8
            xs = torch.arange(ow) * stride + off + tx + 8
            cols.append(xp[:, ys][:, :, xs])
    col = torch.stack(cols, 3).reshape(B * oh * ow, ks * ks * C)
    lhs = (col * dcol).sum()
    out = torch.empty(B, h, w, C, device="cuda")
    DB.col2im(dcol.float().cuda(), C, None, out, C, ks, stride, off, (oh, ow))
    rhs = (x * out.double().cpu()).sum()
    assert abs(lhs - rhs) <= 1e-5 * (abs(lhs) + 1.0)
    out2 = torch.empty_like(out)
    DB.col2im(dcol.float().cuda(), C, None, out2, C, ks, stride, off, (oh, ow))
    assert torch.equal(out, out2)
    # ReLU mask: zero where the forward activation was not positive
    mask = torch.randn(B, h, w, C, generator=g).cuda()
    out3 = torch.empty_like(out)
    DB.col2im(dcol.float().cuda(), C, mask, out3, C, ks, stride, off, (oh, ow))
    assert torch.equal(out3, torch.where(mask > 0, out, torch.zeros_like(out)))


@pytest.mark.gpu
@pytest.mark.parametrize("ks,stride,off,h,oh", [(3, 1, -1, 12, 12), (4, 2, -1, 12, 6), (4, 8, 2, 16, 2)])
def test_im2col_and_group_sum(ks, stride, off, h, oh):
    from dgtd_b200.twig.ops.functions import decoder_bank as DB
    g = torch.Generator().manual_seed(4)
    B, w, ow, D = 3, h, oh, 3
    x = torch.randn(B, h, w, 32 * D, generator=g).to(torch.bfloat16)
    xs_ = x[..., 32:64].float()
    xp = torch.zeros(B, h + 16, w + 16, 32)
    xp[:, 8:8 + h, 8:8 + w] = xs_
    cols = []
    for ty in range(ks):
        for tx in range(ks):
            ys = torch.arange(oh) * stride + off + ty + 8
            xx = torch.arange(ow) * stride + off + tx + 8
            cols.append(xp[:, ys][:, :, xx])
    col = torch.stack(cols, 3).reshape(B * oh * ow, ks * ks * 32)
    xc = x.cuda()
    got = DB.im2col(xc[..., 32:64], ks, stride, off, (oh, ow))
    assert torch.equal(got.float().cpu(), col)
    s = DB.group_sum(xc, D, 32, 24)
    want = x.float().view(-1, D, 32)[:, :, :24].sum(1)
    assert torch.allclose(s.cpu(), want, atol=1e-6)


@pytest.mark.gpu
def test_decoder_bank_falls_back_per_decoder_for_non_power_of_two_grids():
    """200^2 input: PVT grids 50/25/13/7 are not 2^k folds of the 50x50 decoder map -> per-decoder path;
    gradients still reach every decoder parameter."""
    TD = common.package()
    from dgtd_b200.twig.ops.functions import decoder_bank as DB
    assert not DB.bank_supported((50, 50), TD.pvt_token_grids((200, 200)))
    assert DB.bank_supported((96, 96), TD.pvt_token_grids((384, 384)))
    assert DB.bank_supported((88, 88), TD.pvt_token_grids((352, 352)))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_train_step_replays_the_eager_gradients(precision):
    """CUDA-graph replay of fwd+bwd (twig/graphs.py) gives the gradients of the eager step bit for bit
    (eval mode: no DropPath randomness; every reduction in the library has a fixed order), also after the
    inputs and the weights changed between replays."""
    TD = common.package()
    from dgtd_b200.twig import graphs
    enc, dec = TD.build_texture_diffuser(seed=0)
    enc, dec = enc.cuda().eval(), dec.cuda().eval()
    common.perturb_regressor_(enc)
    S, B = 192, 2
    image, depth = common.synthetic_inputs(B, S, seed=5)
    image2, depth2 = common.synthetic_inputs(B, S, seed=6)
    params = [p for p in list(enc.parameters()) + list(dec.parameters()) if p.requires_grad]

    def eager(img, dep):
        for p in params:
            p.grad = None
        graphs.default_loss(*TD.texture_prompts_train(enc, dec, img.cuda(), dep.cuda(), precision=precision)).backward()
        return [None if p.grad is None else p.grad.clone() for p in params]

    want1 = eager(image, depth)
    step = graphs.GraphedTrainStep(enc, dec, image.cuda(), depth.cuda(), precision=precision)
    step()
    got1 = [p.grad.clone() for p in params]
    with torch.no_grad():
        for p in params:
            p.mul_(1.01)                     # an "optimizer step": replays must read the new weights
    step(image2.cuda(), depth2.cuda())
    got2 = [p.grad.clone() for p in params]
    want2 = eager(image2, depth2)
    for want, got in ((want1, got1), (want2, got2)):
        for w, g in zip(want, got):
            if w is None:
                assert float(g.abs().max()) == 0.0
            else:
                assert torch.equal(w, g)

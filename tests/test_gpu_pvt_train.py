"""SURVEY.md 8f-1 in training: gradients of the PVT-v2 backbone that consumes the texture prompts (cod.py:824-1002,
1455-1509 under autograd) on the CUDA path vs float64 autograd of the CPU oracle (oracle/pvt_ref.py) on the same
seeded inputs -- the backward kernels alone, one Block, and every parameter of the backbone + hot path."""
import pytest
import torch
import torch.nn.functional as F

import common
from oracle import pvt_ref as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,N,Nk,heads", [(2, 100, 70, 2), (1, 576, 144, 5), (2, 16, 16, 8), (1, 300, 200, 1), (1, 2304, 144, 2)])
def test_attention_backward_kernel(B, N, Nk, heads, dtype):
    """Ragged query / key counts (N % 64 != 0, Nk % 64 != 0, several key tiles); bf16 inputs with fp32 math."""
    common.package()
    from dgtd_b200.twig.ops.functions import pvt_func as PF, pvt_train_func as PT
    g = torch.Generator().manual_seed(5)
    C = heads * 64
    q = torch.randn(B * N, C, generator=g).to(dtype)
    kv = torch.randn(B * Nk, 2 * C, generator=g).to(dtype)
    do = torch.randn(B * N, C, generator=g)
    qd = q.double().requires_grad_(True)
    kvd = kv.double().requires_grad_(True)
    qh = qd.view(B, N, heads, 64).permute(0, 2, 1, 3)
    kh = kvd.view(B, Nk, 2, heads, 64).permute(2, 0, 3, 1, 4)
    o = ((qh @ kh[0].transpose(-2, -1)) * 64 ** -0.5).softmax(-1) @ kh[1]
    o = o.transpose(1, 2).reshape(B * N, C)
    rq, rkv = torch.autograd.grad((o * do.double()).sum(), [qd, kvd])
    oc = PF.attention(q.cuda(), kv.cuda(), B, N, Nk, heads)
    dq, dkv = PT.attention_bwd(q.cuda(), kv.cuda(), oc, do.cuda(), B, N, Nk, heads)
    # bf16: the forward output the kernel reads for dO.O carries bf16 rounding
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel(dq, rq) <= tol and rel(dkv, rkv) <= tol, (rel(dq, rq), rel(dkv, rkv))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,C", [(2, 9, 13, 328), (3, 20, 37, 128), (1, 24, 24, 320), (2, 5, 3, 64)])
def test_dwconv3_gelu_backward_kernel(dtype, B, h, w, C):
    """328 channels: ragged strip (234 pixels) and channel tail of the pixel-strip kernel; C % 64 == 0 in bf16: the
    persistent TMA-staged kernel over several tiles per CTA (partial tiles in both directions, 1-5 channel groups)."""
    common.package()
    from dgtd_b200.twig.ops.functions import pvt_train_func as PT
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, h, w, C, generator=g).to(dtype)
    wt = torch.randn(C, 1, 3, 3, generator=g) * 0.3
    b = torch.randn(C, generator=g) * 0.1
    go = torch.randn(B, h, w, C, generator=g)
    xd = x.double().requires_grad_(True)
    wd, bd = wt.double().requires_grad_(True), b.double().requires_grad_(True)
    u = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=1, groups=C).permute(0, 2, 3, 1)
    y = 0.5 * u * (1.0 + torch.erf(u / 2 ** 0.5))
    rx, rw, rb = torch.autograd.grad((y * go.double()).sum(), [xd, wd, bd])
    wT = wt.reshape(C, 9).t().contiguous().cuda()
    dx, dwT, db = PT.dwconv3_gelu_bwd(x.cuda(), wT, b.cuda(), go.cuda())
    assert rel(dx, rx) <= 1e-5 and rel(dwT.t().reshape(C, 1, 3, 3), rw) <= 1e-5 and rel(db, rb) <= 1e-5


@pytest.fixture(scope="module")
def net():
    common.package()
    from dgtd_b200.twig.model import pvt
    m = pvt.pvt_v2_b2().eval()       # eval: DropPath off (the oracle has none); gradients still flow
    common.fill_params_(m, seed=0)
    return m.cuda()


@pytest.mark.parametrize("stage,hw", [(0, (16, 24)), (2, (8, 6)), (3, (4, 4))])
def test_block_gradients_match_oracle(net, stage, hw):
    """One Block (sr = 8 / 2 / 1): input, prompt and all parameter gradients at 1e-4 in fp32."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    set_precision(net, "fp32")
    H, W = hw
    C = P.EMBED_DIMS[stage]
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, H * W, C, generator=g)
    pr = torch.randn(2, H * W, C, generator=g) * 0.1
    go = torch.randn(2, H * W, C, generator=g)
    blk = getattr(net, f"block{stage + 1}")[1]
    sd = {k: v.detach().double().cpu().requires_grad_(True) for k, v in blk.state_dict().items()}
    xd, pd = x.double().requires_grad_(True), pr.double().requires_grad_(True)
    want = P.block(xd + pd, H, W, sd, P.NUM_HEADS[stage], P.SR_RATIOS[stage])
    ref = torch.autograd.grad((want * go.double()).sum(), [xd, pd] + list(sd.values()))
    for p in blk.parameters():
        p.grad = None
    xc, pc = x.cuda().requires_grad_(True), pr.cuda().requires_grad_(True)
    got = blk._forward_train(xc, pc, H, W, 0)
    assert rel(got, want) <= 1e-5
    (got * go.cuda()).sum().backward()
    assert rel(xc.grad, ref[0]) <= 1e-4 and rel(pc.grad, ref[1]) <= 1e-4
    named = dict(blk.named_parameters())
    for k, r in zip(sd, ref[2:]):
        assert rel(named[k].grad, r) <= 1e-4, (k, rel(named[k].grad, r))
    set_precision(net, None)


def _backbone_grads(net, image, depth, gouts):
    for p in net.parameters():
        p.grad = None
    e1, outs = net.forward_features(image.cuda(), depth.cuda())      # e1: high-pass of the image, no parameters
    loss = sum((o * go.cuda()).sum() for o, go in zip(outs, gouts))
    loss.backward()
    return e1, outs, {k: p.grad for k, p in net.named_parameters()}


def _backbone_case(S, B):
    image, depth = common.synthetic_inputs(B, S, seed=3)
    g = torch.Generator().manual_seed(17)
    gouts = [torch.randn(B, c, S // r, S // r, generator=g) * 1e-2 for c, r in zip(P.EMBED_DIMS, (4, 8, 16, 32))]
    return image, depth, gouts


def test_backbone_gradients_match_oracle(net):
    """128^2, B = 2: every parameter of `PyramidVisionTransformerImpr.forward_features` that the loss reaches (PVT
    blocks, patch embeds, stage norms AND the prompt encoder / decoders behind the prompts) at the north-star fp32
    tolerance 1e-4 against float64 autograd of the oracle."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    set_precision(net, "fp32")
    common.perturb_regressor_(net.prompt_encoder)
    S, B = 128, 2
    image, depth, gouts = _backbone_case(S, B)
    sd = {k: v.detach().double().cpu().requires_grad_(True) for k, v in net.state_dict().items()}
    r1, routs = P.forward_features(image.double(), depth.double(), sd)
    loss = sum((o * go.double()).sum() for o, go in zip(routs, gouts))
    names = list(sd)
    ref = dict(zip(names, torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)))
    e1, outs, got = _backbone_grads(net, image, depth, gouts)
    assert rel(e1, r1) <= 1e-4
    for o, r in zip(outs, routs):
        assert rel(o, r) <= 1e-4
    worst, n = ("", 0.0), 0
    for k, gr in got.items():
        r = ref.get(k)
        if r is None:
            assert gr is None or float(gr.abs().max()) == 0.0, k
            continue
        assert gr is not None, k
        n += 1
        e = rel(gr, r)
        if e > worst[1]:
            worst = (k, e)
    print("checked", n, "gradients; worst", worst)
    assert n >= 780
    assert worst[1] <= 1e-4, worst
    set_precision(net, None)


def test_backbone_gradients_bf16_track_fp32(net):
    """bf16 operands (tcgen05 forward / dgrad / wgrad, fp32 accumulate): gradient direction and size agree with the
    exact path (cosine >= 0.99 and norm ratio within 5 % for every large parameter tensor)."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    S, B = 128, 2
    image, depth, gouts = _backbone_case(S, B)
    set_precision(net, "fp32")
    _, _, exact = _backbone_grads(net, image, depth, gouts)
    exact = {k: v.clone() for k, v in exact.items() if v is not None}
    set_precision(net, "bf16")
    _, _, got = _backbone_grads(net, image, depth, gouts)
    set_precision(net, None)
    bad = []
    for k, r in exact.items():
        if r.numel() < 4096 or not k.startswith(("block", "patch_embed")):
            continue
        gq = got[k].double().flatten()
        rr = r.double().flatten()
        cos = float((gq @ rr) / (gq.norm() * rr.norm()).clamp_min(1e-300))
        ratio = float(gq.norm() / rr.norm().clamp_min(1e-300))
        if cos < 0.99 or abs(ratio - 1.0) > 0.05:
            bad.append((k, cos, ratio))
    assert not bad, bad[:8]

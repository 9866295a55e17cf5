"""the PVT-v2 backbone that consumes the texture prompts, on the CUDA path, against the
golden fixture from the unmodified reference (tests/golden/pvt_128.npz) and the float64 oracle."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import pvt_ref as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def net():
    common.package()
    from dgtd_b200.twig.model import pvt
    m = pvt.pvt_v2_b2().eval()
    common.fill_params_(m, seed=0)
    return m.cuda()


def test_state_dict_keys_match_the_fixture_recipe(net):
    assert len(net.state_dict()) == 786      # same key set as the reference's pvt_v2_b2 (checked at authoring time)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_features_matches_reference_golden(net, precision):
    common.package()
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    g = np.load(os.path.join(common.GOLDEN, "pvt_128.npz"))
    S, B = int(g["S"]), int(g["B"])
    image, depth = common.synthetic_inputs(B, S, seed=7)
    set_precision(net, precision)
    e1, outs = net.forward_features(image.cuda(), depth.cuda())
    assert [tuple(o.shape) for o in outs] == [(B, 64, S // 4, S // 4), (B, 128, S // 8, S // 8),
                                              (B, 320, S // 16, S // 16), (B, 512, S // 32, S // 32)]
    for s, o in enumerate(outs):
        ref = torch.from_numpy(g[f"out{s}"])
        got = o[:, ::4, ::2, ::2].cpu()
        # fp32: 1e-4; bf16: no worse than 2x the reference's own bf16-autocast error (min 2e-2)
        tol = 1e-4 if precision == "fp32" else max(2.0 * float(g["ref_bf16_relerr"][s]), 2e-2)
        assert rel(got, ref) <= tol, (s, rel(got, ref), tol)
        if precision == "fp32":
            assert np.allclose(common.moments(o.cpu()), g[f"out{s}_moments"], rtol=1e-4, atol=1e-7)
    set_precision(net, None)


def test_block_modules_against_the_oracle(net):
    """Block / Attention / Mlp / OverlapPatchEmbed forwards (reference signatures) vs the float64 restatement,
    stage 1 (sr = 8, one head) and stage 3 (sr = 2, five heads); non-square token grid."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    set_precision(net, "fp32")
    sd = {k: v.detach().double().cpu() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    for s, (H, W) in ((0, (16, 24)), (2, (8, 6))):
        C = P.EMBED_DIMS[s]
        x = torch.randn(2, H * W, C, generator=g)
        blk = getattr(net, f"block{s + 1}")[1]
        want = P.block(x.double(), H, W, P.sub(sd, f"block{s + 1}.1"), P.NUM_HEADS[s], P.SR_RATIOS[s])
        got = blk(x.cuda(), H, W)
        assert rel(got.cpu(), want) <= 1e-5, (s, rel(got.cpu(), want))
        a = blk.attn(x.cuda(), H, W)
        assert rel(a.cpu(), P.attention(x.double(), H, W, P.sub(sd, f"block{s + 1}.1.attn"), P.NUM_HEADS[s], P.SR_RATIOS[s])) <= 1e-5
        m = blk.mlp(x.cuda(), H, W)
        assert rel(m.cpu(), P.mlp(x.double(), H, W, P.sub(sd, f"block{s + 1}.1.mlp"))) <= 1e-5
    img = torch.randn(1, 3, 40, 56, generator=g)
    t, H, W = net.patch_embed1(img.cuda())
    want, Hr, Wr = P.overlap_patch_embed(img.double(), P.sub(sd, "patch_embed1"), 7, 4)
    assert (H, W) == (Hr, Wr) and rel(t.cpu(), want) <= 1e-5
    set_precision(net, None)


def test_attention_kernel_key_tails_and_bf16():
    """Nk not a multiple of the 64-key tile, N not a multiple of the 32-query block, 3 heads; bf16 inputs."""
    common.package()
    from dgtd_b200.twig.ops.functions import pvt_func as PF
    g = torch.Generator().manual_seed(12)
    B, N, Nk, heads = 2, 75, 144, 3
    q = torch.randn(B * N, heads * 64, generator=g)
    kv = torch.randn(B * Nk, 2 * heads * 64, generator=g)

    def ref(q, kv):
        qh = q.double().view(B, N, heads, 64).permute(0, 2, 1, 3)
        k = kv.double()[:, :heads * 64].reshape(B, Nk, heads, 64).permute(0, 2, 1, 3)
        v = kv.double()[:, heads * 64:].reshape(B, Nk, heads, 64).permute(0, 2, 1, 3)
        o = torch.softmax(qh @ k.transpose(-1, -2) / 8.0, -1) @ v
        return o.permute(0, 2, 1, 3).reshape(B * N, heads * 64)
    got = PF.attention(q.cuda(), kv.cuda(), B, N, Nk, heads)
    assert rel(got.cpu(), ref(q, kv)) <= 1e-5
    qb, kvb = q.to(torch.bfloat16), kv.to(torch.bfloat16)
    gotb = PF.attention(qb.cuda(), kvb.cuda(), B, N, Nk, heads)
    assert rel(gotb.float().cpu(), ref(qb.float(), kvb.float())) <= 8e-3      # one bf16 rounding of the output


# ---- shape sweeps of the new token operators against plain torch (float64) --------------------------------
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

_SET = dict(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))


@settings(**_SET)
@given(rows=st.integers(1, 70), c4=st.sampled_from([4, 16, 32, 80, 128, 160, 512]), with_add=st.booleans(),
       bf16_out=st.booleans())
def test_sweep_ln_tokens(rows, c4, with_add, bf16_out):
    common.package()
    from dgtd_b200.twig.ops.functions import pvt_func as PF
    from dgtd_b200.twig.ops.capi import BF16, F32
    C = c4 * 4 if c4 <= 128 else c4      # 16 .. 512 channels (and the 640-wide case below)
    g = torch.Generator().manual_seed(rows * 1000 + C)
    x = torch.randn(rows, C, generator=g) * 2 + 0.5
    add = torch.randn(rows, C, generator=g) if with_add else None
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    out, s = PF.ln_tokens(x.cuda(), w.cuda(), b.cuda(), 1e-6, BF16 if bf16_out else F32,
                          add=add.cuda() if with_add else None, want_sum=True)
    xs = x.double() + (add.double() if with_add else 0)
    ref = torch.nn.functional.layer_norm(xs, (C,), w.double(), b.double(), 1e-6)
    assert rel(s.cpu(), xs) <= 1e-6
    assert rel(out.float().cpu(), ref) <= (8e-3 if bf16_out else 1e-5)


@settings(**_SET)
@given(B=st.integers(1, 2), h=st.integers(1, 9), w=st.integers(1, 11), c8=st.sampled_from([1, 3, 8, 9, 40]))
def test_sweep_dwconv3_gelu(B, h, w, c8):
    common.package()
    from dgtd_b200.twig.ops.functions import pvt_func as PF
    C = 8 * c8
    g = torch.Generator().manual_seed(B * 7919 + h * 131 + w * 17 + C)
    x = torch.randn(B, h, w, C, generator=g)
    wt = torch.randn(C, 1, 3, 3, generator=g) * 0.4
    bias = torch.randn(C, generator=g)
    ref = torch.nn.functional.gelu(torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), wt.double(), bias.double(),
                                                              padding=1, groups=C)).permute(0, 2, 3, 1)
    wT = wt.reshape(C, 9).t().contiguous()
    got = PF.dwconv3_gelu(x.cuda(), wT.cuda(), bias.cuda())
    assert rel(got.cpu(), ref) <= 1e-5
    gotb = PF.dwconv3_gelu(x.to(torch.bfloat16).cuda(), wT.cuda(), bias.cuda())
    refb = torch.nn.functional.gelu(torch.nn.functional.conv2d(x.to(torch.bfloat16).double().permute(0, 3, 1, 2), wt.double(),
                                                               bias.double(), padding=1, groups=C)).permute(0, 2, 3, 1)
    assert rel(gotb.float().cpu(), refb) <= 8e-3


@pytest.mark.parametrize("B,h,w,C", [(3, 20, 37, 128), (2, 24, 24, 320), (1, 96, 96, 512), (5, 12, 12, 2048)])
def test_dwconv3_gelu_tma_tiles(B, h, w, C):
    """bf16 storage, C % 64 == 0: the persistent TMA-staged kernel (csrc/dwconv3_tma.cu) over several tiles per CTA,
    partial tiles in both directions, and the four stage shapes' channel counts."""
    common.package()
    from dgtd_b200.twig.ops.functions import pvt_func as PF
    g = torch.Generator().manual_seed(B + h + w + C)
    x = torch.randn(B, h, w, C, generator=g).to(torch.bfloat16)
    wt = torch.randn(C, 1, 3, 3, generator=g) * 0.4
    bias = torch.randn(C, generator=g)
    ref = torch.nn.functional.gelu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=1,
                                                              groups=C)).permute(0, 2, 3, 1)
    got = PF.dwconv3_gelu(x.cuda(), wt.reshape(C, 9).t().contiguous().cuda(), bias.cuda())
    assert rel(got.float().cpu(), ref) <= 8e-3


@settings(**_SET)
@given(B=st.integers(1, 2), N=st.integers(1, 130), Nk=st.integers(1, 150), heads=st.integers(1, 3))
def test_sweep_attention_fp32_and_bf16(B, N, Nk, heads):
    common.package()
    from dgtd_b200.twig.ops.functions import pvt_func as PF
    g = torch.Generator().manual_seed(B + 10 * N + 1000 * Nk + heads)
    q = torch.randn(B * N, heads * 64, generator=g)
    kv = torch.randn(B * Nk, 2 * heads * 64, generator=g)

    def ref(q, kv):
        qh = q.double().view(B, N, heads, 64).permute(0, 2, 1, 3)
        k = kv.double()[:, :heads * 64].reshape(B, Nk, heads, 64).permute(0, 2, 1, 3)
        v = kv.double()[:, heads * 64:].reshape(B, Nk, heads, 64).permute(0, 2, 1, 3)
        return (torch.softmax(qh @ k.transpose(-1, -2) / 8.0, -1) @ v).permute(0, 2, 1, 3).reshape(B * N, heads * 64)
    assert rel(PF.attention(q.cuda(), kv.cuda(), B, N, Nk, heads).cpu(), ref(q, kv)) <= 1e-5
    qb, kvb = q.to(torch.bfloat16), kv.to(torch.bfloat16)
    assert rel(PF.attention(qb.cuda(), kvb.cuda(), B, N, Nk, heads).float().cpu(), ref(qb.float(), kvb.float())) <= 1e-2


def test_backbone_at_352_and_batch_consistency(net):
    """352^2 (token grids 88/44/22/11, 121 keys: not a multiple of the 48-key tile) in fp32 vs the oracle; bf16
    batch of 3 == the same images one by one (no cross-sample coupling)."""
    from dgtd_b200.twig.model.texture_diffuser import set_precision
    set_precision(net, "fp32")
    sd = {k: v.detach().double().cpu() for k, v in net.state_dict().items()}
    image, depth = common.synthetic_inputs(1, 352, seed=21)
    with torch.no_grad():
        _, want = P.forward_features(image.double(), depth.double(), sd)
    _, outs = net.forward_features(image.cuda(), depth.cuda())
    for s in range(4):
        assert rel(outs[s].cpu(), want[s]) <= 1e-4, (s, rel(outs[s].cpu(), want[s]))
    set_precision(net, "bf16")
    imgs = [common.synthetic_inputs(1, 128, seed=30 + i) for i in range(3)]
    image = torch.cat([a for a, _ in imgs]).cuda()
    depth = torch.cat([d for _, d in imgs]).cuda()
    _, batch = net.forward_features(image, depth)
    for i in range(3):
        _, one = net.forward_features(image[i:i + 1], depth[i:i + 1])
        for s in range(4):
            assert rel(batch[s][i:i + 1].cpu(), one[s].cpu()) <= 2e-2
    set_precision(net, None)

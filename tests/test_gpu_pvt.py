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

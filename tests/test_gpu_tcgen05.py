"""tcgen05 / TMEM / TMA GEMM (bf16 operands, fp32 accumulation) through `dgtd_linear_fwd`.
Reference = float64 product of the SAME bf16-rounded operands, so only accumulation order and
the output rounding differ: fp32 outputs must agree to 1e-4, bf16 outputs to one bf16 ulp."""
import pytest
import torch

import common
from oracle import texture_diffuser_ref as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def OP():
    common.package()
    from dgtd_b200.twig.ops.functions import texture_diffusion_func
    return texture_diffusion_func


def operands(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    return a, w, b


def rel(got, ref):
    return float((got.double().cpu() - ref).abs().max() / ref.abs().max())


# (M, N, K): single tile, M/N/K tails, every BN variant (32/64/128/256), multi-wave persistent
SHAPES = [(128, 128, 64), (128, 128, 512), (200, 128, 128), (144, 24, 128), (384, 64, 192),
          (1000, 320, 288), (4096, 512, 128), (36992, 512, 128), (20000, 2048, 512), (9216, 256, 2048)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_bf16_fp32_out(OP, M, N, K):
    from dgtd_b200.twig.ops.capi import F32
    a, w, b = operands(M, N, K)
    ref = a.double() @ w.double().t() + b.double()
    got = OP.linear(a.cuda(), w.cuda(), b.cuda(), out_dtype=F32)
    assert rel(got, ref) <= 1e-4


@pytest.mark.parametrize("M,N,K", [(300, 512, 128), (5000, 2048, 512), (20000, 2048, 512), (19100, 1280, 320),
                                   (38000, 1024, 256)])   # the last three: more row blocks than CTA pairs
def test_linear_bf16_gelu_bf16_out(OP, M, N, K):
    a, w, b = operands(M, N, K, seed=1)
    ref = O.gelu_erf(a.double() @ w.double().t() + b.double())
    got = OP.linear(a.cuda(), w.cuda(), b.cuda(), act=1)
    assert got.dtype == torch.bfloat16
    # bf16 rounding (2^-9 relative) + 2.8e-5 absolute of the polynomial GELU
    err = (got.double().cpu() - ref).abs()
    assert float((err - ref.abs() * 2 ** -8).max()) <= 1e-3


def test_linear_residual_bf16(OP):
    g = torch.Generator().manual_seed(2)
    B, rows, N, K = 4, 600, 128, 512
    M = B * rows
    a, w, b = operands(M, N, K, seed=2)
    gam = torch.randn(N, generator=g)
    keep = torch.tensor([1.0, 0.0, 1.6, 1.6])
    res = torch.randn(M, N, generator=g)
    ref = res.double() + keep.double().repeat_interleave(rows)[:, None] * (gam.double() * (a.double() @ w.double().t() + b.double()))
    r = res.clone().cuda()
    OP.linear_residual_(a.cuda(), w.cuda(), b.cuda(), gam.cuda(), keep.cuda(), rows, r)
    assert rel(r, ref) <= 1e-4


def test_repeated_launches_are_deterministic(OP):
    from dgtd_b200.twig.ops.capi import F32
    a, w, b = operands(3000, 512, 256, seed=3)
    a, w, b = a.cuda(), w.cuda(), b.cuda()
    y0 = OP.linear(a, w, b, out_dtype=F32)
    for _ in range(5):
        assert torch.equal(OP.linear(a, w, b, out_dtype=F32), y0)


# ---- implicit-GEMM convolution (TMA im2col producer, 64B swizzle, grouped) ----------------------
CONV_CASES = [  # (B, h, w, groups, Cout, ks, stride, off, relu)
    (2, 16, 16, 3, 32, 3, 1, -1, True),      # decoder conv2 shape class (N=32 tile)
    (1, 96, 96, 1, 512, 3, 1, -1, True),     # conv1 of all 16 decoders (N=512)
    (2, 96, 96, 2, 64, 3, 1, -1, False),     # stage-1 prompts, no resize
    (2, 96, 96, 2, 128, 4, 2, -1, False),    # folded, ratio 2
    (1, 96, 96, 3, 320, 4, 4, 0, False),     # folded, ratio 4
    (3, 96, 96, 2, 512, 4, 8, 2, False),     # folded, ratio 8
    (1, 88, 88, 2, 64, 4, 2, -1, False),     # 352^2 input: 88 -> 44
    (1, 20, 28, 1, 40, 3, 1, -1, False),     # odd tile shapes / N tail
    (2, 24, 40, 5, 32, 3, 1, -1, True),      # five groups: the resident weight tiles are swapped four times
]


@pytest.mark.parametrize("B,h,w,G,Cout,ks,stride,off,relu", CONV_CASES)
def test_conv_nhwc_grouped_bf16(OP, B, h, w, G, Cout, ks, stride, off, relu):
    import torch.nn.functional as F
    from dgtd_b200.twig.ops.capi import ACT_NONE, ACT_RELU
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, h, w, 32 * G, generator=g).to(torch.bfloat16)
    x[..., 24::32] = 0                                   # a few zero pad channels like the real layout
    wt = (torch.randn(G, Cout, ks, ks, 32, generator=g) / (ks * 6.0)).to(torch.bfloat16)
    bias = torch.randn(G, Cout, generator=g)
    oh = (h - 1) // stride + 1 if ks == 3 else h // stride
    ow = (w - 1) // stride + 1 if ks == 3 else w // stride
    out = torch.empty(G, B, oh, ow, Cout, dtype=torch.float32).cuda()
    OP.conv_nhwc_grouped(x.cuda(), wt.reshape(G * Cout, ks * ks * 32).cuda().contiguous(), bias.reshape(-1).cuda(),
                         32, (oh, ow), ks, stride, off, ACT_RELU if relu else ACT_NONE, out, Cout, Cout, G, 32,
                         Cout, B * oh * ow * Cout)
    torch.cuda.synchronize()
    for gi in range(G):
        xi = x[..., 32 * gi:32 * gi + 32].double().permute(0, 3, 1, 2)
        wi = wt[gi].double().permute(0, 3, 1, 2)          # (Cout, 32, ks, ks)
        pad = 4 * stride + ks
        xp = F.pad(xi, (pad, pad, pad, pad))
        ref = F.conv2d(xp[:, :, pad + off:, pad + off:], wi, bias[gi].double(), stride=stride)[:, :, :oh, :ow]
        ref = ref.clamp_min(0) if relu else ref
        got = out[gi].permute(0, 3, 1, 2).double().cpu()
        assert rel(got, ref) <= 1e-4, (gi, rel(got, ref))
    # bf16 output (the decoder activations): TMA-store epilogue of the halo kernel; one output rounding
    outb = torch.full((G, B, oh, ow, Cout), float("nan"), dtype=torch.bfloat16).cuda()
    OP.conv_nhwc_grouped(x.cuda(), wt.reshape(G * Cout, ks * ks * 32).cuda().contiguous(), bias.reshape(-1).cuda(),
                         32, (oh, ow), ks, stride, off, ACT_RELU if relu else ACT_NONE, outb, Cout, Cout, G, 32,
                         Cout, B * oh * ow * Cout)
    assert torch.isfinite(outb.float()).all()
    assert rel(outb.float().cpu().double(), out.cpu().double()) <= 5e-3


@pytest.mark.parametrize("M,K", [(9216, 48), (300, 48), (5000, 128), (256, 64)])
def test_linear_with_row_layernorm_epilogue(OP, M, K):
    """The stem GEMM with its LayerNorm fused on the accumulator read-back (two epilogue warps per TMEM lane quadrant
    exchange their partial sums): equal to GEMM + two-pass LayerNorm of the same bf16 operands to 2e-5 of max|ref|,
    rows with a large common offset included (one-pass variance in fp32)."""
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(128, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(128, generator=g) + 3.0          # a common offset of ~3 sigma in every row
    lw, lb = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    y = a.double() @ w.double().t() + b.double()
    ref = torch.nn.functional.layer_norm(y, (128,), lw.double(), lb.double(), 1e-6)
    got = OP.linear_ln(a.cuda(), w.cuda(), b.cuda(), lw.cuda(), lb.cuda(), 1e-6)
    assert rel(got, ref) <= 2e-5


@pytest.mark.parametrize("M,N,K", [(4096, 24, 128), (1000, 24, 256), (36864, 24, 512), (300, 64, 1024), (129, 8, 36)])
def test_linear_tf32_reads_fp32_operands(OP, M, N, K):
    """Thin projections on kind::tf32 (the head of cod.py:1174 in bf16 mode): fp32 operands, TF32 products (10-bit
    mantissas, truncated by the tensor core), fp32 accumulation: within 2e-3 of the float64 product relative to max|ref|
    -- and exact when the operands are representable in TF32."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = a.double() @ w.double().t() + b.double()
    got = OP.linear_tf32(a.cuda(), w.cuda(), b.cuda())
    assert rel(got, ref) <= 2e-3
    a16, w16 = a.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()   # 8-bit mantissas: exact products in TF32
    ref16 = a16.double() @ w16.double().t() + b.double()
    assert rel(OP.linear_tf32(a16.cuda(), w16.cuda(), b.cuda()), ref16) <= 2e-6


def test_conv_nhwc_group_major_operands_match_interleaved(OP):
    """Round 2: the decoder bank keeps its hidden maps group-major (every group a dense (B,h,w,32) tensor).  conv1 with
    the chunk-scattered output, conv2 / the folded strided conv3 reading group-major input: bit-identical to the
    interleaved-slice form (same MMAs on the same values, only the addresses differ)."""
    from dgtd_b200.twig.ops.capi import ACT_NONE, ACT_RELU
    g = torch.Generator().manual_seed(11)
    B, h, w, D = 2, 24, 40, 5
    gs = B * h * w * 32
    emb = torch.randn(B, h, w, 32, generator=g).to(torch.bfloat16).cuda()
    w1 = (torch.randn(32 * D, 9 * 32, generator=g) / 18.0).to(torch.bfloat16).cuda()
    b1 = torch.randn(32 * D, generator=g).cuda()
    w2 = (torch.randn(32 * D, 9 * 32, generator=g) / 18.0).to(torch.bfloat16).cuda()
    b2 = torch.randn(32 * D, generator=g).cuda()
    E = 64
    w3 = (torch.randn(D * E, 16 * 32, generator=g) / 24.0).to(torch.bfloat16).cuda()
    b3 = torch.randn(D * E, generator=g).cuda()
    # interleaved slices
    h1 = torch.full((B, h, w, 32 * D), float("nan"), dtype=torch.bfloat16).cuda()
    OP.conv_nhwc_grouped(emb, w1, b1, 32, (h, w), 3, 1, -1, ACT_RELU, h1, 32 * D, 32 * D, 1, 0, 32 * D, 0)
    h2 = torch.full_like(h1, float("nan"))
    OP.conv_nhwc_grouped(h1, w2, b2, 32, (h, w), 3, 1, -1, ACT_RELU, h2, 32, 32 * D, D, 32, 32, 32)
    o = torch.full((D, B, (h // 2) * (w // 2), E), float("nan"), dtype=torch.bfloat16).cuda()
    OP.conv_nhwc_grouped(h2, w3, b3, 32, (h // 2, w // 2), 4, 2, -1, ACT_NONE, o, E, E, D, 32, E, B * (h // 2) * (w // 2) * E)
    # group-major
    g1 = torch.full((D, B, h, w, 32), float("nan"), dtype=torch.bfloat16).cuda()
    OP.conv_nhwc_grouped(emb, w1, b1, 32, (h, w), 3, 1, -1, ACT_RELU, g1, 32 * D, 32, 1, 0, 32 * D, gs)
    g2 = torch.full_like(g1, float("nan"))
    OP.conv_nhwc_grouped(g1[0], w2, b2, 32, (h, w), 3, 1, -1, ACT_RELU, g2, 32, 32, D, gs, 32, gs)
    og = torch.full_like(o, float("nan"))
    OP.conv_nhwc_grouped(g2[0], w3, b3, 32, (h // 2, w // 2), 4, 2, -1, ACT_NONE, og, E, E, D, gs, E,
                         B * (h // 2) * (w // 2) * E)
    torch.cuda.synchronize()
    assert torch.equal(g1.permute(1, 2, 3, 0, 4).reshape(B, h, w, 32 * D), h1)
    assert torch.equal(g2.permute(1, 2, 3, 0, 4).reshape(B, h, w, 32 * D), h2)
    assert torch.isfinite(og.float()).all() and torch.equal(og, o)


@pytest.mark.parametrize("Mo,No,Kr,lda,ldb,tr", [
    (512, 128, 4096, 512, 128, False),     # trunk shape (stage 0 pwconv)
    (256, 1024, 2304, 256, 1024, True),
    (288, 32, 2304, 288, 512, False),      # decoder conv2: 32-column slice of a 512-wide gradient
    (32, 288, 2304, 512, 288, False),      # ... and with the slice as the A operand (narrower than one TMA box)
    (512, 288, 1000, 512, 288, False),     # reduction length not a multiple of the 64-row k-block
    (64, 320, 72, 64, 320, False),         # tiny token grid (12x12 x B... ) and Mo below one pair tile
    (4096, 1024, 9216, 4096, 1024, False),
])
def test_wgrad_tc_mn_reads_untransposed_operands(Mo, No, Kr, lda, ldb, tr):
    """out = a^T b with a, b row-major activations consumed as MN-major tcgen05 operands; fp32 torch
    reference on the same bf16-rounded inputs: |err| <= 2e-3 * max|ref| (accumulation order only)."""
    from dgtd_b200.twig.ops.functions import train_func as TF
    g = torch.Generator().manual_seed(Mo + No + Kr)
    A = torch.randn(Kr, lda, generator=g).to(torch.bfloat16).cuda()
    Bm = torch.randn(Kr, ldb, generator=g).to(torch.bfloat16).cuda()
    off_a, off_b = (64 if lda > Mo else 0), (64 if ldb > No else 0)
    a, b = A[:, off_a:off_a + Mo], Bm[:, off_b:off_b + No]
    got = TF.wgrad_tc_mn(a, b, transpose_out=tr)
    ref = a.float().t() @ b.float()
    if tr:
        ref = ref.t()
    assert got.shape == ref.shape
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err <= 2e-3, err
    got2 = TF.wgrad_tc_mn(a, b, transpose_out=tr)
    assert torch.equal(got, got2)


@pytest.mark.parametrize("M,C,mean", [(600, 256, 0.0), (2304, 128, 0.7), (4096, 512, 2.0)])
def test_layernorm_folded_into_pwconv1(M, C, mean):
    """cod.py:1108-1110 as ONE GEMM: act(rstd * (y W'^T - mean * rowsum(W')) + (W1 ln_b + b1)) against float64
    GELU(LN(y) W1^T + b1) on the same bf16-rounded y; also against the two-pass form (LayerNorm kernel -> bf16 ->
    GEMM).  Rows with a channel mean up to 2 sigma: the folded form rounds y BEFORE centring, so its error grows with
    |mean| / sigma -- bounded here at 2x the two-pass error + 1e-2."""
    common.package()
    from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
    from dgtd_b200.twig.ops import capi
    g = torch.Generator().manual_seed(C + M)
    y = (torch.randn(M, C, generator=g) * 1.3 + mean + 0.3 * torch.randn(M, 1, generator=g)).to(torch.bfloat16)
    ln_w = 1.0 + 0.2 * torch.randn(C, generator=g)
    ln_b = 0.2 * torch.randn(C, generator=g)
    w1 = torch.randn(4 * C, C, generator=g) / C ** 0.5
    b1 = 0.1 * torch.randn(4 * C, generator=g)
    y64 = y.double()
    mu, var = y64.mean(1, keepdim=True), y64.var(1, unbiased=False, keepdim=True)
    a64 = (y64 - mu) / torch.sqrt(var + 1e-6) * ln_w.double() + ln_b.double()
    ref = torch.nn.functional.gelu(a64 @ w1.double().t() + b1.double())
    # folded
    wq = (w1 * ln_w[None, :]).to(torch.bfloat16)
    col_s = wq.float().sum(1)
    cbias = w1 @ ln_b + b1
    stats = torch.stack([mu.float().squeeze(1), (1.0 / torch.sqrt(var + 1e-6)).float().squeeze(1)], 1).contiguous()
    got = OP.linear_lnfold(y.cuda(), wq.cuda(), cbias.cuda(), col_s.cuda(), stats.cuda(), act=capi.ACT_GELU)
    e_fold = common.rel_err(got, ref)
    # two-pass: normalised activation rounded to bf16, plain GEMM
    two = OP.linear(a64.float().to(torch.bfloat16).cuda(), w1.to(torch.bfloat16).cuda(), b1.cuda(), act=capi.ACT_GELU)
    e_two = common.rel_err(two, ref)
    print(f"M={M} C={C} mean={mean}: folded {e_fold:.2e}, two-pass {e_two:.2e}")
    assert e_fold <= 2.0 * e_two + 1e-2


@pytest.mark.parametrize("M", [128, 1280, 148 * 128 * 2 + 384])
def test_convnext_mlp_fused_matches_the_two_gemms(M):
    """dgtd_convnext_mlp_fused_fwd (stage 0, C = 128): x + gamma * (GELU(LN(y) W1^T + b1) W2^T + b2) with the hidden tensor on
    chip, against (a) the two un-fused tcgen05 GEMMs (same bf16 rounding points: must agree to accumulation order) and
    (b) the float64 formula on the same bf16-rounded y.  One tile, a partial wave, and more tiles than SMs x 2 (both TMEM
    accumulators and every ring wrap)."""
    common.package()
    from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
    from dgtd_b200.twig.ops import capi
    C = 128
    g = torch.Generator().manual_seed(M)
    y = (torch.randn(M, C, generator=g) * 1.3 + 0.4 + 0.3 * torch.randn(M, 1, generator=g)).to(torch.bfloat16)
    ln_w, ln_b = 1.0 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    w1, b1 = torch.randn(4 * C, C, generator=g) / C ** 0.5, 0.1 * torch.randn(4 * C, generator=g)
    w2, b2 = torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5, 0.1 * torch.randn(C, generator=g)
    gamma = 0.5 + torch.rand(C, generator=g)
    x = torch.randn(M, C, generator=g)
    y64 = y.double()
    mu, var = y64.mean(1, keepdim=True), y64.var(1, unbiased=False, keepdim=True)
    a64 = (y64 - mu) / torch.sqrt(var + 1e-6) * ln_w.double() + ln_b.double()
    hid64 = torch.nn.functional.gelu(a64 @ w1.double().t() + b1.double())
    ref = x.double() + gamma.double() * (hid64 @ w2.double().t() + b2.double())
    wq = (w1 * ln_w[None, :]).to(torch.bfloat16).cuda()
    col_s, cbias = wq.float().sum(1), (w1 @ ln_b + b1).cuda()
    stats = torch.stack([mu.float().squeeze(1), (1.0 / torch.sqrt(var + 1e-6)).float().squeeze(1)], 1).contiguous().cuda()
    w2q = w2.to(torch.bfloat16).cuda()
    hid = OP.linear_lnfold(y.cuda(), wq, cbias, col_s, stats, act=capi.ACT_GELU)
    two = OP.linear_residual_(hid, w2q, b2.cuda(), gamma.cuda(), None, M, x.cuda().clone())
    got = OP.convnext_mlp_fused_(y.cuda(), stats, wq, col_s, cbias, w2q, b2.cuda(), gamma.cuda(), x.cuda().clone())
    torch.cuda.synchronize()
    delta = common.rel_err(got - x.cuda(), (two - x.cuda()).double().cpu())      # the update, not the residual stream
    e_f, e_t = common.rel_err(got - x.cuda(), ref - x.double()), common.rel_err(two - x.cuda(), ref - x.double())
    print(f"M={M}: fused vs two-GEMM {delta:.2e}; vs float64: fused {e_f:.2e}, two-GEMM {e_t:.2e}")
    assert delta <= 1e-5 and e_f <= 1.05 * e_t + 1e-4
    # in place on the residual stream, and gamma = None
    xin = x.cuda().clone()
    out = OP.convnext_mlp_fused_(y.cuda(), stats, wq, col_s, cbias, w2q, b2.cuda(), None, xin)
    assert out.data_ptr() == xin.data_ptr()
    two1 = OP.linear_residual_(hid, w2q, b2.cuda(), None, None, M, x.cuda().clone())
    assert common.rel_err(out, two1.double().cpu()) <= 1e-5


def test_dwconv7_stats_matches_conv_and_row_moments():
    """dgtd_dwconv7_stats_tma_fwd: y = depthwise 7x7 (bf16 store), stats = (mean, rstd) of the STORED row."""
    common.package()
    from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
    g = torch.Generator().manual_seed(5)
    B, h, w, C = 2, 24, 40, 256
    x = torch.randn(B, h, w, C, generator=g)
    wt = torch.randn(C, 1, 7, 7, generator=g) * 0.2
    bias = torch.randn(C, generator=g) * 0.1
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), bias.double(), padding=3, groups=C)
    y, stats = OP.dwconv7_stats_tma(x.cuda(), wt.reshape(C, 49).t().contiguous().cuda(), bias.cuda(), 1e-6)
    assert y.dtype == torch.bfloat16
    assert common.rel_err(y.float().permute(0, 3, 1, 2), ref) <= 5e-3
    yf = y.double().cpu().view(-1, C)
    mu, var = yf.mean(1), yf.var(1, unbiased=False)
    assert common.rel_err(stats[:, 0], mu) <= 1e-5 and common.rel_err(stats[:, 1], 1.0 / torch.sqrt(var + 1e-6)) <= 1e-5

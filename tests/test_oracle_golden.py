"""The oracle (oracle/texture_diffuser_ref.py) against golden vectors produced by the unmodified
reference modules (tests/golden/make_golden.py).  CPU only; float64 => agreement to rounding."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import texture_diffuser_ref as O

TOL = 1e-11


def t(a):
    return torch.from_numpy(np.asarray(a)).double()


def close(a, b, tol=TOL):
    assert a.shape == b.shape
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert err <= tol, err


def test_surface_normals(golden_ops):
    close(O.surface_normals(t(golden_ops["normals_in"])), t(golden_ops["normals_out"]))


def test_fft_highpass_both_forms(golden_ops):
    x, ref = t(golden_ops["fft_in"]), t(golden_ops["fft_out"])
    close(O.fft_highpass(x), ref)
    close(O.fft_highpass_projector(x), ref, 1e-10)


def test_regressor(golden_ops):
    close(O.regress_weights(t(golden_ops["reg_in"]), t(golden_ops["reg_w"]), t(golden_ops["reg_b"])),
          t(golden_ops["reg_out"]))


@pytest.mark.parametrize("tag", ["mp24", "mp1"])
def test_message_passing(golden_ops, tag):
    x, w = t(golden_ops[f"{tag}_x"]), t(golden_ops[f"{tag}_w"])
    close(O.message_passing_core(x, w), t(golden_ops[f"{tag}_core"]))
    full = O.message_passing(x, w, t(golden_ops[f"{tag}_convw"]), t(golden_ops[f"{tag}_convb"]), (48, 48))
    close(full, t(golden_ops[f"{tag}_full"]))


@pytest.mark.parametrize("tag", ["mp24", "mp1"])
def test_message_passing_gradients(golden_ops, tag):
    """autograd through the oracle's explicit stencil == autograd through the reference's unfold."""
    x = t(golden_ops[f"{tag}_x"]).requires_grad_(True)
    w = t(golden_ops[f"{tag}_w"]).requires_grad_(True)
    gx, gw = torch.autograd.grad(O.message_passing_core(x, w), [x, w], t(golden_ops[f"{tag}_gout"]))
    close(gx, t(golden_ops[f"{tag}_gx"]), 1e-10)
    close(gw, t(golden_ops[f"{tag}_gw"]), 1e-10)


def test_layer_norm(golden_ops):
    g = golden_ops
    close(O.layer_norm_channels_first(t(g["ln_channels_first_in"]), t(g["ln_channels_first_w"]),
                                      t(g["ln_channels_first_b"])), t(g["ln_channels_first_out"]))
    close(O.layer_norm_channels_last(t(g["ln_channels_last_in"]), t(g["ln_channels_last_w"]),
                                     t(g["ln_channels_last_b"])), t(g["ln_channels_last_out"]))


def test_convnext_block(golden_ops):
    p = {k[len("blk_p_"):]: t(v) for k, v in golden_ops.items() if k.startswith("blk_p_")}
    close(O.convnext_block(t(golden_ops["blk_in"]), p), t(golden_ops["blk_out"]))


def test_decoder_and_injection(golden_ops):
    p = {k[len("dec_p_"):]: t(v) for k, v in golden_ops.items() if k.startswith("dec_p_")}
    y = O.shape_prop_decoder(t(golden_ops["dec_in"]), p)
    close(y, t(golden_ops["dec_out"]))
    for n in (8, 4, 2):
        close(O.prompt_to_tokens(y, (n, n)), t(golden_ops[f"dec_tokens{n}"]))


def test_seeded_parameters_match_reference_checksums():
    """torch.manual_seed(0) + our module construction == the reference's random init."""
    TD = common.package()
    enc, dec = TD.build_texture_diffuser(seed=0)
    fx = common.load_params_fixture()
    sd = {"prompt_encoder." + k: v for k, v in enc.state_dict().items()}
    sd.update({"prompt_decoder." + k: v for k, v in dec.state_dict().items()})
    assert list(sd) == list(fx)
    for k, v in sd.items():
        shape, s, sa = fx[k]
        assert list(v.shape) == shape, k
        assert abs(float(v.double().sum()) - s) <= 1e-6 * max(1.0, sa), k
        assert abs(float(v.double().abs().sum()) - sa) <= 1e-6 * max(1.0, sa), k


@pytest.mark.parametrize("name,S,w20", [("path_384", 384, False), ("path_384_w20", 384, True),
                                        ("path_352_w20", 352, True)])
def test_full_path_against_reference(name, S, w20):
    """Whole hot path (B=1) through the oracle vs the reference run recorded in the fixture."""
    TD = common.package()
    enc, dec = TD.build_texture_diffuser(seed=0)
    if w20:
        common.perturb_regressor_(enc)
    pe, pd = common.oracle_params(enc, dec)
    image, depth = common.synthetic_inputs(1, S)
    with torch.no_grad():
        e1, e3, toks = O.texture_prompts(image.double(), depth.double(), pe, pd)
    outs = common.flatten_outputs(e1, e3, toks)
    fx = np.load(os.path.join(common.GOLDEN, name + ".npz"))
    assert len(outs) == 18
    for k, v in outs.items():
        close(common.subsample(k, v), t(fx[k + ".sub"]), 1e-9)
        mom = common.moments(v)
        assert np.allclose(mom, fx[k + ".mom"], rtol=1e-9, atol=1e-12), k


def test_pvt_backbone_oracle_matches_reference_golden():
    """oracle/pvt_ref.py (with the texture-prompt oracle inside) on the mirror module's
    state dict == the unmodified reference `pvt_v2_b2.forward_features` fixture (float64, 1e-9)."""
    import os
    import numpy as np
    from oracle import pvt_ref as P
    common.package()
    from dgtd_b200.twig.model import pvt
    g = np.load(os.path.join(common.GOLDEN, "pvt_128.npz"))
    S, B = int(g["S"]), int(g["B"])
    net = pvt.pvt_v2_b2().eval()
    common.fill_params_(net, seed=0)
    sd = {k: v.detach().double() for k, v in net.state_dict().items()}
    image, depth = common.synthetic_inputs(B, S, seed=7)
    with torch.no_grad():
        _, outs = P.forward_features(image.double(), depth.double(), sd)
    for s, o in enumerate(outs):
        ref = torch.from_numpy(g[f"out{s}"])
        err = float((o[:, ::4, ::2, ::2] - ref).abs().max() / ref.abs().max())
        assert err <= 1e-9, (s, err)
        assert np.allclose(common.moments(o), g[f"out{s}_moments"], rtol=1e-9)


def test_structure_loss_oracle_matches_reference_golden():
    """oracle/loss_ref.py == `cod.cal_loss` fixture (value and autograd gradient, float64)."""
    import os
    import numpy as np
    from oracle import loss_ref as L
    g = np.load(os.path.join(common.GOLDEN, "loss_small.npz"))
    for tag, shape in (("a", (2, 48, 64)), ("b", (3, 40, 40))):
        preds, gts = common.loss_inputs(*shape, seed=ord(tag))
        p = preds.double().requires_grad_(True)
        loss = L.structure_loss(p, gts.double())
        (gr,) = torch.autograd.grad(loss, p)
        assert abs(float(loss.detach()) - float(g[f"{tag}_loss"])) <= 1e-12
        assert float((gr - torch.from_numpy(g[f"{tag}_grad"])).abs().max()) <= 1e-13


def test_ssim_constant_oracle_matches_reference_golden():
    """oracle/loss_ref.ssim_constant against the value recorded from the unmodified `SSIM` module (cod.py:316-351)."""
    import os
    import numpy as np
    import torch
    from oracle import loss_ref as L
    import common
    g = np.load(os.path.join(common.GOLDEN, "loss_small.npz"))
    v = L.ssim_constant(torch.from_numpy(g["ssim_emb"]), torch.from_numpy(g["ssim_img"]))
    assert abs(float(v) - float(g["ssim_value"])) < 1e-13


def test_hitnet_train_oracle_matches_reference_golden():
    """SURVEY.md 8f-2 in training: oracle/hitnet_ref.py with train=True (batch-statistics BatchNorm) + oracle/loss_ref.py
    on the mirror module's state dict == the loss and the gradients (norm + 8 samples of each of the 845 tensors) the
    UNMODIFIED reference `Hitnet` produced in train() (tests/golden/make_golden_hitnet_train.py), float64."""
    import os
    import numpy as np
    from oracle import hitnet_ref as H
    from oracle import loss_ref as L
    common.package()
    from dgtd_b200.twig.model import hitnet
    fx = np.load(os.path.join(common.GOLDEN, "hitnet_train_128.npz"))
    S, B = int(fx["S"]), int(fx["B"])
    net = hitnet.Hitnet()
    common.hitnet_fixture_params_(net, seed=0)
    common.perturb_regressor_(net.backbone.prompt_encoder)
    sd = {k: v.detach().double().requires_grad_("running" not in k) for k, v in net.state_dict().items()
          if v.dtype.is_floating_point}
    image, depth = common.synthetic_inputs(B, S, seed=7)
    _, label = common.loss_inputs(B, S, S, seed=11)
    _, P1, P2 = H.hitnet_forward(image.double(), depth.double(), sd, train=True)
    loss = L.deep_supervision_loss(P1, P2, label.double())
    assert abs(float(loss.detach()) - float(fx["loss"])) <= 1e-10 * abs(float(fx["loss"]))
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = dict(zip(names, torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)))
    slopes = [k for k in names if k.endswith("body.1.weight")]
    grads[slopes[0]] = sum(grads[k] for k in slopes)             # ONE shared nn.PReLU() (cod.py:686)
    n = 0
    for k in fx.files:
        if not k.startswith("g/"):
            continue
        flat = grads[k[2:]].flatten()
        step = max(1, flat.numel() // 8)
        have = torch.cat([flat.norm().reshape(1), flat[::step][:8]])
        want = torch.from_numpy(fx[k])
        assert float((have - want).abs().max() / want.abs().max()) <= 1e-8, k
        n += 1
    assert n == 845
    for k in [str(s) for s in fx["unused"]]:
        assert grads[k] is None or float(grads[k].abs().max()) < 1e-13, k

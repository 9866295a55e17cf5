"""Drop-in boundary checks that need no GPU: the C-ABI library exports every symbol the header
declares, the module mirror has the reference's class names / signatures / state_dict keys, and
the product path fails loudly (no CPU fallback) when asked to compute without CUDA."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import common

HEADER = os.path.join(common.ROOT, "include", "dgtd_ops.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgtd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    common.package()
    from dgtd_b200.twig.ops import capi
    assert os.path.exists(capi.LIB_PATH), "build the extension first (python __graft_entry__.py)"
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dgtd_ops.h but not exported"
    assert set(capi.SIGNATURES) == set(syms), set(capi.SIGNATURES) ^ set(syms)
    capi.load()
    assert capi.load().dgtd_version() == 100
    assert capi.launch_count() >= 0


def test_header_argument_counts_match_binding():
    common.package()
    from dgtd_b200.twig.ops import capi
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, argtypes in capi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(argtypes), (name, n, len(argtypes))


def test_state_dict_keys_and_constructor_signatures():
    TD = common.package()
    enc, dec = TD.build_texture_diffuser(seed=0)
    keys = list(enc.state_dict())
    assert len(keys) == 358 and len(dec.state_dict()) == 96
    for k in ("propagation_weight_regressor.reg.weight", "encoder1.weight", "message_passing.conv.weight",
              "adaptor.weight", "encoder2.downsample_layers.0.0.weight", "encoder2.stages.2.26.pwconv2.bias",
              "encoder2.stages.3.2.gamma", "encoder2.convs.3.weight", "encoder2.fusion_conv.bias"):
        assert k in keys, k
    assert "3.decoder.2.decoder.4.weight" in dec.state_dict()
    assert enc.state_dict()["propagation_weight_regressor.reg.weight"].shape == (1176, 3, 1, 1)
    assert dec.state_dict()["2.decoder.5.decoder.4.weight"].shape == (320, 24, 3, 3)
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(TD.prompt_encoder.__init__) == ["self", "latent_dim", "embed_dim", "depth", "fusion"]
    assert sig(TD.prompt_encoder.forward) == ["self", "image", "cues", "cross"]
    assert sig(TD.prompt_decoder.forward) == ["self", "embedding", "cross"]
    assert sig(TD.MessagePassing.__init__) == ["self", "latent_dim", "img_size", "k", "max_step", "sym_norm"]
    assert sig(TD.MessagePassing.forward) == ["self", "input", "weight"]
    assert sig(TD.convnext_Block.__init__) == ["self", "dim", "drop_path", "layer_scale_init_value"]
    assert sig(TD.LayerNorm.__init__) == ["self", "normalized_shape", "eps", "data_format"]
    assert sig(TD.ShapePropEncoder.__init__) == ["self", "in_channels", "out_dim"]
    assert sig(TD.ShapePropDecoder.__init__) == ["self", "out_dim", "latent_dim"]
    assert sig(TD.ShapePropWeightRegressor.__init__) == ["self", "in_channels", "latent_dim"]
    # drop-path schedule of the trunk (cod.py:1140-1150)
    rates = [b.drop_path.drop_prob if isinstance(b.drop_path, TD.DropPath) else 0.0
             for st in enc.encoder2.stages for b in st]
    assert len(rates) == 36 and rates[0] == 0.0 and abs(rates[-1] - 0.4) < 1e-6


def test_checkpoint_round_trip_with_reference_layout(tmp_path):
    TD = common.package()
    enc, dec = TD.build_texture_diffuser(seed=0)
    sd = {"hitnet.backbone.prompt_encoder." + k: v for k, v in enc.state_dict().items()}
    sd.update({"hitnet.backbone.prompt_decoder." + k: v for k, v in dec.state_dict().items()})
    path = tmp_path / "epoch_1.pth"
    torch.save({"state_dict": sd}, path)
    loaded = torch.load(path)["state_dict"]
    enc2, dec2 = TD.build_texture_diffuser(seed=1)
    pre_e, pre_d = "hitnet.backbone.prompt_encoder.", "hitnet.backbone.prompt_decoder."
    enc2.load_state_dict({k[len(pre_e):]: v for k, v in loaded.items() if k.startswith(pre_e)}, strict=True)
    dec2.load_state_dict({k[len(pre_d):]: v for k, v in loaded.items() if k.startswith(pre_d)}, strict=True)
    assert all(torch.equal(a, b) for a, b in zip(enc.state_dict().values(), enc2.state_dict().values()))


def test_token_grids_and_fold_parameters():
    TD = common.package()
    from dgtd_b200.twig.model.texture_diffuser import _fold_params
    assert TD.pvt_token_grids((384, 384)) == [(96, 96), (48, 48), (24, 24), (12, 12)]
    assert TD.pvt_token_grids((352, 352)) == [(88, 88), (44, 44), (22, 22), (11, 11)]
    assert TD.pvt_token_grids((768, 768)) == [(192, 192), (96, 96), (48, 48), (24, 24)]
    assert _fold_params((96, 96), (48, 48)) == (2, -1)
    assert _fold_params((96, 96), (24, 24)) == (4, 0)
    assert _fold_params((96, 96), (12, 12)) == (8, 2)
    assert _fold_params((96, 96), (96, 96)) is None and _fold_params((96, 96), (32, 32)) is None


def test_folded_conv_weights_equal_conv_then_bilinear():
    """The 4x4 stride-r fold checked on CPU with torch reference ops."""
    import torch.nn.functional as F
    from dgtd_b200.twig.model.texture_diffuser import _fold_conv3_bilinear, _fold_params
    g = torch.Generator().manual_seed(0)
    w3 = torch.randn(5, 24, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, 24, 16, 16, generator=g, dtype=torch.float64)
    w4 = _fold_conv3_bilinear(w3.float()).double().reshape(5, 4, 4, 24).permute(0, 3, 1, 2)
    for n in (8, 4, 2):
        r, off = _fold_params((16, 16), (n, n))
        ref = F.interpolate(F.conv2d(x, w3, padding=1), size=(n, n), mode="bilinear")
        xp = F.pad(x, (2, 2, 2, 2))
        got = F.conv2d(xp[:, :, 2 + off:, 2 + off:], w4, stride=r)[:, :, :n, :n]
        assert float((got - ref).abs().max()) < 1e-6   # fold computed in fp32


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    TD = common.package()
    enc, dec = TD.build_texture_diffuser(seed=0)
    image, depth = common.synthetic_inputs(1, 96)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        with torch.no_grad():
            enc(image, depth)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        TD.texture_prompts(enc, dec, image, depth)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        TD.MessagePassing(24)(torch.randn(1, 24, 12, 12), torch.rand(1, 1176, 12, 12))


def test_paramwise_options_follow_the_custom_keys_of_the_training_config():
    """config/sod.yml:62-76 through mmengine's rule (longest matching key wins)."""
    import common
    common.package()
    from dgtd_b200.twig.optim import SOD_CUSTOM_KEYS, paramwise_options
    f = lambda n: paramwise_options(n, 5e-4, 0.1, SOD_CUSTOM_KEYS)  # noqa: E731
    assert f("hitnet.backbone.prompt_encoder.encoder2.stages.2.5.pwconv1.weight") == (5e-4 * 0.02, 0.1)
    assert f("hitnet.backbone.prompt_encoder.encoder2.downsample_layers.1.1.weight") == (5e-4 * 0.02, 0.1)
    assert f("hitnet.backbone.prompt_encoder.message_passing.conv.weight") == (5e-4 * 0.2, 0.1)
    assert f("hitnet.backbone.block1.0.attn.q.weight") == (5e-4 * 0.2, 0.1)
    assert f("hitnet.decoder_level1.0.body.0.weight") == (5e-4, 0.1)
    assert paramwise_options("x", 1.0, 0.5, None) == (1.0, 0.5)

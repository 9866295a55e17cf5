"""Fused AdamW (twig/optim.py, csrc/optim_ops.cu) against torch.optim.AdamW -- the optimizer the reference's
config builds (config/sod.yml:56-76) -- on the same parameters, gradients and per-prefix options."""
import pytest
import torch

import common

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch_adamw():
    common.package()
    from dgtd_b200.twig.optim import FusedAdamW, paramwise_options
    g = torch.Generator().manual_seed(0)
    shapes = {"hitnet.backbone.prompt_encoder.encoder2.stages.0.w": (37, 129), "hitnet.backbone.block1.w": (5000,),
              "hitnet.head.w": (3, 7, 11), "hitnet.backbone.prompt_encoder.reg.bias": (1,), "hitnet.x": (4096 * 2 + 3,)}
    keys = {"hitnet.backbone": {"lr_mult": 0.2}, "hitnet.backbone.prompt_encoder.encoder2.stages.0": {"lr_mult": 0.02}}
    ours = {n: torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for n, s in shapes.items()}
    ref = {n: torch.nn.Parameter(p.detach().double().cpu().clone()) for n, p in ours.items()}
    groups = []
    for n, p in ref.items():
        lr, wd = paramwise_options(n, 5e-4, 0.1, keys)
        groups.append({"params": [p], "lr": lr, "weight_decay": wd})
    assert [round(gp["lr"] / 5e-4, 6) for gp in groups] == [0.02, 0.2, 1.0, 0.2, 1.0]
    topt = torch.optim.AdamW(groups, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    opt = FusedAdamW(ours.items(), lr=5e-4, weight_decay=0.1, custom_keys=keys)
    for step in range(4):
        for n in shapes:
            gr = torch.randn(shapes[n], generator=g) * (10.0 ** (step - 2))
            ours[n].grad.copy_(gr.cuda() * 2.0)              # grad_scale 0.5 below (the 1 / world_size of a sum reduce)
            ref[n].grad = gr.double()
        opt.step(grad_scale=0.5)
        topt.step()
        for n in shapes:
            a, b = ours[n].detach().double().cpu(), ref[n].detach()
            assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max()), (step, n)
    # parameters are views of one flat buffer, values preserved by the re-homing
    assert all(p.data_ptr() >= opt.flat_param.data_ptr() for p in ours.values())


def test_train_step_with_fused_adamw_changes_the_inference_result():
    """Graph replay -> fused AdamW -> the eval-mode forward sees the updated weights (version counters bumped)."""
    TD = common.package()
    from dgtd_b200.twig import graphs
    from dgtd_b200.twig.optim import SOD_CUSTOM_KEYS, FusedAdamW
    enc, dec = TD.build_texture_diffuser(seed=0)
    enc, dec = enc.cuda().train(), dec.cuda().train()
    image, depth = common.synthetic_inputs(1, 128, seed=3)
    image, depth = image.cuda(), depth.cuda()
    enc.message_passing.img_size = 128
    named = [("hitnet.backbone.prompt_encoder." + n, p) for n, p in enc.named_parameters()] + \
            [("hitnet.backbone.prompt_decoder." + n, p) for n, p in dec.named_parameters()]
    opt = FusedAdamW(named, lr=5e-4, weight_decay=0.1, custom_keys=SOD_CUSTOM_KEYS)
    lrs = {n: o[0] for n, o in zip(opt.names, opt.options)}
    assert abs(lrs["hitnet.backbone.prompt_encoder.encoder2.stages.2.5.pwconv1.weight"] - 1e-5) < 1e-12
    assert abs(lrs["hitnet.backbone.prompt_decoder.0.decoder.0.decoder.0.weight"] - 1e-4) < 1e-12
    step = graphs.GraphedTrainStep(enc, dec, image, depth, precision="fp32", flat_grad=opt.flat_grad)
    enc.eval(); dec.eval()
    with torch.no_grad():
        before = TD.texture_prompts(enc, dec, image, depth, precision="fp32")[1].clone()
    enc.train(); dec.train()
    l0 = float(step())
    opt.step()
    l1 = float(step())
    assert l0 == l0 and l1 == l1 and l0 != l1                 # finite, and the replay reads the updated parameters
    enc.eval(); dec.eval()
    with torch.no_grad():
        after = TD.texture_prompts(enc, dec, image, depth, precision="fp32")[1]
    assert float((after - before).abs().max()) > 0
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None

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


def test_flat_layout_is_aligned_and_lr_scale_follows_a_cosine_schedule():
    """(1) every parameter re-homed into the flat buffer starts on a 128-byte boundary whatever the sizes
    before it (a 1-element PReLU slope / 1-channel head bias must not shift later tensors off the float4 / TMA
    alignment the kernels assume); (2) `step(lr_scale=...)` reproduces torch.optim.AdamW under CosineAnnealingLR
    (config/sod.yml param_scheduler)."""
    common.package()
    from dgtd_b200.twig.optim import FusedAdamW
    g = torch.Generator().manual_seed(1)
    shapes = {"a.prelu": (1,), "b.weight": (24, 3, 3, 3), "c.bias": (1,), "d.weight": (129, 5), "e.bias": (3,)}
    ours = {n: torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for n, s in shapes.items()}
    ref = {n: torch.nn.Parameter(p.detach().double().cpu().clone()) for n, p in ours.items()}
    opt = FusedAdamW(ours.items(), lr=5e-4, weight_decay=0.1)
    for p in ours.values():
        assert p.data_ptr() % 128 == 0 and p.grad.data_ptr() % 128 == 0
    topt = torch.optim.AdamW(list(ref.values()), lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(topt, T_max=6, eta_min=5e-6)
    for step in range(6):
        for n in shapes:
            gr = torch.randn(shapes[n], generator=g)
            ours[n].grad.copy_(gr.cuda())
            ref[n].grad = gr.double()
        opt.step(lr_scale=sched.get_last_lr()[0] / 5e-4)
        topt.step()
        sched.step()
        for n in shapes:
            a, b = ours[n].detach().double().cpu(), ref[n].detach()
            assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max()), (step, n)


def test_graph_capture_guards_against_re_homed_parameters():
    """an optimizer built AFTER the graph capture would re-home p.data and leave the replay reading freed
    storage: the constructor refuses, and a replay after any other storage move raises instead of computing garbage."""
    TD = common.package()
    from dgtd_b200.twig import graphs
    from dgtd_b200.twig.optim import FusedAdamW
    enc, dec = TD.build_texture_diffuser(seed=0)
    enc, dec = enc.cuda().train(), dec.cuda().train()
    image, depth = common.synthetic_inputs(1, 96, seed=3)
    step = graphs.GraphedTrainStep(enc, dec, image.cuda(), depth.cuda(), precision="fp32", warmup=1)
    step()
    named = [("enc." + n, p) for n, p in enc.named_parameters()] + [("dec." + n, p) for n, p in dec.named_parameters()]
    with pytest.raises(RuntimeError, match="already captured"):
        FusedAdamW(named, flat_grad=step.flat_grad)
    p = enc.encoder1.weight
    p.data = p.data.clone()                         # any storage move
    with pytest.raises(RuntimeError, match="storage moved"):
        step()
    for q in list(enc.parameters()) + list(dec.parameters()):
        q.grad = None

"""Generate the committed golden fixtures by running the UNMODIFIED reference modules
(`/root/reference/twig/model/cod.py`, imported through `oracle/ref_loader.py`) on CPU.

    python tests/golden/make_golden.py          # authoring container only (needs /root/reference)

The reference ships no tests or golden vectors for this path, so these files are
what pins the oracle (`oracle/texture_diffuser_ref.py`) and, through it, the CUDA kernels.

Outputs (all small, float64 masters unless noted):
  params_seed0.json   per-tensor checksums of the seed-0 random-init parameters (454 tensors)
  ops_small.npz       inputs, parameters and reference outputs of each module at small sizes,
                      including autograd gradients of MessagePassing
  path_<S>[_w20].npz  full hot path, B=1, SxS, seed-0 weights: sub-sampled outputs + full-tensor
                      moments; `_w20` = regressor weights x20 and random bias (non-trivial
                      diffusion weights)
Everything a test needs besides these files is regenerated from seeds by `tests/common.py`.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_loader import load_reference  # noqa: E402
import common  # noqa: E402  (tests/common.py: seeded inputs / parameter recipes shared with the tests)

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def np64(t):
    return t.detach().double().cpu().numpy()


def build_reference(m, seed=0):
    torch.manual_seed(seed)
    pe = m.prompt_encoder(24, list(common.PVT_EMBED_DIMS), list(common.PVT_DEPTHS), True)
    pd = nn.Sequential(*[m.prompt_decoder(24, e, d, True)
                         for e, d in zip(common.PVT_EMBED_DIMS, common.PVT_DEPTHS)])
    init = m.PyramidVisionTransformerImpr._init_weights
    pe.apply(lambda mod: init(None, mod))
    pd.apply(lambda mod: init(None, mod))
    return pe, pd


def params_fixture(m):
    pe, pd = build_reference(m)
    sd = {"prompt_encoder." + k: v for k, v in pe.state_dict().items()}
    sd.update({"prompt_decoder." + k: v for k, v in pd.state_dict().items()})
    rec = {k: [list(v.shape), float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}
    with open(os.path.join(OUT, "params_seed0.json"), "w") as f:
        json.dump(rec, f, indent=0)
    print("params_seed0.json", len(rec))


def ops_small(m):
    g = torch.Generator().manual_seed(1234)
    R = {}

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).float()

    # a1 surface normals (cod.py:96-109)
    d = torch.rand(2, 1, 9, 11, generator=g).float()
    R["normals_in"] = np64(d)
    R["normals_out"] = np64(m.cod.compute_surface_normals(None, d.double()))

    # a2 fft high-pass (cod.py:1256-1271), non-square
    pe = m.prompt_encoder(24, [64, 128, 320, 512], [3, 4, 6, 3], True)
    x = rnd(2, 3, 24, 20)
    R["fft_in"] = np64(x)
    R["fft_out"] = np64(pe.fft(x.double(), 0.3))

    # a4 regressor (cod.py:1051-1060)
    reg = m.ShapePropWeightRegressor(3, 24)
    with torch.no_grad():
        reg.reg.weight.normal_(0, 1.0, generator=g)
        reg.reg.bias.normal_(0, 1.0, generator=g)
    xx = torch.rand(1, 3, 6, 6, generator=g).float()
    R["reg_w"], R["reg_b"], R["reg_in"] = np64(reg.reg.weight), np64(reg.reg.bias), np64(xx)
    wts = reg.double()(xx.double())
    R["reg_out"] = np64(wts)

    # a6 MessagePassing: core (wc = c and wc = 1), full module, and autograd gradients
    for tag, (n, c, h, w, wc) in {"mp24": (1, 24, 9, 8, 24), "mp1": (1, 8, 10, 14, 1)}.items():
        mp = m.MessagePassing(c, img_size=48).double()
        with torch.no_grad():
            mp.conv.weight.normal_(0, 0.5, generator=g)
            mp.conv.bias.normal_(0, 0.5, generator=g)
        xin = rnd(n, c, h, w).double().requires_grad_(True)
        wgt = torch.rand(n, wc * 49, h, w, generator=g).double().requires_grad_(True)
        # the diffusion core's output is the input of `mp.conv`: capture it with a forward hook
        core = {}
        hk = mp.conv.register_forward_hook(lambda mod, inp, o: core.__setitem__("x", inp[0]))
        out = mp(xin, wgt)
        hk.remove()
        gout = torch.randn(core["x"].shape, generator=g).double()
        gx, gw = torch.autograd.grad(core["x"], [xin, wgt], gout)
        R[f"{tag}_x"], R[f"{tag}_w"] = np64(xin), np64(wgt)
        R[f"{tag}_core"], R[f"{tag}_full"] = np64(core["x"]), np64(out)
        R[f"{tag}_convw"], R[f"{tag}_convb"] = np64(mp.conv.weight), np64(mp.conv.bias)
        R[f"{tag}_gout"], R[f"{tag}_gx"], R[f"{tag}_gw"] = np64(gout), np64(gx), np64(gw)

    # a12 LayerNorm both formats (cod.py:1025-1049)
    for fmt, shape in (("channels_first", (2, 32, 5, 7)), ("channels_last", (2, 5, 7, 32))):
        ln = m.LayerNorm(32, eps=1e-6, data_format=fmt).double()
        with torch.no_grad():
            ln.weight.normal_(1.0, 0.3, generator=g)
            ln.bias.normal_(0, 0.3, generator=g)
        xi = rnd(*shape, scale=2.0) + 0.7
        R[f"ln_{fmt}_in"], R[f"ln_{fmt}_w"], R[f"ln_{fmt}_b"] = np64(xi), np64(ln.weight), np64(ln.bias)
        R[f"ln_{fmt}_out"] = np64(ln(xi.double()))

    # a8 convnext_Block (cod.py:1082-1117), eval mode, gamma randomised
    blk = m.convnext_Block(32, drop_path=0.2, layer_scale_init_value=1.0).double().eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.normal_(0, 0.2, generator=g)
        blk.norm.weight.add_(1.0)
    xi = rnd(2, 32, 10, 12)
    R["blk_in"] = np64(xi)
    for k, v in blk.state_dict().items():
        R["blk_p_" + k] = np64(v)
    R["blk_out"] = np64(blk(xi.double()))

    # a10 ShapePropDecoder (cod.py:1210-1226) + a11 injection (cod.py:1471)
    dec = m.ShapePropDecoder(40, 24).double()
    with torch.no_grad():
        for p in dec.parameters():
            p.normal_(0, 0.1, generator=g)
    emb = rnd(1, 24, 16, 16)
    R["dec_in"] = np64(emb)
    for k, v in dec.state_dict().items():
        R["dec_p_" + k] = np64(v)
    y = dec(emb.double())
    R["dec_out"] = np64(y)
    R["dec_tokens8"] = np64(F.interpolate(y, size=(8, 8), mode="bilinear").flatten(2).permute(0, 2, 1))
    R["dec_tokens4"] = np64(F.interpolate(y, size=(4, 4), mode="bilinear").flatten(2).permute(0, 2, 1))
    R["dec_tokens2"] = np64(F.interpolate(y, size=(2, 2), mode="bilinear").flatten(2).permute(0, 2, 1))
    np.savez_compressed(os.path.join(OUT, "ops_small.npz"), **R)
    print("ops_small.npz", len(R))


def full_path(m, S, w20):
    pe, pd = build_reference(m)
    if w20:
        common.perturb_regressor_(pe)
    pe.message_passing.img_size = S          # the reference hard-codes 384 (cod.py:1252)
    image, depth = common.synthetic_inputs(1, S)
    rec = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        pe_, pd_ = pe.to(dt).eval(), pd.to(dt).eval()
        with torch.no_grad():
            e1, e3 = pe_(image.to(dt), depth.to(dt))
            grids = common.pvt_token_grids((S, S))
            toks = []
            for s in range(4):
                ps = pd_[s](e3)
                toks.append([F.interpolate(p, size=grids[s], mode="bilinear").flatten(2).permute(0, 2, 1)
                             for p in ps])
        outs = common.flatten_outputs(e1, e3, toks)
        if tag == "f64":
            master = outs
            for k, v in outs.items():
                rec[k + ".sub"] = np64(common.subsample(k, v))
                rec[k + ".mom"] = common.moments(v)
        else:   # how far the reference's own fp32 run is from the fp64 master (context for tolerances)
            rec["ref_f32_relerr"] = np.array(
                [float((outs[k].double() - master[k]).abs().max() / master[k].abs().max()) for k in sorted(outs)])
            rec["ref_f32_keys"] = np.array(sorted(outs))
    # the reference under bf16 autocast (what `AmpOptimWrapper` training / a bf16 deployment of the
    # stock model computes): its distance to the fp64 master sets the bf16 tolerance of the kernels
    pe_, pd_ = pe.to(torch.float32).eval(), pd.to(torch.float32).eval()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        e1, e3 = pe_(image, depth)
        toks = [[F.interpolate(p, size=grids[s], mode="bilinear").flatten(2).permute(0, 2, 1)
                 for p in pd_[s](e3)] for s in range(4)]
    outs = common.flatten_outputs(e1, e3, toks)
    rec["ref_bf16_relerr"] = np.array(
        [float((outs[k].double() - master[k]).abs().max() / master[k].abs().max()) for k in sorted(outs)])
    name = f"path_{S}{'_w20' if w20 else ''}.npz"
    np.savez_compressed(os.path.join(OUT, name), **rec)
    print(name, "max ref fp32 relerr", rec["ref_f32_relerr"].max(), "ref bf16-autocast relerr",
          rec["ref_bf16_relerr"].min(), rec["ref_bf16_relerr"].max())


def main():
    m = load_reference()
    params_fixture(m)
    ops_small(m)
    full_path(m, 384, False)
    full_path(m, 384, True)
    full_path(m, 352, True)


if __name__ == "__main__":
    main()

"""Golden fixture of the PVT-v2 backbone with the texture prompts, from the UNMODIFIED
reference classes (`pvt_v2_b2`, cod.py:1782) run on CPU in float64.

    python tests/golden/make_golden_pvt.py          # authoring container only (needs /root/reference)

Parameters come from `common.fill_params_` (seeded per tensor name), inputs from `common.synthetic_inputs`.
Also asserts that the restatement `oracle/pvt_ref.py` equals the reference to 1e-10 before writing.
Output: pvt_<S>.npz with sub-sampled stage outputs + moments, and the float32 / bf16-autocast errors of
the reference against its own float64 result (the tolerances of the GPU tests are tied to those).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_loader import load_reference  # noqa: E402
from oracle import pvt_ref as P  # noqa: E402
import common  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main(S=128, B=1):
    m = load_reference()
    torch.manual_seed(0)
    net = m.pvt_v2_b2().eval()
    net.prompt_encoder.message_passing.img_size = S          # the reference hard-codes 384
    common.fill_params_(net, seed=0)
    image, depth = common.synthetic_inputs(B, S, seed=7)
    with torch.no_grad():
        net64 = net.double()
        e1, outs = net64.forward_features(image.double(), depth.double())
        params = {k: v.detach().double() for k, v in net64.state_dict().items()}
        oe1, oouts = P.forward_features(image.double(), depth.double(), params)
        assert rel(oe1, e1) < 1e-10, rel(oe1, e1)
        for a, b in zip(oouts, outs):
            assert rel(a, b) < 1e-10, rel(a, b)
        net32 = net.float()
        _, outs32 = net32.forward_features(image, depth)
        f32_err = [rel(a.double(), b) for a, b in zip(outs32, outs)]
        with torch.autocast("cpu", dtype=torch.bfloat16):
            _, outs16 = net32.forward_features(image, depth)
        bf16_err = [rel(a.double(), b) for a, b in zip(outs16, outs)]
    rec = {"S": np.array(S), "B": np.array(B), "ref_f32_relerr": np.array(f32_err), "ref_bf16_relerr": np.array(bf16_err)}
    for s, o in enumerate(outs):
        rec[f"out{s}"] = o[:, ::4, ::2, ::2].numpy()
        rec[f"out{s}_moments"] = common.moments(o)
    np.savez_compressed(os.path.join(OUT, f"pvt_{S}.npz"), **rec)
    print("reference fp32 err", f32_err, "bf16 autocast err", bf16_err)
    print("wrote", f"pvt_{S}.npz", os.path.getsize(os.path.join(OUT, f"pvt_{S}.npz")), "bytes")


if __name__ == "__main__":
    main()

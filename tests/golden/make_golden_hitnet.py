"""Golden fixture of the Hitnet decoder + predict head, from the UNMODIFIED reference class
(`Hitnet`, cod.py:685-807) run on CPU in float64 in eval mode.

    python tests/golden/make_golden_hitnet.py       # authoring container only (needs /root/reference)

Parameters come from `common.hitnet_fixture_params_` (seeded per tensor name; BatchNorm running statistics included),
inputs from `common.synthetic_inputs`.  Asserts that the restatement `oracle/hitnet_ref.py` equals the
reference to 1e-10 before writing.  Output: hitnet_<S>.npz with the four stage predictions, the SAM prediction,
the predict logits `P1[-1] + P2` (cod.py:149), their binarised masks at 0.5 and at `binary_thresh` 0.2 (packed
bits), and the errors / mask Hamming distances of the reference's own fp32 and bf16-autocast runs against its
float64 result (the tolerances of the GPU tests are tied to those).
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
from oracle import hitnet_ref as H  # noqa: E402
import common  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def masks(logits, thr):
    return (torch.sigmoid(logits) > thr).numpy()


def main(S=128, B=2):
    m = load_reference()
    torch.manual_seed(0)
    net = m.Hitnet().eval()
    net.backbone.prompt_encoder.message_passing.img_size = S      # the reference hard-codes 384
    common.hitnet_fixture_params_(net, seed=0)
    image, depth = common.synthetic_inputs(B, S, seed=7)
    with torch.no_grad():
        net64 = net.double()
        e1, P1, P2 = net64(image.double(), depth.double())
        logits = P1[-1] + P2
        params = {k: v.detach().double() for k, v in net64.state_dict().items() if v.dtype.is_floating_point}
        oe1, oP1, oP2 = H.hitnet_forward(image.double(), depth.double(), params)
        for a, b in zip(list(oP1) + [oP2], list(P1) + [P2]):
            assert rel(a, b) < 1e-10, rel(a, b)
        assert rel(H.predict_logits(oP1, oP2, (S, S)), logits) < 1e-10
        net32 = net.float()
        _, P1f, P2f = net32(image, depth)
        f32_err = rel((P1f[-1] + P2f).double(), logits)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            _, P1h, P2h = net32(image, depth)
        lh = (P1h[-1].float() + P2h.float()).double()
        bf16_err = rel(lh, logits)
    rec = {"S": np.array(S), "B": np.array(B), "ref_f32_relerr": np.array(f32_err),
           "ref_bf16_relerr": np.array(bf16_err)}
    sub = 2 if S <= 128 else 8
    for i, t in enumerate(P1):
        rec[f"P1_{i}"] = t[:, :, ::sub, ::sub].numpy()
    rec["P2"] = P2[:, :, ::sub, ::sub].numpy()
    rec["sub"] = np.array(sub)
    rec["logits"] = logits.numpy().astype(np.float64)
    for thr, key in ((0.5, "mask50"), (0.2, "mask20")):
        rec[key] = np.packbits(masks(logits, thr))
        rec[f"ref_f32_hamming_{key}"] = np.array(int((masks((P1f[-1] + P2f).double(), thr) != masks(logits, thr)).sum()))
        rec[f"ref_bf16_hamming_{key}"] = np.array(int((masks(lh, thr) != masks(logits, thr)).sum()))
    # decision margin: how close the float64 logits come to the two thresholds (a mask bit can only flip
    # legitimately when the logit error exceeds this)
    rec["margin50"] = np.array(float(logits.abs().min()))
    rec["margin20"] = np.array(float((logits - np.log(0.2 / 0.8)).abs().min()))
    path = os.path.join(OUT, f"hitnet_{S}.npz")
    np.savez_compressed(path, **rec)
    print("reference fp32 err", f32_err, "bf16 autocast err", bf16_err)
    print({k: int(v) for k, v in rec.items() if "hamming" in k}, "margins", float(rec["margin50"]), float(rec["margin20"]))
    print("logit range", float(logits.min()), float(logits.max()), "fg fraction", float(masks(logits, 0.5).mean()))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main(S=352, B=1)        # BASELINE configs[0]: COD forward, batch 1, 352 x 352

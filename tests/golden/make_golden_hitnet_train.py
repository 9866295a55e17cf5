"""Golden fixture of the TRAINING step of the whole model (SURVEY.md 8f-2 / 8f-1 under autograd): loss and the
gradient of every parameter of the UNMODIFIED reference `Hitnet` (cod.py:685-807) run on CPU in float64 with the
decoder's BatchNorms in train() (batch statistics) and the deep-supervision structure loss of cod.py:135-141.

    python tests/golden/make_golden_hitnet_train.py       # authoring container only (needs /root/reference)

The backbone is put in eval() so that its DropPath is the identity (the only stochastic element; the backbone has no
BatchNorm); the decoder stays in train().  Asserts that the restatement (`oracle/hitnet_ref.py` with train=True +
`oracle/loss_ref.py`) reproduces the reference's loss, every gradient and the BatchNorm running-statistics update to
1e-9 before writing.  Output hitnet_train_128.npz: the loss, per parameter the gradient's L2 norm and 8 strided
samples, and the updated running statistics of two BatchNorms.
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
from oracle import loss_ref as L  # noqa: E402
import common  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main(S=128, B=2):
    m = load_reference()
    torch.manual_seed(0)
    net = m.Hitnet()
    net.backbone.prompt_encoder.message_passing.img_size = S      # the reference hard-codes 384
    common.hitnet_fixture_params_(net, seed=0)
    common.perturb_regressor_(net.backbone.prompt_encoder)
    net = net.double().train()
    net.backbone.eval()
    image, depth = common.synthetic_inputs(B, S, seed=7)
    _, label = common.loss_inputs(B, S, S, seed=11)
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
              for k, v in net.state_dict().items() if v.dtype.is_floating_point}
    before = {k: v.detach().clone() for k, v in net.state_dict().items() if "running" in k}
    e1, P1, P2 = net(image.double(), depth.double())
    loss = L.deep_supervision_loss(P1, P2, label.double())
    loss.backward()
    oe1, oP1, oP2 = H.hitnet_forward(image.double(), depth.double(), params, train=True)
    oloss = L.deep_supervision_loss(oP1, oP2, label.double())
    names = [k for k, v in params.items() if v.requires_grad]
    ograds = dict(zip(names, torch.autograd.grad(oloss, [params[k] for k in names], allow_unused=True)))
    assert abs(float(oloss) - float(loss)) < 1e-10 * abs(float(loss)), (float(oloss), float(loss))
    rec = {"S": np.array(S), "B": np.array(B), "loss": np.array(float(loss))}
    named = dict(net.named_parameters())          # shared parameters (the ONE PReLU slope, cod.py:686) appear once
    owner = {}
    for k, v in net.state_dict(keep_vars=True).items():
        owner.setdefault(id(v), k)
    shared = {}
    for k, v in net.state_dict(keep_vars=True).items():
        if owner[id(v)] != k:
            shared.setdefault(owner[id(v)], []).append(k)
    n = 0
    worst = 0.0
    unused = []
    for k in names:
        if k not in named:
            continue                               # an alias of a shared parameter: summed into its owner below
        g = named[k].grad
        og = ograds[k]
        for alias in shared.get(k, []):
            og = og + ograds[alias]
        # a bias in front of conv -> train-mode BatchNorm (norm4.bias -> Translayer4_1 ...) has an exactly zero
        # gradient (the batch mean removes it); both sides hold rounding noise there
        if g is None or float(g.abs().max()) < 1e-13:
            assert og is None or float(og.abs().max()) < 1e-13, k
            unused.append(k)
            continue
        e = rel(og, g)
        worst = max(worst, e)
        assert e < 1e-8, (k, e)
        flat = g.detach().flatten()
        step = max(1, flat.numel() // 8)
        rec["g/" + k] = np.concatenate([[float(flat.norm())], flat[::step][:8].numpy()])
        n += 1
    # running-statistics update of nn.BatchNorm2d in train(): momentum 0.1, unbiased variance
    after = {k: v.detach().clone() for k, v in net.state_dict().items() if "running" in k}
    for k in ("Translayer2_1.bn.running_mean", "Translayer2_1.bn.running_var", "conv4.bn.running_mean",
              "conv4.bn.running_var"):
        rec["before/" + k] = before[k].numpy()
        rec["after/" + k] = after[k].numpy()
    rec["unused"] = np.array(unused)
    rec["shared"] = np.array([f"{k}<-{','.join(v)}" for k, v in shared.items()])
    path = os.path.join(OUT, f"hitnet_train_{S}.npz")
    np.savez_compressed(path, **rec)
    print("loss", float(loss), "gradients pinned", n, "worst oracle-vs-reference", worst, "grad-less", len(unused))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Import the UNMODIFIED reference model file for golden-vector generation.

TEST INFRASTRUCTURE ONLY.  Nothing on the product path imports this module.

The reference (`/root/reference/twig/model/cod.py`) pulls in third-party packages that
are not installed in this image (timm, mmengine, nest, segment_anything, torchcam,
matplotlib, mmseg ...).  Only three of those symbols are *used* by the hot path
(`DropPath`, `to_2tuple`, `trunc_normal_`, cod.py:816); the rest are stubbed with empty
modules.  The reference also hard-codes `.cuda()` (cod.py:1259); on a CPU-only box that
call is shimmed to the identity.  The reference source itself is never copied or edited.

`/root/reference` exists only in the authoring container.  `oracle/make_ref.py` stages a
byte-identical, git-ignored copy of the ONE file under `oracle/_ref/` (it ships to the GPU box like
the built `.so`); this loader prefers that copy, so nothing reads `/root/reference` at run time.
Users: `tests/golden/make_golden*.py` (fixtures), `tests/test_dropin_reference.py` (the reference's own
`forward_features` with this repo's classes patched in), and the baseline legs of `bench.py`
(`--impl reference` on the host cores, `gpu_eager_reference` on the same B200).
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

import contextlib

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve_root() -> str:
    env = os.environ.get("DGTD_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(_HERE, "_ref")
    if os.path.isfile(os.path.join(staged, "twig", "model", "cod.py")):
        return staged
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()
_REF_FILE = os.path.join(REFERENCE_ROOT, "twig", "model", "cod.py")
_CACHE = {}


@contextlib.contextmanager
def on_cpu():
    """Run reference code on the host cores of a box that HAS a GPU: `prompt_encoder.fft` hard-codes
    `.cuda()` for its mask (cod.py:1259); inside this context that call is the identity."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def reference_available() -> bool:
    return os.path.isfile(_REF_FILE)


class _DropPath(nn.Module):
    """timm 0.6.13 `DropPath` semantics: per-sample Bernoulli keep mask / keep_prob."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []  # behave like a package so that sub-imports resolve
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _install_stubs() -> None:
    ident = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    class _Any:
        def __init__(self, *a, **k):
            pass

    _stub("timm", create_model=lambda *a, **k: None)
    _stub("timm.models")
    _stub("timm.models.resnet", Bottleneck=_Any)
    _stub("timm.models.layers", DropPath=_DropPath, to_2tuple=to_2tuple,
          trunc_normal_=torch.nn.init.trunc_normal_)
    _stub("timm.models.registry", register_model=lambda f: f)
    _stub("timm.models.vision_transformer", _cfg=lambda **k: dict(k))
    _stub("mmengine")
    _stub("mmengine.model", BaseModel=nn.Module, MMDistributedDataParallel=_Any)
    _stub("mmengine.hooks", Hook=object)
    _stub("nest", export=ident)
    _stub("transformers", AutoImageProcessor=_Any, DPTForDepthEstimation=_Any)
    _stub("segment_anything", sam_model_registry={})
    _stub("segment_anything.utils")
    _stub("segment_anything.utils.transforms", ResizeLongestSide=_Any)
    _stub("torchcam")
    _stub("torchcam.methods", CAM=_Any)
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("mmseg")
    for name in ("cv2", "tqdm"):
        try:
            __import__(name)
        except Exception:  # pragma: no cover
            _stub(name, tqdm=lambda x, *a, **k: x)


def load_reference():
    """Return the reference `cod` python module (cached)."""
    if "mod" in _CACHE:
        return _CACHE["mod"]
    if not reference_available():
        raise FileNotFoundError(_REF_FILE)
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k.split(".")[0] in ("timm", "mmengine", "nest", "transformers",
                                    "segment_anything", "torchcam", "matplotlib", "mmseg")}
    _install_stubs()
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # cod.py:1259 shim
    spec = importlib.util.spec_from_file_location("_dgtd_reference_cod", _REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # un-stub `transformers` (a real package in this image) for everyone else
    for k in [k for k in sys.modules if k.split(".")[0] == "transformers"]:
        del sys.modules[k]
    for k, v in saved.items():
        if v is not None and k.split(".")[0] == "transformers":
            sys.modules[k] = v
    _CACHE["mod"] = mod
    return mod


# ------------------------------------------------------------------------------------------------
# The reference's hot path as the reference itself runs it (cod.py:1394-1396 construction,
# :1399-1414 init, :1467-1505 minus the PVT blocks).  Used for fixtures and as the timed baseline.
PVT_EMBED_DIMS = (64, 128, 320, 512)
PVT_DEPTHS = (3, 4, 6, 3)


def build_reference_hot_path(m, seed: int = 0, img_size: int = 384):
    """`prompt_encoder(24, embed_dims, depths, True)` + 4 x `prompt_decoder` built and initialised exactly as
    `PyramidVisionTransformerImpr.__init__` does (cod.py:1394-1414)."""
    torch.manual_seed(seed)
    pe = m.prompt_encoder(24, list(PVT_EMBED_DIMS), list(PVT_DEPTHS), True)
    pd = nn.Sequential(*[m.prompt_decoder(24, e, d, True) for e, d in zip(PVT_EMBED_DIMS, PVT_DEPTHS)])
    init = m.PyramidVisionTransformerImpr._init_weights
    pe.apply(lambda mod: init(None, mod))
    pd.apply(lambda mod: init(None, mod))
    pe.message_passing.img_size = img_size      # the reference hard-codes 384 (cod.py:1252)
    return pe, pd


def reference_hot_path(pe, pd, image, depth, grids):
    """embedding1, embedding3 and the 16 injected prompt tensors in token layout -- the statements of
    `forward_features` (cod.py:1467-1505) with the PVT blocks left out."""
    import torch.nn.functional as F
    e1, e3 = pe(image, depth)
    toks = []
    for s in range(4):
        ps = pd[s](e3)
        toks.append([F.interpolate(p, size=grids[s], mode="bilinear").flatten(2).permute(0, 2, 1) for p in ps])
    return e1, e3, toks

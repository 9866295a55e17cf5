"""Stage the UNMODIFIED reference model file under ``oracle/_ref/`` so that it travels to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY -- nothing on the product path imports anything under ``oracle/``.

    python oracle/make_ref.py            # authoring container only (needs /root/reference)

The reference is a Python module (``/root/reference/twig/model/cod.py``); there is nothing to compile.
"Building" the reference arm therefore means placing a byte-identical copy of that ONE file at
``oracle/_ref/twig/model/cod.py``.  ``oracle/_ref/`` is git-ignored (the reference source never enters
this repository's history) but not gpurun-ignored, so -- like the built ``libdgtd_ops.so`` -- it ships
with the tree to the GPU box, where ``/root/reference`` does not exist.  ``oracle/ref_loader.py`` imports
it there with the third-party imports stubbed (timm / mmengine / nest ...), and ``bench.py`` times it:

  * ``bench.py --impl reference``      the unmodified module on the host cores (``cpu_baseline.kind =
                                       "reference"``)
  * ``gpu_eager_reference`` (our arm)  the same module in eager PyTorch on the same B200 (the bar of
                                       SURVEY.md 8(d) config 2 (iii) / BASELINE.md 4)

``__graft_entry__.build()`` calls :func:`stage` whenever ``/root/reference`` is present; a SHA-256 of the
staged file is written next to it so a stale or edited copy is detectable.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("DGTD_REFERENCE_SRC", "/root/reference")
REL = os.path.join("twig", "model", "cod.py")
DST_ROOT = os.path.join(HERE, "_ref")


def staged_path() -> str:
    return os.path.join(DST_ROOT, REL)


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(verbose: bool = False) -> str | None:
    """Copy the reference file (byte-identical) into ``oracle/_ref``; returns the staged path or None when the
    reference tree is not present (GPU box: the prebuilt copy is used as it came)."""
    src = os.path.join(SRC_ROOT, REL)
    dst = staged_path()
    if not os.path.isfile(src):
        return dst if os.path.isfile(dst) else None
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if not os.path.isfile(dst) or _sha(dst) != _sha(src):
        shutil.copyfile(src, dst)
    with open(os.path.join(DST_ROOT, "SHA256"), "w") as f:
        f.write(f"{_sha(dst)}  {REL}\n")
    if verbose:
        print(f"staged {src} -> {dst} ({_sha(dst)[:16]})")
    return dst


if __name__ == "__main__":
    p = stage(verbose=True)
    sys.exit(0 if p else 1)

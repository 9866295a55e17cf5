"""Import alias: ``import dgtd_b200`` loads the package that lives in the directory
``depth-guided-texture-diffusion-for-image-semantic-segmentation_b200/`` (a name Python's
import statement cannot spell because of the hyphens)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                     "depth-guided-texture-diffusion-for-image-semantic-segmentation_b200")
_spec = _ilu.spec_from_file_location("dgtd_b200", _os.path.join(_DIR, "__init__.py"),
                                     submodule_search_locations=[_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["dgtd_b200"] = _mod
_spec.loader.exec_module(_mod)

"""B200-native depth-guided texture-diffusion hot path (sm_100a).

Layout mirrors the reference tree for this path only:

* ``csrc/``            hand-written CUDA kernels + the C-ABI (``include/dgtd_ops.h``)
* ``twig/ops/``        ctypes binding of ``libdgtd_ops.so`` and the operator functions
                       (the reference's ``twig/ops/functions`` convention)
* ``twig/model/``      ``nn.Module`` mirror of ``twig/model/cod.py:1025-1323`` (same class
                       names, constructor/forward signatures and ``state_dict`` keys)

There is no CPU or PyTorch fallback: every forward raises if the extension is missing.
"""
__version__ = "0.1.0"

"""`twig/ops`: the native operator extension of the texture-diffusion hot path.

* ``libdgtd_ops.so``   C-ABI library (``include/dgtd_ops.h``) built by ``build.py`` / ``make.sh``
* ``capi``             ctypes binding (loads the library lazily, raises if it is missing)
* ``functions``        operator functions / ``torch.autograd.Function`` wrappers
"""

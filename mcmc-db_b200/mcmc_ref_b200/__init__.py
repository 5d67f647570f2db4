"""mcmc_ref_b200 -- B200-native drop-in for mcmc-ref's draw-comparison hot path.

Mirrors the reference's module names for that path (``diagnostics``, ``compare``,
``backends``, ``reference``, ``convert``) and adds the batched forms the reference
lacks (``batch``).  Every number is produced by the sm_100a kernels in
``csrc/`` through the C-ABI in ``include/b200ref.h``; there is no CPU fallback:
importing ``_native`` fails loudly when the shared library or a CUDA device is
missing.
"""

__version__ = "0.1.0"

__all__ = ["batch", "backends", "compare", "convert", "diagnostics", "reference"]

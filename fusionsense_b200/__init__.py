"""fusionsense_b200 — B200 (sm_100a) hot path of FusionSense / DN-Splatter.

Hand-written CUDA kernels behind a C ABI (include/fsb200.h, libfsb200.so) plus the thin Python host layer
that mirrors the reference-facing interfaces: the `gsplat` symbols `dn_splatter/dn_model.py:29-35` imports,
the DN-Splatter loss / optimiser / densification hooks, and `utils/VisualHull.py`.
There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_gsplat_shim(force: bool = False) -> None:
    """Make `import gsplat` (and the sub-modules DN-Splatter / splatfacto import) resolve to this package."""
    if "gsplat" in sys.modules and not force:
        mod = sys.modules["gsplat"]
        if getattr(mod, "__name__", "").startswith("fusionsense_b200"):
            return
        raise RuntimeError("another `gsplat` is already imported; pass force=True to replace it")
    from . import gsplat as shim
    from .gsplat import cuda_legacy, rendering
    from .gsplat.cuda_legacy import _torch_impl, _wrapper

    sys.modules["gsplat"] = shim
    sys.modules["gsplat.rendering"] = rendering
    sys.modules["gsplat.cuda_legacy"] = cuda_legacy
    sys.modules["gsplat.cuda_legacy._wrapper"] = _wrapper
    sys.modules["gsplat.cuda_legacy._torch_impl"] = _torch_impl

"""Import-compatible stand-in for the `gsplat==1.0.0` symbols FusionSense uses.

`dn_splatter/dn_model.py:29-35` (and nerfstudio 1.1.3's splatfacto) import exactly:
`gsplat.rendering.rasterization`, `gsplat.rasterize_gaussians`,
`gsplat.cuda_legacy._torch_impl.quat_to_rotmat`, `gsplat.cuda_legacy._wrapper.num_sh_bases`.
Use `fusionsense_b200.install_gsplat_shim()` (or put `<repo>/shim` on PYTHONPATH) to make
`import gsplat` resolve here.
"""
from .rendering import rasterization, rasterization_from_params  # noqa: F401
from .cuda_legacy._wrapper import num_sh_bases, rasterize_gaussians  # noqa: F401
from .cuda_legacy._torch_impl import quat_to_rotmat  # noqa: F401

__version__ = "1.0.0"

__all__ = ["rasterization", "rasterize_gaussians", "quat_to_rotmat", "num_sh_bases", "rasterization_from_params"]

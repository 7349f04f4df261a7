"""`gsplat.cuda_legacy` stand-in (the 0.1.x API that gsplat 1.0.0 keeps for old callers)."""
from ._torch_impl import quat_to_rotmat  # noqa: F401
from ._wrapper import num_sh_bases, rasterize_gaussians  # noqa: F401

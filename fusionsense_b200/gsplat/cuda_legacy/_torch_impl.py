"""`gsplat.cuda_legacy._torch_impl` stand-in: the one helper DN-Splatter imports from it."""
import torch
import torch.nn.functional as F
from torch import Tensor


def quat_to_rotmat(quat: Tensor) -> Tensor:
    """[..., 4] (w, x, y, z) quaternions, normalised internally -> [..., 3, 3] rotation matrices.

    Call sites: /root/reference/dn_splatter/dn_model.py:286,623,1489,1699,1770,2146.
    """
    assert quat.shape[-1] == 4, quat.shape
    w, x, y, z = torch.unbind(F.normalize(quat, dim=-1), dim=-1)
    mat = torch.stack(
        [
            1 - 2 * (y**2 + z**2),
            2 * (x * y - w * z),
            2 * (x * z + w * y),
            2 * (x * y + w * z),
            1 - 2 * (x**2 + z**2),
            2 * (y * z - w * x),
            2 * (x * z - w * y),
            2 * (y * z + w * x),
            1 - 2 * (x**2 + y**2),
        ],
        dim=-1,
    )
    return mat.reshape(quat.shape[:-1] + (3, 3))

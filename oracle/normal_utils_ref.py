"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's depth -> normal path.

Follows /root/reference/dn_splatter/utils/camera_utils.py:69-89 (get_camera_coords), :92-144
(get_means3d_backproj) and /root/reference/dn_splatter/utils/normal_utils.py:7-20 (pcd_to_normal), :23-46
(normal_from_depth_image).  Pinned: tests/golden/pseudo_normals.npz holds outputs of the unmodified reference
functions (oracle/make_golden_pseudo_normals.py).  Only tests/, smoke() and bench.py's cpu_baseline may import it.
"""
import numpy as np


def get_camera_coords(img_size, pixel_offset=0.5):
    """camera_utils.py:69-89: [H*W, 2] pixel centres, column 0 = u (width), column 1 = v (height)."""
    W, H = img_size
    u, v = np.meshgrid(np.arange(W), np.arange(H), indexing="xy")
    return (np.stack([u, v], axis=-1).reshape(-1, 2) + pixel_offset).astype(np.float32)


def get_means3d_backproj(depths, fx, fy, cx, cy, img_size, c2w):
    """camera_utils.py:92-144 (no mask): fp32, ((u - cx) * d) / fx, then `p @ inv(R) + t`."""
    d = np.asarray(depths, dtype=np.float32).reshape(-1)
    uv = get_camera_coords(img_size)
    p = np.empty((d.shape[0], 3), dtype=np.float32)
    p[:, 0] = (uv[:, 0] - np.float32(cx)) * d / np.float32(fx)
    p[:, 1] = (uv[:, 1] - np.float32(cy)) * d / np.float32(fy)
    p[:, 2] = d
    c2w = np.asarray(c2w, dtype=np.float32)
    rinv = np.linalg.inv(c2w[:3, :3]).astype(np.float32)
    return (p @ rinv + c2w[:3, 3]).astype(np.float32)


def pcd_to_normal(xyz):
    """normal_utils.py:7-20."""
    xyz = np.asarray(xyz, dtype=np.float32)
    hd, wd, _ = xyz.shape
    bottom = xyz[2:hd, 1:wd - 1]
    top = xyz[0:hd - 2, 1:wd - 1]
    right = xyz[1:hd - 1, 2:wd]
    left = xyz[1:hd - 1, 0:wd - 2]
    n = np.cross(right - left, top - bottom).astype(np.float32)
    n = n / np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), np.float32(1e-12))
    out = np.zeros((hd, wd, 3), dtype=np.float32)
    out[1:hd - 1, 1:wd - 1] = n
    return out


def normal_from_depth_image(depths, fx, fy, cx, cy, img_size, c2w):
    """normal_utils.py:23-46 with smooth=False."""
    pts = get_means3d_backproj(depths, fx, fy, cx, cy, img_size, c2w)
    return pcd_to_normal(pts.reshape(img_size[1], img_size[0], 3))

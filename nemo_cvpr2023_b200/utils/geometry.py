"""Mirror of the reference's `utils/geometry.py` (the epsilon variant of the projection)."""
import torch

from .. import functional as Fn
from ..hmr.geometry import batch_rodrigues, quat_to_rotmat, rot6d_to_rotmat  # noqa: F401


def perspective_projection(points, rotation, translation, focal_length, camera_center, epsilon=1e-8):
    """utils/geometry.py:63-93."""
    return Fn.perspective_projection(points, rotation, translation, focal_length, camera_center, epsilon)


def perspective_projection_with_K(points, rotation, translation, K, epsilon=1e-8):
    """utils/geometry.py:95-116.  Arbitrary K: rotation/translation/divide by the kernel with unit
    intrinsics, then the 2x3 part of K applied (p_z/(p_z+eps) is the homogeneous third component)."""
    B = points.shape[0]
    one = torch.ones(B, device=points.device)
    xy = Fn.perspective_projection(points, rotation, translation, one, torch.zeros(B, 2, device=points.device), epsilon)
    cam_z = (torch.einsum('bj,bkj->bk', rotation[:, 2], points) + translation[:, 2:3])
    w = (cam_z / (cam_z + epsilon)).unsqueeze(-1)
    h = torch.cat([xy, w], -1)
    return torch.einsum('bij,bkj->bki', K, h)[..., :2]

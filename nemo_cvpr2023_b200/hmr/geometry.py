"""Mirror of the reference's `hmr/geometry.py` (same names, argument meaning and shapes); the
arithmetic runs in libnemo_b200.so forward and backward.  CUDA tensors only."""
import torch

from .. import functional as Fn


def batch_rodrigues(theta):
    """hmr/geometry.py:9-23: axis-angle (B,3) -> rotation matrices (B,3,3), quaternion form."""
    return Fn.batch_rodrigues(theta, kind=0)


def quat_to_rotmat(quat):
    """hmr/geometry.py:25-45 (not on the fit path; plain tensor expression)."""
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
                        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
                        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z],
                       dim=1).view(-1, 3, 3)


def rot6d_to_rotmat(x):
    """hmr/geometry.py:47-61: (B,6) -> (B,3,3)."""
    return Fn.rot6d_to_rotmat(x)


def apply_extrinsics(points, rotation, translation, inverse=False):
    """hmr/geometry.py:63-76 (not on the fit path)."""
    if not inverse:
        return torch.einsum('bij,bkj->bki', rotation, points) + translation.unsqueeze(1)
    return torch.einsum('bij,bkj->bki', torch.inverse(rotation), points - translation.unsqueeze(1))


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """hmr/geometry.py:78-106 (no epsilon in the divide)."""
    return Fn.perspective_projection(points, rotation, translation, focal_length, camera_center, 0.0)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """hmr/geometry.py:181-210: (N,3,3) or (N,3,4) -> (N,3), NaN entries replaced by 0."""
    if rotation_matrix.shape[1:] == (3, 4):
        rotation_matrix = rotation_matrix[:, :, :3]
    return Fn.rotmat_to_aa(rotation_matrix, nan_to_zero=True)

"""Analytic identities the oracle must satisfy whatever the reference's fixtures say (last row): a second,
independent pin of `oracle/nemo_oracle.py` next to the reference-generated goldens of tests/test_oracle_golden.py.  All in
float64 on the CPU."""
import numpy as np
import torch

from nemo_cvpr2023_b200 import assets
from oracle import nemo_oracle as O

torch.set_default_dtype(torch.float32)


def _sm(sparse=False):
    return O.to_torch(assets.make_smpl_assets(0, sparse=sparse), dtype=torch.float64)


def test_rest_pose_gives_the_shaped_template():
    """identity rotations: vertices == v_template + shapedirs . betas, joints == J_regressor . v_shaped
    (lbs.py:221-252 with R = I: zero pose offsets, rest transforms)"""
    sm = _sm()
    g = torch.Generator().manual_seed(0)
    betas = torch.randn(3, 10, generator=g, dtype=torch.float64)
    eye = torch.eye(3, dtype=torch.float64).expand(3, 24, 3, 3)
    v, j = O.lbs(betas, eye, sm['v_template'], sm['shapedirs'], sm['posedirs'], sm['J_regressor'], sm['parents'],
                 sm['lbs_weights'], pose2rot=False)
    v_shaped = sm['v_template'] + torch.einsum('bl,mkl->bmk', betas, sm['shapedirs'])
    # (the synthetic skinning weights are float32 values: a row sums to 1 +- 1e-7, which is all that is left)
    assert float((v - v_shaped).abs().max()) < 1e-6
    assert float((j - torch.einsum('jv,bvk->bjk', sm['J_regressor'], v_shaped)).abs().max()) < 1e-12


def test_global_rotation_and_translation_are_rigid():
    """a root rotation turns the whole mesh about the root joint; `transl` shifts vertices and all 49 joints"""
    sm = _sm()
    g = torch.Generator().manual_seed(1)
    betas = torch.randn(1, 10, generator=g, dtype=torch.float64)
    body = O.batch_rodrigues_quat(0.4 * torch.randn(23, 3, generator=g, dtype=torch.float64)).view(1, 23, 3, 3)
    eye = torch.eye(3, dtype=torch.float64).view(1, 1, 3, 3)
    R = O.batch_rodrigues_quat(torch.tensor([[0.3, -0.8, 0.5]], dtype=torch.float64)).view(1, 1, 3, 3)
    t = torch.tensor([[0.7, -1.2, 3.0]], dtype=torch.float64)
    a = O.smpl_forward(sm, betas=betas, body_pose=body, global_orient=eye, pose2rot=False)
    b = O.smpl_forward(sm, betas=betas, body_pose=body, global_orient=R, transl=t, pose2rot=False)
    v_shaped = sm['v_template'] + torch.einsum('bl,mkl->bmk', betas, sm['shapedirs'])
    root = torch.einsum('v,bvk->bk', sm['J_regressor'][0], v_shaped)          # rest position of joint 0
    want_v = torch.einsum('ij,bvj->bvi', R[0, 0], a.vertices - root[:, None]) + root[:, None] + t[:, None]
    want_j = torch.einsum('ij,bvj->bvi', R[0, 0], a.joints - root[:, None]) + root[:, None] + t[:, None]
    assert float((b.vertices - want_v).abs().max()) < 1e-6
    assert float((b.joints - want_j).abs().max()) < 1e-6


def test_body_model_equals_lbs_with_skew_rodrigues():
    """`BodyModel.forward` (body_model.py:182-283) == `lbs` on the concatenated axis-angle pose + trans"""
    sm = _sm()
    g = torch.Generator().manual_seed(2)
    ro, pb, ph = (0.3 * torch.randn(2, n, generator=g, dtype=torch.float64) for n in (3, 63, 6))
    betas, tr = torch.randn(2, 10, generator=g, dtype=torch.float64), torch.randn(2, 3, generator=g, dtype=torch.float64)
    bm = O.body_model_forward(sm, ro, pb, ph, betas, tr)
    v, j = O.lbs(betas, torch.cat([ro, pb, ph], 1), sm['v_template'], sm['shapedirs'], sm['posedirs'],
                 sm['J_regressor'], sm['parents'], sm['lbs_weights'])
    assert float((bm['v'] - (v + tr[:, None])).abs().max()) < 1e-12
    assert float((bm['Jtr'] - (j + tr[:, None])).abs().max()) < 1e-12


def test_the_two_rodrigues_agree_and_invert():
    """quaternion (hmr/geometry.py:9-45) and skew-matrix (lbs.py:303-334) Rodrigues give the same rotation, it is
    orthonormal, and `rotation_matrix_to_angle_axis` takes it back for ||r|| < pi (both variants of the NaN handling)"""
    g = torch.Generator().manual_seed(3)
    r = torch.randn(200, 3, generator=g, dtype=torch.float64)
    r = r / r.norm(dim=1, keepdim=True) * torch.rand(200, 1, generator=g, dtype=torch.float64) * 3.0
    Rq, Rs = O.batch_rodrigues_quat(r), O.batch_rodrigues_skew(r)
    assert float((Rq - Rs).abs().max()) < 1e-7        # the +1e-8 inside the angle is the only difference
    eye = torch.eye(3, dtype=torch.float64)
    assert float((Rq @ Rq.transpose(1, 2) - eye).abs().max()) < 1e-12
    assert float((torch.linalg.det(Rq) - 1).abs().max()) < 1e-12
    for ntz in (True, False):
        # the conversion builds its quaternion in float32 (hmr/geometry.py:203 hard-codes the dtype)
        back = O.rotation_matrix_to_angle_axis(Rq, nan_to_zero=ntz).double()
        assert float((back - r).abs().max()) < 5e-6


def test_rot6d_is_a_gram_schmidt_of_the_interleaved_columns():
    """rot6d_to_rotmat (hmr/geometry.py:47-61): proper rotation whose first column is the normalised elements 0, 2, 4; the
    6D vector of a rotation's first two columns reproduces the rotation"""
    g = torch.Generator().manual_seed(4)
    x = torch.randn(50, 6, generator=g, dtype=torch.float64)
    R = O.rot6d_to_rotmat(x)
    eye = torch.eye(3, dtype=torch.float64)
    assert float((R.transpose(1, 2) @ R - eye).abs().max()) < 1e-12
    assert float((torch.linalg.det(R) - 1).abs().max()) < 1e-12
    a1 = x[:, [0, 2, 4]]
    assert float((R[:, :, 0] - a1 / a1.norm(dim=1, keepdim=True)).abs().max()) < 1e-12
    again = O.rot6d_to_rotmat(R[:, :, :2].reshape(50, 6))       # row-major (3,2) = interleaved a1 / a2
    assert float((again - R).abs().max()) < 1e-12


def test_projection_of_the_optical_axis_and_of_a_known_point():
    """perspective_projection (hmr/geometry.py:78-106): a point on the optical axis lands on the principal point, a point
    (x, y, z) in camera coordinates on (f x / z + cx, f y / z + cy); the camera of `learned_camera_projection` uses
    (IMG_D0 // 2, IMG_D1 // 2) as (cx, cy) (:3104-3106)"""
    eye = torch.eye(3, dtype=torch.float64)[None]
    c = torch.tensor([[540.0, 960.0]], dtype=torch.float64)
    pts = torch.tensor([[[0.0, 0.0, 7.0], [1.0, -2.0, 4.0]]], dtype=torch.float64)
    uv = O.perspective_projection(pts, eye, torch.zeros(1, 3, dtype=torch.float64), 5000.0, c)
    assert torch.allclose(uv[0, 0], c[0])
    assert torch.allclose(uv[0, 1], torch.tensor([5000.0 * 1 / 4 + 540, 5000.0 * -2 / 4 + 960], dtype=torch.float64))
    cams = torch.zeros(2, 9, dtype=torch.float64)
    cams[:, 3], cams[:, 6] = 1.0, 1.0                        # rot6d of the identity: a1 = e_x (0,2,4), a2 = e_y (1,3,5)
    cams[1, :3] = torch.tensor([0.0, 0.0, 3.0])
    p = torch.tensor([[[0.0, 0.0, 5.0]], [[0.0, 0.0, 5.0]]], dtype=torch.float64)
    out = O.learned_camera_projection(p, torch.tensor([0, 1]), cams, 5000.0, 1080, 1920)
    assert torch.allclose(out[:, 0], torch.tensor([[540.0, 960.0], [540.0, 960.0]], dtype=torch.float64))


def test_sparse_skinning_weights_change_nothing_but_the_weights():
    """real-asset-shaped weights (<= 4 non-zeros per vertex, rows sum to 1): skinning is a convex combination, so with all
    joints at the same rigid transform the mesh moves rigidly -- for the dense and for the sparse weights alike"""
    for sparse in (False, True):
        sm = _sm(sparse)
        w = sm['lbs_weights']
        assert float((w.sum(1) - 1).abs().max()) < 1e-6 and float(w.min()) >= 0
        if sparse:
            assert int((w > 0).sum(1).max()) <= 4
        R = O.batch_rodrigues_quat(torch.tensor([[0.2, 0.1, -0.4]], dtype=torch.float64)).view(1, 1, 3, 3)
        eye = torch.eye(3, dtype=torch.float64).expand(1, 23, 3, 3)
        out = O.smpl_forward(sm, global_orient=R, body_pose=eye, pose2rot=False)
        root = torch.einsum('v,vk->k', sm['J_regressor'][0], sm['v_template'])
        want = torch.einsum('ij,vj->vi', R[0, 0], sm['v_template'] - root) + root
        assert float((out.vertices[0] - want).abs().max()) < 1e-6     # weights are float32 values: rows sum to 1 +- 1e-7

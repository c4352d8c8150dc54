"""Mirror of `human_body_prior/body_model/lbs.py` (the smplx LBS the reference vendors).
Same function names and argument order; tensors in the reference's (interleaved) layouts.
The composite `lbs` runs the fused kernels; the small helpers are exposed for API parity."""
import torch

from ... import functional as Fn

_CONST_CACHE = {}


def batch_rodrigues(rot_vecs, epsilon=1e-8, dtype=torch.float32):
    """lbs.py:303-334 (skew-matrix form)."""
    return Fn.batch_rodrigues(rot_vecs, kind=1)


def blend_shapes(betas, shape_disps):
    """lbs.py:277-300: (B,nb) x (V,3,nb) -> (B,V,3) through the library GEMM."""
    B, nb = betas.shape
    V3 = shape_disps.shape[0] * 3
    out = torch.empty(B, V3, device=betas.device)
    sd = shape_disps.reshape(V3, nb).contiguous()
    Fn.gemm(betas.contiguous(), sd, out, B, V3, nb, nb, nb, V3, transB=True)
    return out.view(B, -1, 3)


def vertices2joints(J_regressor, vertices):
    """lbs.py:257-274: (J,V) x (B,V,3) -> (B,J,3)."""
    B, V = vertices.shape[:2]
    J = J_regressor.shape[0]
    out = torch.empty(B, J, 3, device=vertices.device)
    v = vertices.contiguous()
    jr = J_regressor.contiguous()
    for b in range(B):
        Fn.gemm(jr, v[b], out[b], J, 3, V, V, 3, 3)
    return out


def batch_rigid_transform(rot_mats, joints, parents, dtype=torch.float32):
    """lbs.py:350-404 -> (posed_joints (B,J,3), rel_transforms (B,J,4,4))."""
    from ..._lib import lib
    B = rot_mats.shape[0]
    R = rot_mats.reshape(B, 24, 9).contiguous()
    J = joints.reshape(B, 72).contiguous()
    A = torch.empty(B, 24, 12, device=R.device)
    JP = torch.empty(B, 24, 3, device=R.device)
    lib.chain_fwd(R, J, 72, parents.to(torch.int32).contiguous(), A, JP, B)
    T = torch.zeros(B, 24, 4, 4, device=R.device)
    T[:, :, :3] = A.view(B, 24, 3, 4)
    T[:, :, 3, 3] = 1
    return JP, T


def _consts_for(v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    key = (v_template.data_ptr(), shapedirs.data_ptr(), posedirs.data_ptr(), J_regressor.data_ptr(),
           lbs_weights.data_ptr())
    if key not in _CONST_CACHE:
        vt = v_template[0] if v_template.dim() == 3 else v_template
        sm = dict(v_template=vt.detach().cpu().numpy(), shapedirs=shapedirs.detach().cpu().numpy(),
                  posedirs=posedirs.detach().cpu().numpy(), J_regressor=J_regressor.detach().cpu().numpy(),
                  lbs_weights=lbs_weights.detach().cpu().numpy(),
                  J_regressor_extra=torch.zeros(0, vt.shape[0]).numpy(), vertex_joint_ids=[],
                  parents=parents.detach().cpu().numpy(), joint_map=list(range(24)),
                  faces=torch.zeros(0, 3, dtype=torch.long).numpy())
        _CONST_CACHE[key] = Fn.SMPLConstants(sm, v_template.device)
    return _CONST_CACHE[key]


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, joints=None,
        pose2rot=True, v_shaped=None, dtype=torch.float32):
    """lbs.py:164-254 -> (verts (B,V,3), posed joints (B,J,3))."""
    if joints is not None or v_shaped is not None:
        raise NotImplementedError('precomputed joints / v_shaped are not used by the NeMo fit path')
    C = _consts_for(v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights)
    B = max(betas.shape[0], pose.shape[0])
    if pose2rot:
        rot = Fn.batch_rodrigues(pose.reshape(-1, 3), kind=1).view(B, 24, 3, 3)
    else:
        rot = pose.reshape(B, 24, 3, 3)
    V, JP, _ = Fn.lbs_rotmats(rot, betas, C, readout=C.no_readout, want_vertices=True)
    return V, JP

"""CPU oracle for the NeMo per-iteration fitting hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain PyTorch on the CPU, of the algorithm the
reference (`wangkua1/nemo-cvpr2023`, mounted at /root/reference in the build
container) executes for one fitting iteration.  Every function cites the
reference file:line it follows.  It is *not* part of the product: only
`tests/`, `__graft_entry__.smoke()` and the CPU-baseline / `--impl reference`
legs of `bench.py` may import it.  The product package never does (it fails
loudly when its CUDA library is missing instead of falling back to this).

Pinning status: the reference ships no test, golden vector or fixture for this
path, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF: `oracle/make_golden.py` imports the unmodified reference
modules in the build container, runs them on the seeded synthetic assets of
`nemo_cvpr2023_b200/assets.py`, and stores the results under `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against those fixtures on every
run (no reference needed).  The one piece that cannot be executed is the
third-party `smplx==0.1.28` package (`environment.yml:295`), which is neither
vendored nor installed: `SMPL.forward` + `VertexJointSelector` are restated
from its published API; its arithmetic (`smplx.lbs`) is
vendored by the reference at `human_body_prior/body_model/lbs.py` and *is*
executed for the fixtures.

Unlike the reference, every function here is dtype-generic so the same code
gives a float64 "truth" (the reference is hard-wired to float32,
`hmr/geometry.py:203`).
"""
from __future__ import annotations

import math
from collections import namedtuple

import torch
import torch.nn.functional as F

ModelOutput = namedtuple('ModelOutput', ['vertices', 'joints', 'full_pose', 'betas',
                                         'global_orient', 'body_pose'])


# --------------------------------------------------------------------------
# rotations
# --------------------------------------------------------------------------
def rot6d_to_rotmat(x):
    """hmr/geometry.py:47-61 (also human_body_prior/models/vposer_model.py:32-45).
    Interleaved 3x2 view: a1 = elements 0,2,4, a2 = elements 1,3,5; Gram-Schmidt;
    the three unit vectors become the COLUMNS of R.  F.normalize eps = 1e-12."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[..., 0], x[..., 1]
    b1 = a1 / a1.norm(dim=1, keepdim=True).clamp_min(1e-12)
    u = a2 - (b1 * a2).sum(1, keepdim=True) * b1
    b2 = u / u.norm(dim=1, keepdim=True).clamp_min(1e-12)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def batch_rodrigues_quat(theta):
    """hmr/geometry.py:9-45: axis-angle -> unit quaternion -> rotation matrix.
    The angle is ||theta + 1e-8|| (the epsilon is added to every component)."""
    ang = (theta + 1e-8).norm(dim=1, keepdim=True)
    axis = theta / ang
    half = 0.5 * ang
    q = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    r = torch.stack([
        w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z], dim=1)
    return r.view(-1, 3, 3)


def batch_rodrigues_skew(rot_vecs):
    """human_body_prior/body_model/lbs.py:303-334: R = I + sin(a) K + (1-cos(a)) K^2,
    a = ||r + 1e-8||, K the skew matrix of r / a."""
    ang = (rot_vecs + 1e-8).norm(dim=1, keepdim=True)
    d = rot_vecs / ang
    c = torch.cos(ang)[:, None]
    s = torch.sin(ang)[:, None]
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    z = torch.zeros_like(rx)
    k = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return eye + s * k + (1 - c) * torch.bmm(k, k)


def rotmat_to_quat(r, eps=1e-6):
    """hmr/geometry.py:266-346 == human_body_prior/tools/tgm_conversion.py:241-323,
    written on the entries of R directly (the reference transposes a 3x4 pad first).
    Four branches chosen by the diagonal; q = 0.5 * q_sel / sqrt(t_sel)."""
    r00, r01, r02 = r[:, 0, 0], r[:, 0, 1], r[:, 0, 2]
    r10, r11, r12 = r[:, 1, 0], r[:, 1, 1], r[:, 1, 2]
    r20, r21, r22 = r[:, 2, 0], r[:, 2, 1], r[:, 2, 2]
    d2 = r22 < eps
    d01 = r00 > r11
    d0n1 = r00 < -r11
    t0 = 1 + r00 - r11 - r22
    t1 = 1 - r00 + r11 - r22
    t2 = 1 - r00 - r11 + r22
    t3 = 1 + r00 + r11 + r22
    q0 = torch.stack([r21 - r12, t0, r10 + r01, r02 + r20], -1)
    q1 = torch.stack([r02 - r20, r10 + r01, t1, r21 + r12], -1)
    q2 = torch.stack([r10 - r01, r02 + r20, r21 + r12, t2], -1)
    q3 = torch.stack([t3, r21 - r12, r02 - r20, r10 - r01], -1)
    c0 = (d2 & d01).to(r.dtype)[:, None]
    c1 = (d2 & ~d01).to(r.dtype)[:, None]
    c2 = (~d2 & d0n1).to(r.dtype)[:, None]
    c3 = (~d2 & ~d0n1).to(r.dtype)[:, None]
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    t = t0[:, None] * c0 + t1[:, None] * c1 + t2[:, None] * c2 + t3[:, None] * c3
    return 0.5 * q / torch.sqrt(t)


def quat_to_angle_axis(q):
    """hmr/geometry.py:213-263 == tgm_conversion.py:325-376 (ceres formula)."""
    q1, q2, q3 = q[..., 1], q[..., 2], q[..., 3]
    s2 = q1 * q1 + q2 * q2 + q3 * q3
    s = torch.sqrt(s2)
    c = q[..., 0]
    two_theta = 2.0 * torch.where(c < 0.0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0.0, two_theta / s, 2.0 * torch.ones_like(s))
    return torch.stack([q1 * k, q2 * k, q3 * k], -1)


def rotation_matrix_to_angle_axis(r, nan_to_zero=True):
    """hmr/geometry.py:181-210 (nan_to_zero=True: `aa[isnan(aa)] = 0`, :209) and
    human_body_prior/tools/rotation_tools.py:73-81 -> tgm_conversion.py:219-239
    (nan_to_zero=False)."""
    aa = quat_to_angle_axis(rotmat_to_quat(r.reshape(-1, 3, 3)))
    if nan_to_zero:
        aa = torch.where(torch.isnan(aa), torch.zeros_like(aa), aa)
    return aa


# --------------------------------------------------------------------------
# projection and 2D / 3D losses
# --------------------------------------------------------------------------
def perspective_projection(points, rotation, translation, focal_length, camera_center,
                           epsilon=0.0):
    """hmr/geometry.py:78-106 (epsilon=0) and utils/geometry.py:63-93 (epsilon=1e-8)."""
    p = torch.einsum('bij,bkj->bki', rotation, points) + translation[:, None]
    p = p / (p[..., 2:3] + epsilon)
    fl = torch.as_tensor(focal_length, dtype=points.dtype, device=points.device).reshape(-1, 1)
    u = fl * p[..., 0] + camera_center[:, 0:1]
    v = fl * p[..., 1] + camera_center[:, 1:2]
    return torch.stack([u, v], -1)


def perspective_projection_with_K(points, rotation, translation, K, epsilon=1e-8):
    """utils/geometry.py:95-116."""
    p = torch.einsum('bij,bkj->bki', rotation, points) + translation[:, None]
    p = p / (p[..., 2:3] + epsilon)
    return torch.einsum('bij,bkj->bki', K, p)[..., :2]


def learned_camera_projection(points3d, view_idx, cameras, focal, img_d0, img_d1):
    """MultiViewModel.learned_camera_projection, nemo/neural_motion_model.py:3073-3124.
    Per view: t = cam[:3], R = rot6d(cam[3:]), centre = (IMG_D0//2, IMG_D1//2)
    [sic: height/2 is used as cx], hmr projection without epsilon."""
    out = torch.zeros(points3d.shape[0], points3d.shape[1], 2, dtype=points3d.dtype)
    centre = torch.tensor([[float(img_d0 // 2), float(img_d1 // 2)]], dtype=points3d.dtype)
    for v in range(cameras.shape[0]):
        sel = torch.where(view_idx == v)[0]
        if sel.numel() == 0:
            continue
        n = sel.numel()
        rot = rot6d_to_rotmat(cameras[v, 3:][None]).expand(n, -1, -1)
        out[sel] = perspective_projection(points3d[sel], rot, cameras[v, :3][None].expand(n, -1),
                                          focal, centre.expand(n, -1))
    return out


def gmof(residual, rho=100.0, sqrt=False):
    """nemo/utils/misc_utils.py:91-105."""
    sq = residual ** 2
    if sqrt:
        sq = torch.sqrt(sq.sum(-1)).unsqueeze(-1)
    return rho ** 2 * (sq / (sq + rho ** 2))


def keypoint_loss(pred, gt, weight, gt_size=None, loss_type='mse_robust', rho=100.0):
    """MultiViewModel.keypoint_loss, nemo/neural_motion_model.py:2806-2843."""
    m = (weight > 0.5).to(pred.dtype)
    if loss_type in ('rmse', 'rmse_resized'):
        if loss_type == 'rmse_resized':
            s = gt_size[:, None, None]
            pred, gt = pred / s, gt / s
        return m * torch.sqrt(1e-6 + ((pred - gt) ** 2).sum(-1, keepdim=True))
    if loss_type == 'mse':
        return m * (pred - gt) ** 2
    if loss_type == 'rmse_robust':
        return m * gmof(pred - gt, rho, sqrt=True)
    if loss_type == 'mse_robust':
        return m * gmof(pred - gt, rho, sqrt=False)
    if loss_type == 'mse_robust_resized':
        s = gt_size[:, None, None]
        return m * gmof(pred / s * 1000 - gt / s * 1000, rho, sqrt=False)
    raise ValueError(loss_type)


def per_view_mean(loss_all, conf, view_idx):
    """The reduction inside step(), nemo/neural_motion_model.py:3551-3558:
    mean over the views present of mean_{n in view, joint, xy}(loss * conf)."""
    total = 0
    views = view_idx.unique()
    for v in views:
        sel = view_idx == v
        total = total + (loss_all[sel] * conf[sel]).mean()
    return total / len(views)


# --------------------------------------------------------------------------
# SMPL linear blend skinning
# --------------------------------------------------------------------------
def blend_shapes(betas, shapedirs):
    """lbs.py:277-300."""
    return torch.einsum('bl,mkl->bmk', betas, shapedirs)


def vertices2joints(regressor, vertices):
    """lbs.py:257-274."""
    return torch.einsum('bik,ji->bjk', vertices, regressor)


def batch_rigid_transform(rot_mats, joints, parents):
    """lbs.py:350-404.  G_i = G_parent(i) [R_i | J_i - J_parent(i)];
    A_i = G_i - [0 | G_i J_i]; posed joint i = translation of G_i."""
    b, nj = rot_mats.shape[:2]
    rel = joints.clone()
    rel[:, 1:] = joints[:, 1:] - joints[:, parents[1:]]
    g = torch.zeros(b, nj, 4, 4, dtype=rot_mats.dtype)
    g[:, :, :3, :3] = rot_mats
    g[:, :, :3, 3] = rel
    g[:, :, 3, 3] = 1
    chain = [g[:, 0]]
    for i in range(1, nj):
        chain.append(chain[int(parents[i])] @ g[:, i])
    world = torch.stack(chain, 1)
    posed = world[:, :, :3, 3]
    jh = torch.cat([joints, torch.zeros_like(joints[..., :1])], -1)[..., None]
    corr = world @ jh
    rel_t = world.clone()
    rel_t[..., 3:4] = world[..., 3:4] - corr
    return posed, rel_t


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights,
        joints=None, pose2rot=True, v_shaped=None):
    """lbs.py:164-254."""
    b = max(betas.shape[0], pose.shape[0])
    if v_shaped is None:
        v_shaped = v_template + blend_shapes(betas, shapedirs)
    j = joints if joints is not None else vertices2joints(J_regressor, v_shaped)
    eye = torch.eye(3, dtype=betas.dtype)
    if pose2rot:
        rot = batch_rodrigues_skew(pose.reshape(-1, 3)).view(b, -1, 3, 3)
    else:
        rot = pose.reshape(b, -1, 3, 3)
    feat = (rot[:, 1:] - eye).reshape(b, -1)
    v_posed = v_shaped + (feat @ posedirs).view(b, -1, 3)
    j_posed, a = batch_rigid_transform(rot, j, parents)
    nj = J_regressor.shape[0]
    t = (lbs_weights[None].expand(b, -1, -1) @ a.view(b, nj, 16)).view(b, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones_like(v_posed[..., :1])], -1)
    verts = (t @ vh[..., None])[:, :, :3, 0]
    return verts, j_posed


def smpl_forward(sm, betas=None, body_pose=None, global_orient=None, transl=None,
                 pose2rot=True):
    """hmr/smpl.py:29-43 on top of smplx.SMPL.forward (smplx 0.1.28, restated):
    full_pose = cat(global_orient, body_pose); betas broadcast to the batch; lbs;
    VertexJointSelector appends 21 picked vertices; +transl; then 9 extra joints
    regressed from the (translated) vertices and the 49-entry joint_map.
    `sm` is a dict of torch tensors as produced by assets.make_smpl_assets."""
    dt = sm['v_template'].dtype
    if global_orient is None:
        global_orient = torch.zeros(1, 3, dtype=dt) if pose2rot else torch.eye(3, dtype=dt).view(1, 1, 3, 3)
    if body_pose is None:
        body_pose = torch.zeros(1, 69, dtype=dt) if pose2rot else torch.eye(3, dtype=dt).expand(1, 23, 3, 3)
    if betas is None:
        betas = torch.zeros(1, 10, dtype=dt)
    full_pose = torch.cat([global_orient, body_pose], dim=1)
    b = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
    if betas.shape[0] != b:
        betas = betas.expand(int(b / betas.shape[0]), -1)
    verts, joints = lbs(betas, full_pose, sm['v_template'], sm['shapedirs'], sm['posedirs'],
                        sm['J_regressor'], sm['parents'], sm['lbs_weights'], pose2rot=pose2rot)
    joints = torch.cat([joints, verts[:, sm['vertex_joint_ids']]], dim=1)
    if transl is not None:
        joints = joints + transl[:, None]
        verts = verts + transl[:, None]
    extra = vertices2joints(sm['J_regressor_extra'], verts)
    joints = torch.cat([joints, extra], dim=1)[:, sm['joint_map']]
    return ModelOutput(vertices=verts, joints=joints, full_pose=full_pose, betas=betas,
                       global_orient=global_orient, body_pose=body_pose)


def body_model_forward(sm, root_orient=None, pose_body=None, pose_hand=None, betas=None, trans=None):
    """human_body_prior/body_model/body_model.py:182-283 for model_type 'smpl':
    axis-angle full pose (72), skew Rodrigues inside lbs, trans added to both outputs."""
    dt = sm['v_template'].dtype
    b = 1
    for a in (root_orient, pose_body, pose_hand, betas, trans):
        if a is not None:
            b = a.shape[0]
            break
    root_orient = torch.zeros(b, 3, dtype=dt) if root_orient is None else root_orient
    pose_body = torch.zeros(b, 63, dtype=dt) if pose_body is None else pose_body
    pose_hand = torch.zeros(b, 6, dtype=dt) if pose_hand is None else pose_hand
    betas = torch.zeros(b, 10, dtype=dt) if betas is None else betas
    trans = torch.zeros(b, 3, dtype=dt) if trans is None else trans
    full_pose = torch.cat([root_orient, pose_body, pose_hand], -1)
    v, jtr = lbs(betas, full_pose, sm['v_template'][None].expand(b, -1, -1), sm['shapedirs'],
                 sm['posedirs'], sm['J_regressor'], sm['parents'], sm['lbs_weights'])
    return dict(v=v + trans[:, None], Jtr=jtr + trans[:, None], full_pose=full_pose)


# --------------------------------------------------------------------------
# motion field: phase warp, RBF, MLP
# --------------------------------------------------------------------------
def monotonic_forward(x, shifts, scales):
    """monotonic_network.py:23-39."""
    sh, sc = torch.relu(shifts), torch.relu(scales)

    def net(u):
        return torch.sigmoid(sc * (u.expand(-1, sh.shape[0]) - sh)).mean(-1, keepdim=True)

    y, y0, y1 = net(x), net(torch.zeros_like(x)), net(torch.ones_like(x))
    return (y - y0) / (y1 - y0 + 1e-6)


def _rbf_basis(name):
    """nemo/rbf.py:62-118 (the layer picks one by `eval(name)`, :41)."""
    s3, s5 = 3 ** 0.5, 5 ** 0.5
    return {
        'gaussian': lambda a: torch.exp(-a.pow(2)),
        'linear': lambda a: a,
        'quadratic': lambda a: a.pow(2),
        'inverse_quadratic': lambda a: 1 / (1 + a.pow(2)),
        'multiquadric': lambda a: (1 + a.pow(2)).pow(0.5),
        'inverse_multiquadric': lambda a: 1 / (1 + a.pow(2)).pow(0.5),
        'spline': lambda a: a.pow(2) * torch.log(a + 1),
        'poisson_one': lambda a: (a - 1) * torch.exp(-a),
        'poisson_two': lambda a: ((a - 2) / 2) * a * torch.exp(-a),
        'matern32': lambda a: (1 + s3 * a) * torch.exp(-s3 * a),
        'matern52': lambda a: (1 + s5 * a + (5 / 3) * a.pow(2)) * torch.exp(-s5 * a),
    }[name]


def rbf_forward(x, centres, log_sigmas, basis='quadratic'):
    """nemo/rbf.py:47-56: phi((x - c_i)^2 / exp(log_sigma_i)), x (N,1), centres (D,1)."""
    d = (x[:, None, :] - centres[None]).pow(2).sum(-1) / torch.exp(log_sigmas)[None]
    return _rbf_basis(basis)(d)


def motionnet_forward(p, x, prefix='learned_motion.'):
    """FCNN + MotionNet.forward, nemo/neural_motion_model.py:58-71, :130-148.
    Returns rot6d (N,144), rotmat (N,24,3,3), pose axis-angle (N,72), trans (N,3)."""
    g = lambda k: p[prefix + k]
    h = torch.relu(F.linear(x, g('net.net.0.weight'), g('net.net.0.bias')))
    h = torch.relu(F.linear(h, g('net.net.2.weight'), g('net.net.2.bias')))
    z = torch.relu(F.linear(h, g('net.net.4.weight'), g('net.net.4.bias')))
    rot6d = F.linear(z, g('rot_out.weight'), g('rot_out.bias'))
    rotmat = rot6d_to_rotmat(rot6d).view(x.shape[0], 24, 3, 3)
    pose = rotation_matrix_to_angle_axis(rotmat.reshape(-1, 3, 3)).reshape(-1, 72)
    trans = F.linear(z, g('linear_out.weight'), g('linear_out.bias'))
    return rot6d, rotmat, pose, trans


# --------------------------------------------------------------------------
# priors
# --------------------------------------------------------------------------
def vposer_encode(vp, pose_body):
    """VPoser.encode, human_body_prior/models/vposer_model.py:63-75,90-96 in eval
    mode (model_loader.py:83): BN uses running stats, Dropout is off.
    Returns (mu, sigma) with sigma = softplus(logvar-layer) used as the std."""
    def bn(x, pre):
        return F.batch_norm(x, vp[pre + '.running_mean'], vp[pre + '.running_var'],
                            vp[pre + '.weight'], vp[pre + '.bias'], False, 0.0, 1e-5)
    x = bn(pose_body.reshape(pose_body.shape[0], -1), 'encoder_net.1')
    x = F.leaky_relu(F.linear(x, vp['encoder_net.2.weight'], vp['encoder_net.2.bias']), 0.01)
    x = bn(x, 'encoder_net.4')
    x = F.linear(x, vp['encoder_net.6.weight'], vp['encoder_net.6.bias'])
    x = F.linear(x, vp['encoder_net.7.weight'], vp['encoder_net.7.bias'])
    mu = F.linear(x, vp['encoder_net.8.mu.weight'], vp['encoder_net.8.mu.bias'])
    sigma = F.softplus(F.linear(x, vp['encoder_net.8.logvar.weight'], vp['encoder_net.8.logvar.bias']))
    return mu, sigma


def vposer_decode(vp, z):
    """VPoser.decode, vposer_model.py:77-88,98-106: MLP -> 21x6D -> rotmat -> axis-angle
    through the tgm conversion (no NaN patch).  Returns (pose_body (N,21,3), matrot (N,21,9))."""
    x = F.leaky_relu(F.linear(z, vp['decoder_net.0.weight'], vp['decoder_net.0.bias']), 0.01)
    x = F.leaky_relu(F.linear(x, vp['decoder_net.3.weight'], vp['decoder_net.3.bias']), 0.01)
    x = F.linear(x, vp['decoder_net.5.weight'], vp['decoder_net.5.bias'])
    r = rot6d_to_rotmat(x)
    aa = rotation_matrix_to_angle_axis(r, nan_to_zero=False)
    return aa.view(z.shape[0], -1, 3), r.view(z.shape[0], -1, 9)


def gmm_prior(gmm, pose):
    """MaxMixturePrior.merged_log_likelihood, hmr/smplify/prior.py:181-196."""
    d = pose[:, None] - gmm['means']
    pd = torch.einsum('mij,bmj->bmi', gmm['precisions'], d)
    ll = 0.5 * (pd * d).sum(-1) - torch.log(gmm['nll_weights'])
    return ll.min(dim=1)[0]


# --------------------------------------------------------------------------
# the fit model (NemoV1..V4 semantics)
# --------------------------------------------------------------------------
class FitOracle:
    """NemoV2 / NemoV3 / NemoV4 `get_preds_batch` + `step`
    (nemo/neural_motion_model.py:3364-3453, :3511-3598, :3637-3672, :3675-3781,
    :3796-3956, :3959-4058) on plain tensors.

    `params`: dict name -> leaf tensor (reference state_dict names, see
    assets.make_fit_params).  `sm`, `vp`, `gmm`, `data`: dicts of tensors."""

    def __init__(self, args, params, sm, vp, gmm, data, version=2, focal=5000.0):
        self.args, self.p, self.sm, self.vp, self.gmm, self.data = args, params, sm, vp, gmm, data
        self.version = version
        self.focal = focal
        self.num_views = params['learned_cameras'].shape[0]
        self.num_frames = data['points2d_gt_all'].shape[1]
        self.training = False
        self.optimizers = None

    # optimizers exactly as NemoV1/V2.__init__ builds them (:3404-3431, :3694-3721)
    def build_optimizers(self):
        a, p = self.args, self.p
        for t in p.values():
            t.requires_grad_(True)
        p['phase_rbf.centres'].requires_grad_(False) if 'phase_rbf.centres' in p else None
        cls = torch.optim.Adam if a.opt_human == 'adam' else torch.optim.AdamW
        motion = [p[k] for k in p if k.startswith('learned_motion.')]
        if a.phase_rbf_dim > 0:
            motion = motion + [p['phase_rbf.log_sigmas']]
        phase = [p[k] for k in p if k.startswith('phase_networks.')]
        self.opt_cameras = torch.optim.Adam([p['learned_cameras']], lr=a.lr_camera, weight_decay=0)
        self.opt_motion = cls(motion, lr=a.lr_human, weight_decay=a.wd_human)
        self.opt_phase = torch.optim.Adam(phase, lr=a.lr_phase, weight_decay=0.0)
        self.optimizers = [self.opt_cameras, self.opt_motion, self.opt_phase]
        if a.instance_code_size > 0:
            self.opt_instance = torch.optim.Adam([p['learned_instance_code']], lr=a.lr_instance, weight_decay=0)
            self.optimizers.append(self.opt_instance)
        return self

    def raw_phase(self, frame_idx):
        """:2978-2984"""
        return torch.linspace(0, 1, self.num_frames, dtype=self.p['learned_cameras'].dtype)[frame_idx]

    def motion_input(self, view_idx, phases):
        a, p = self.args, self.p
        x = phases
        if a.phase_rbf_dim > 0:
            x = rbf_forward(x, p['phase_rbf.centres'], p['phase_rbf.log_sigmas'], a.rbf_kernel)
        if a.instance_code_size > 0:
            code = p['learned_instance_code'][view_idx]
            if self.version >= 3 and self.training and a.code_noise > 0:
                code = code + a.code_noise * torch.randn_like(code)
            x = torch.cat([x, code], 1)
        return x

    def smpl_given_poses(self, pose_input, orient6d, pose_type):
        """:2963-2976"""
        if pose_type == 'aa':
            rot = batch_rodrigues_quat(pose_input.reshape(-1, 3)).reshape(-1, 23, 3, 3)
        else:
            rot = pose_input
        return smpl_forward(self.sm, betas=self.p['learned_betas'], body_pose=rot,
                            global_orient=rot6d_to_rotmat(orient6d)[:, None], pose2rot=False)

    def get_preds_batch(self, view_idx, frame_idx, add_trans=True, phases=None, detach_pose=False):
        a, p = self.args, self.p
        n = len(view_idx)
        raw = (self.raw_phase(frame_idx) if phases is None else phases)[:, None]
        warped = torch.stack([monotonic_forward(raw, p['phase_networks.%d.shifts' % v],
                                                p['phase_networks.%d.scales' % v])
                              for v in range(self.num_views)])
        ph = warped[view_idx, torch.arange(n)]
        x = self.motion_input(view_idx, ph)
        rot6d, rotmat, pose, trans = motionnet_forward(p, x)
        body_rot = rotmat[:, 1:].detach() if detach_pose else rotmat[:, 1:]
        out = self.smpl_given_poses(body_rot, rot6d[:, :6], 'rotmat')
        dt = x.dtype
        if a.phase_rbf_dim > 0:
            x0 = torch.cat([rbf_forward(torch.zeros(1, 1, dtype=dt), p['phase_rbf.centres'],
                                        p['phase_rbf.log_sigmas'], a.rbf_kernel),
                            torch.zeros(1, a.instance_code_size, dtype=dt)], 1)
        else:
            x0 = torch.zeros(1, 1 + a.instance_code_size, dtype=dt)
        trans = trans - motionnet_forward(p, x0)[3]
        v, j = out.vertices, out.joints
        if add_trans:
            v, j = v + trans[:, None], j + trans[:, None]
        idx = list(range(25)) if self.version >= 4 else [38] + list(range(1, 25))
        return dict(view_idx=view_idx, frame_idx=frame_idx, v=v, j=j[:, idx], poses=pose[:, 3:],
                    orient=rot6d[:, :6], orient_aa=pose[:, :3], trans=trans)

    def vposer_loss(self, poses, orient):
        """:2775-2804"""
        n = poses.shape[0]
        mu, sigma = vposer_encode(self.vp, poses[:, :63])
        dec, _ = vposer_decode(self.vp, mu)
        recon = torch.cat([dec.reshape(n, -1), poses[:, 63:]], 1)
        v_orig = self.smpl_given_poses(poses, orient, 'aa').vertices
        v_rec = self.smpl_given_poses(recon, orient, 'aa').vertices
        v2v = (v_rec.detach() - v_orig).abs().mean()
        kl = (-torch.log(sigma) + 0.5 * (sigma ** 2 + mu ** 2) - 0.5).sum(1).mean()
        return v2v, kl

    def losses(self, view_idx, frame_idx):
        a, d = self.args, self.data
        pred = self.get_preds_batch(view_idx, frame_idx)
        pts2d = learned_camera_projection(pred['j'], view_idx, self.p['learned_cameras'], self.focal,
                                          d['IMG_D0'], d['IMG_D1'])
        gt = d['points2d_gt_all'][view_idx, frame_idx]
        size = d['gt_bbox_size'][view_idx, frame_idx]
        loss_all = keypoint_loss(pts2d, gt[..., :2], gt[..., 2:], size, a.loss)
        kp = per_view_mean(loss_all, gt[..., 2:], view_idx)
        out = dict(kp_loss=kp)
        loss = kp
        v2v, kl = self.vposer_loss(pred['poses'], pred['orient'])
        if a.weight_vp_loss:
            loss = loss + a.weight_vp_loss * v2v
        if a.weight_vp_z_loss:
            loss = loss + a.weight_vp_z_loss * kl
        if self.version >= 3:
            if a.weight_instance_loss:
                inst = (self.p['learned_instance_code'] ** 2).mean()
                loss = loss + a.weight_instance_loss * inst
                out['instance_loss'] = inst
            if a.weight_3d_loss:
                l3 = keypoint_loss(pred['poses'], d['hmr_theta'][view_idx, frame_idx],
                                   d['hmr_mask'][view_idx, frame_idx], loss_type='mse_robust').mean()
                loss = loss + a.weight_3d_loss * l3
                out['loss_3d'] = l3
        g = gmm_prior(self.gmm, pred['poses']).mean()
        if a.weight_gmm_loss:
            loss = loss + a.weight_gmm_loss * g
        out.update(gmm_loss=g, vp_recon_loss=v2v, vp_kl_loss=kl, total_loss=loss)
        return out, dict(view_idx=view_idx, frame_idx=frame_idx, loss_all=loss_all, points2d_gt=gt,
                         points2d=pts2d, pred=pred)

    def step(self, view_idx, frame_idx, update=True):
        """:3511-3598 / :3796-3909"""
        if update and self.version >= 3:
            self.training = True
        out, info = self.losses(view_idx, frame_idx)
        if update:
            for o in self.optimizers:
                o.zero_grad()
            out['total_loss'].backward()
            for o in self.optimizers:
                o.step()
        self.training = False
        return {k: float(v.detach()) for k, v in out.items()}, info

    def warmup_loss(self, view_idx, frame_idx):
        """NemoV1.warmup body, :3455-3509."""
        pred = self.get_preds_batch(view_idx, frame_idx)
        d = self.data
        return keypoint_loss(pred['poses'], d['hmr_theta'][view_idx, frame_idx],
                             d['hmr_mask'][view_idx, frame_idx], loss_type='mse_robust').mean()


def to_torch(d, dtype=torch.float32):
    """numpy asset dict -> torch dict (floats cast to `dtype`, ints to int64)."""
    import numpy as np
    out = {}
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            t = torch.from_numpy(v.copy())
            out[k] = t.to(dtype) if t.is_floating_point() else t.long()
        else:
            out[k] = v
    return out

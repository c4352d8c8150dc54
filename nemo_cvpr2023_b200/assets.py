"""Deterministic synthetic stand-ins for the licensed assets of the NeMo fit path.

The reference needs SMPL (`hmr/hmr_config.py:70-76`), the SPIN extra joint
regressor (`hmr/smpl.py:23`), the VPoser V02_05 snapshot
(`nemo/neural_motion_model.py:225-231`), the 8-component GMM pose prior
(`hmr/smplify/prior.py:116-160`) and the NeMo-MoCap keypoints
(`nemo/multi_view_sequence.py:415-443`).  None of them can be shipped, so every
consumer in this repo (product, oracle, reference harness, bench) draws the
same tensors from here.  Shapes are the exact SMPL / VPoser / GMM shapes; values
follow SURVEY.md section 8(d).

Everything is generated with ``numpy.random.RandomState`` (frozen legacy
generator, bit-stable across numpy versions) in float64 and cast to float32, so
the GPU box and this container see identical bytes without shipping 20 MB files.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_BASIS = 207
NUM_FACES = 13776

# SMPL kinematic tree (kintree_table[0]); consumed by the reference at
# human_body_prior/body_model/body_model.py:247 and smplx.SMPL.parents.
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17,
                18, 19, 20, 21]

# smplx 0.1.28 `VERTEX_IDS['smplh']` in the order VertexJointSelector appends
# them (face 5, feet 6, left finger tips 5, right finger tips 5) -> raw joints
# 24..44.  SURVEY.md section 8(c).
VERTEX_JOINT_IDS = [332, 6260, 2800, 4071, 583,
                    3216, 3226, 3387, 6617, 6624, 6787,
                    2746, 2319, 2445, 2556, 2673,
                    6191, 5782, 5905, 6016, 6133]

# hmr/hmr_constants.py:42-149: JOINT_MAP[JOINT_NAMES[i]] for the 49 output joints,
# as indices into the 54 raw joints (24 kinematic + 21 vertex picks + 9 extra).
JOINT_MAP_49 = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7,
                25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20,
                47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]

FOCAL_LENGTH = 5000.0  # hmr/hmr_constants.py:1


def _rs(seed):
    return np.random.RandomState(seed)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def make_smpl_assets(seed: int = 0, sparse: bool = False) -> dict:
    """SMPL-shaped random model.  All float32 numpy arrays.

    sparse=True: the sparsity pattern of the licensed SMPL files -- at most 4 non-zero skinning weights per vertex
    (`human_body_prior/body_model/lbs.py:241-245` multiplies the dense (V, 24) matrix regardless) and joint regressors
    that each look at a few hundred vertices -- so that real-asset-shaped weights are exercised by the tests.

    v_template (6890,3)  shapedirs (6890,3,10)  posedirs (207,20670)
    J_regressor (24,6890)  J_regressor_extra (9,6890)  lbs_weights (6890,24)
    parents (24,) int64  vertex_joint_ids (21,) int64  joint_map (49,) int64
    faces (13776,3) int64
    """
    r = _rs(seed)
    v_template = 0.3 * r.standard_normal((NUM_VERTS, 3))
    shapedirs = 0.01 * r.standard_normal((NUM_VERTS, 3, NUM_BETAS))
    posedirs = 0.001 * r.standard_normal((NUM_POSE_BASIS, NUM_VERTS * 3))
    jr = r.uniform(0.0, 1.0, (NUM_JOINTS, NUM_VERTS))
    jr /= jr.sum(1, keepdims=True)
    jre = r.uniform(0.0, 1.0, (9, NUM_VERTS))
    jre /= jre.sum(1, keepdims=True)
    logits = 4.0 * r.standard_normal((NUM_VERTS, NUM_JOINTS))
    logits -= logits.max(1, keepdims=True)
    w = np.exp(logits)
    w /= w.sum(1, keepdims=True)
    if sparse:
        keep = np.argsort(-w, axis=1)[:, :4]
        m = np.zeros_like(w, dtype=bool)
        np.put_along_axis(m, keep, True, axis=1)
        m[::7, :] &= (np.cumsum(m[::7, :], axis=1) <= 2)          # some vertices follow two joints only, some one
        m[::31, :] &= (np.cumsum(m[::31, :], axis=1) <= 1)
        w = np.where(m, w, 0.0)
        w /= w.sum(1, keepdims=True)
        for reg in (jr, jre):
            for row in reg:
                row[r.permutation(NUM_VERTS)[300:]] = 0.0
            reg /= reg.sum(1, keepdims=True)
    faces = r.randint(0, NUM_VERTS, size=(NUM_FACES, 3)).astype(np.int64)
    return dict(
        v_template=_f32(v_template), shapedirs=_f32(shapedirs), posedirs=_f32(posedirs),
        J_regressor=_f32(jr), J_regressor_extra=_f32(jre), lbs_weights=_f32(w),
        parents=np.asarray(SMPL_PARENTS, dtype=np.int64),
        vertex_joint_ids=np.asarray(VERTEX_JOINT_IDS, dtype=np.int64),
        joint_map=np.asarray(JOINT_MAP_49, dtype=np.int64),
        faces=faces,
    )


def _linear(r, fan_out, fan_in):
    b = 1.0 / math.sqrt(fan_in)
    return _f32(r.uniform(-b, b, (fan_out, fan_in))), _f32(r.uniform(-b, b, (fan_out,)))


def make_vposer_weights(seed: int = 2, num_neurons: int = 512, latentD: int = 32) -> dict:
    """State-dict shaped like human_body_prior/models/vposer_model.py:59-88
    (keys are the reference's `state_dict()` keys).  BatchNorm running stats are
    non-trivial so that the eval-mode affine fold is actually exercised."""
    r = _rs(seed)
    nf = 63
    sd = {}

    def bn(prefix, n):
        sd[prefix + '.weight'] = _f32(r.uniform(0.8, 1.2, n))
        sd[prefix + '.bias'] = _f32(0.1 * r.standard_normal(n))
        sd[prefix + '.running_mean'] = _f32(0.1 * r.standard_normal(n))
        sd[prefix + '.running_var'] = _f32(r.uniform(0.5, 1.5, n))

    def lin(prefix, o, i):
        w, b = _linear(r, o, i)
        sd[prefix + '.weight'] = w
        sd[prefix + '.bias'] = b

    bn('encoder_net.1', nf)
    lin('encoder_net.2', num_neurons, nf)
    bn('encoder_net.4', num_neurons)
    lin('encoder_net.6', num_neurons, num_neurons)
    lin('encoder_net.7', num_neurons, num_neurons)
    lin('encoder_net.8.mu', latentD, num_neurons)
    lin('encoder_net.8.logvar', latentD, num_neurons)
    lin('decoder_net.0', num_neurons, latentD)
    lin('decoder_net.3', num_neurons, num_neurons)
    lin('decoder_net.5', 21 * 6, num_neurons)
    # keep the decoded 6D vectors away from the degenerate (zero-norm) region
    ident = np.tile(np.asarray([1, 0, 0, 1, 0, 0], np.float32), 21)
    sd['decoder_net.5.bias'] = _f32(sd['decoder_net.5.bias'] + ident)
    return sd


def make_gmm(seed: int = 3, num_gaussians: int = 8, dim: int = 69) -> dict:
    """Buffers of hmr/smplify/prior.py:142-160 (`means`, `precisions`, `nll_weights`)."""
    r = _rs(seed)
    means = 0.2 * r.standard_normal((num_gaussians, dim))
    a = 0.05 * r.standard_normal((num_gaussians, dim, dim))
    prec = np.einsum('mij,mkj->mik', a, a) + np.eye(dim)[None]
    lw = r.standard_normal(num_gaussians)
    lw -= lw.max()
    w = np.exp(lw)
    w /= w.sum()
    return dict(means=_f32(means), precisions=_f32(prec), nll_weights=_f32(w[None]))


def make_fit_data(num_views: int, num_frames: int, seed: int = 1) -> dict:
    """Targets in the layout of MultiViewModel.collate_gt_2d
    (nemo/neural_motion_model.py:2908-2961) and NemoV1.__init__ (:3440-3453)."""
    r = _rs(seed)
    x = r.uniform(0.0, 1920.0, (num_views, num_frames, 25, 1))
    y = r.uniform(0.0, 1080.0, (num_views, num_frames, 25, 1))
    c = r.uniform(0.0, 1.0, (num_views, num_frames, 25, 1))
    pts = np.concatenate([x, y, c], -1)
    theta = 0.2 * r.standard_normal((num_views, num_frames, 69))
    return dict(
        points2d_gt_all=_f32(pts),
        gt_bbox_size=_f32(np.full((num_views, num_frames), 500.0)),
        hmr_theta=_f32(theta),
        hmr_mask=_f32(np.ones((num_views, num_frames, 1))),
        IMG_D0=1080, IMG_D1=1920,
    )


def default_args(**over) -> SimpleNamespace:
    """The argparse/YAML namespace the fit models read (`self.args`).  Defaults
    follow config/default-v1.yml + run_scripts_examples/nemomocap-example.sh."""
    d = dict(
        batch_size=512, loss='mse_robust', label_type='op',
        h_dim=1000, phase_rbf_dim=100, rbf_kernel='quadratic', instance_code_size=5,
        monotonic_network_n_nodes=200, phase_init='linear',
        lr_camera=0.1, lr_human=1e-4, lr_phase=1e-3, lr_instance=1e-3, wd_human=1e-3,
        opt_human='adam', lr_factor=1.0,
        weight_vp_loss=10.0, weight_vp_z_loss=1.0, weight_gmm_loss=1.0,
        weight_instance_loss=0.0, weight_3d_loss=0.0, code_noise=0.0,
        load_ckpt_path=None, out_dir=None,
    )
    d.update(over)
    return SimpleNamespace(**d)


def make_fit_params(args, num_views: int, seed: int = 4, img_d0: int = 1080,
                    rot_out_std: float = 0.02) -> dict:
    """Initial trainable parameters with the reference's state_dict key names
    (SURVEY.md section 5).  Initialisation follows NemoV1/V2.__init__
    (nemo/neural_motion_model.py:3375-3402, :3685-3692) with `rot_out.weight`
    drawn N(0, 0.02^2) instead of xavier(gain=1e-5) to sit away from the
    rotmat->axis-angle identity singularity."""
    r = _rs(seed)
    h = args.h_dim
    in_dim = (args.phase_rbf_dim if args.phase_rbf_dim > 0 else 1) + args.instance_code_size
    p = {}
    cam = 1e-4 * r.standard_normal((num_views, 9))
    cam[:, 3] += 1
    cam[:, 6] += 1
    cam[:, 2] += 2 * FOCAL_LENGTH / (img_d0 * 1 + 1e-9)
    p['learned_cameras'] = _f32(cam)
    if args.instance_code_size > 0:
        p['learned_instance_code'] = _f32(1e-4 * r.standard_normal((num_views, args.instance_code_size)))
    for idx, (o, i) in zip((0, 2, 4), ((h, in_dim), (h, h), (h, h))):
        w, b = _linear(r, o, i)
        p['learned_motion.net.net.%d.weight' % idx] = w
        p['learned_motion.net.net.%d.bias' % idx] = b
    p['learned_motion.rot_out.weight'] = _f32(rot_out_std * r.standard_normal((144, h)))
    p['learned_motion.rot_out.bias'] = _f32(np.tile(np.asarray([1, 0, 0, 1, 0, 0], np.float32), 24))
    w, b = _linear(r, 3, h)
    p['learned_motion.linear_out.weight'] = w
    p['learned_motion.linear_out.bias'] = b
    p['learned_betas'] = _f32(np.zeros((1, 10)))
    k = args.monotonic_network_n_nodes
    for v in range(num_views):
        if args.phase_init == 'linear':
            s = np.linspace(0.0, 1.0, k)
        else:
            s = r.uniform(0.0, 1.0, k)
        p['phase_networks.%d.shifts' % v] = _f32(np.clip(s, 0, 1))
        p['phase_networks.%d.scales' % v] = _f32(np.full(k, 15.0))
    if args.phase_rbf_dim > 0:
        p['phase_rbf.log_sigmas'] = _f32(np.zeros(args.phase_rbf_dim))
        p['phase_rbf.centres'] = _f32(np.linspace(0.0, 1.0, args.phase_rbf_dim)[:, None])
    return p

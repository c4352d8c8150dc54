"""Size-independent properties at BASELINE.json's full sizes (N = 4096 poses per GPU, 6890 vertices), where the CPU
oracle is too slow to be the checker: rest pose, linearity / adjointness of the skinning kernels, idempotence of
`step(update=False)`, and the full-batch step of the custom-video configuration against the oracle at a small size."""
import numpy as np
import pytest
import torch

from nemo_cvpr2023_b200 import assets
from oracle import nemo_oracle as O
from tests.util import TOL_LOSS_REL, TOL_VERT_ABS, assert_grad, assert_rel, npy

pytestmark = pytest.mark.gpu
DEV = 'cuda'
N = 4096


@pytest.fixture(scope='module')
def ctx():
    import nemo_cvpr2023_b200.functional as Fn
    from nemo_cvpr2023_b200._lib import lib
    sm = assets.make_smpl_assets(0)
    return Fn, lib, Fn.SMPLConstants(sm, DEV), sm


def test_rest_pose_is_the_shaped_template(ctx):
    """lbs.py:212-252 at identity rotations: vertices == v_template + S.beta for every one of 4096 frames, joints ==
    J_regressor . v_shaped (identity chain)."""
    from nemo_cvpr2023_b200.hmr.smpl import SMPL
    Fn, lib, C, sm = ctx
    smpl = SMPL('synthetic')
    g = torch.Generator(device=DEV).manual_seed(0)
    betas = torch.randn(1, 10, device=DEV, generator=g)
    eye = torch.eye(3, device=DEV)
    out = smpl(betas=betas, body_pose=eye.expand(N, 23, 3, 3).contiguous(), global_orient=eye.expand(N, 1, 3, 3).contiguous(),
               pose2rot=False)
    v_shaped = torch.tensor(sm['v_template'], device=DEV).double() + torch.einsum(
        'vcl,l->vc', torch.tensor(sm['shapedirs'], device=DEV).double(), betas[0].double())
    assert out.vertices.shape == (N, 6890, 3)
    assert float((out.vertices.double() - v_shaped[None]).abs().max()) < TOL_VERT_ABS
    J = torch.tensor(sm['J_regressor'], device=DEV).double() @ v_shaped
    jm = np.asarray(sm['joint_map'])
    kin = [(i, int(j)) for i, j in enumerate(jm) if j < 24]
    got = out.joints.double()
    for i, j in kin[:8]:
        assert float((got[:, i] - J[j][None]).abs().max()) < TOL_VERT_ABS


def test_skinning_is_linear_and_its_backward_is_the_adjoint(ctx):
    """v = T(A) [base + PO; 1] is affine in PO and linear in A: for random directions the forward differences equal the
    inner products with the backward's outputs, <dV, V(PO + d) - V(PO)> = <dPO, d> and <dV, V(A + E) - V(A)> = <dA, E>
    (N = 4096: 28 M vertex-frames; float64 accumulation of the inner products)."""
    Fn, lib, C, sm = ctx
    g = torch.Generator(device=DEV).manual_seed(1)
    A = torch.randn(N, 24, 12, device=DEV, generator=g)
    PO = 0.01 * torch.randn(N, 3, Fn.VP, device=DEV, generator=g)
    PO[:, :, 6890:] = 0
    dV = torch.randn(N, 6890, 3, device=DEV, generator=g)
    d = 0.01 * torch.randn(N, 3, Fn.VP, device=DEV, generator=g)
    d[:, :, 6890:] = 0
    E = torch.randn(N, 24, 12, device=DEV, generator=g)
    V0, V1, V2 = (torch.empty(N, 6890, 3, device=DEV) for _ in range(3))
    lib.skin_fwd(A, PO.view(N, -1), C.vt, C.Wt, V0, N)
    lib.skin_fwd(A, (PO + d).view(N, -1), C.vt, C.Wt, V1, N)
    lib.skin_fwd(A + E, PO.view(N, -1), C.vt, C.Wt, V2, N)
    dPO, dA = torch.empty(N, 3 * Fn.VP, device=DEV), torch.empty(N, 24, 12, device=DEV)
    lib.skin_bwd(A, PO.view(N, -1), C.vt, C.Wt, dV, None, 0.0, dPO, dA, N)
    lhs = float((dV.double() * (V1.double() - V0.double())).sum())
    rhs = float((dPO.view(N, 3, -1).double() * d.double()).sum())
    assert abs(lhs - rhs) <= 2e-4 * max(abs(lhs), abs(rhs)), (lhs, rhs)
    lhs = float((dV.double() * (V2.double() - V0.double())).sum())
    rhs = float((dA.double() * E.double()).sum())
    assert abs(lhs - rhs) <= 2e-4 * max(abs(lhs), abs(rhs)), (lhs, rhs)
    assert float(dPO.view(N, 3, -1)[:, :, 6890:].abs().max()) == 0.0


def test_fused_l1_equals_the_two_meshes(ctx):
    """the fused vertex L1 (two meshes in one accumulator) == mean |V_a - V_b| of two separate forward launches, and
    its sign bits reproduce sign(V_a - V_b) through the backward (N = 4096)."""
    Fn, lib, C, sm = ctx
    g = torch.Generator(device=DEV).manual_seed(2)
    A, A2 = (torch.randn(N, 24, 12, device=DEV, generator=g) for _ in range(2))
    PO, PO2 = (0.01 * torch.randn(N, 3 * Fn.VP, device=DEV, generator=g) for _ in range(2))
    Va, Vb = torch.empty(N, 6890, 3, device=DEV), torch.empty(N, 6890, 3, device=DEV)
    lib.skin_fwd(A, PO, C.vt, C.Wt, Va, N)
    lib.skin_fwd(A2, PO2, C.vt, C.Wt, Vb, N)
    signs = torch.zeros(N, 3, 2, Fn.VP // 32, dtype=torch.int32, device=DEV)
    l1 = torch.zeros(1, device=DEV)
    lib.skin_l1_fwd(A, PO, A2, PO2, C.vt, C.Wt, signs, l1, N)
    ref = float((Va.double() - Vb.double()).abs().sum())
    assert abs(float(l1) - ref) <= 1e-5 * ref
    dPO_s, dA_s = torch.empty(N, 3 * Fn.VP, device=DEV), torch.empty(N, 24, 12, device=DEV)
    lib.skin_bwd(A, PO, C.vt, C.Wt, None, signs, 0.5, dPO_s, dA_s, N)
    dPO_v, dA_v = torch.empty_like(dPO_s), torch.empty_like(dA_s)
    lib.skin_bwd(A, PO, C.vt, C.Wt, (0.5 * torch.sign(Va - Vb)).contiguous(), None, 0.0, dPO_v, dA_v, N)
    assert_grad(npy(dPO_s), npy(dPO_v), what='dPO signs vs dV')
    assert_grad(npy(dA_s), npy(dA_v), what='dA signs vs dV')


def test_bench_size_step_is_idempotent_without_update():
    """NemoV2.step at the bench size (8 views x 512 frames): update=False leaves every parameter untouched and
    repeats its losses; update=True changes them and lowers nothing to NaN."""
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    V, Fr = 8, 512
    args = assets.default_args(batch_size=V * Fr)
    m = nmm.NemoV2(args, SyntheticMultiViewSequence(V, Fr, seed=1), DEV, init_params=assets.make_fit_params(args, V, seed=4))
    vi, fi = torch.arange(V).repeat_interleave(Fr), torch.arange(Fr).repeat(V)
    before = {k: v.clone() for k, v in m.engine.named_params().items()}
    a, info = m.step(vi, fi, update=False)
    b, _ = m.step(vi, fi, update=False)
    for k, v in m.engine.named_params().items():
        assert torch.equal(v, before[k]), k
    for k in a:
        assert np.isfinite(a[k]) and abs(float(a[k]) - float(b[k])) <= 1e-6 * abs(float(a[k])) + 1e-12, k
    assert info['loss_all'].shape == (V * Fr, 25, 2)
    m.step(vi, fi, update=True)
    changed = [k for k, v in m.engine.named_params().items() if not torch.equal(v, before[k])]
    assert 'learned_cameras' in changed and any(k.startswith('learned_motion.') for k in changed)
    assert all(torch.isfinite(v).all() for v in m.engine.named_params().values())


def test_full_batch_custom_video_step_matches_oracle():
    """configs[4] shape at oracle size: NemoV3, full_batch=True (every (view, frame) pair), weight_3d_loss = 1000,
    instance codes -- losses and gradients against the CPU oracle."""
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    V, Fr = 3, 8
    args = assets.default_args(h_dim=40, phase_rbf_dim=10, instance_code_size=5, monotonic_network_n_nodes=12,
                               batch_size=16, weight_3d_loss=1000.0, weight_instance_loss=0.05)
    params = assets.make_fit_params(args, V, seed=41)
    data = assets.make_fit_data(V, Fr, seed=42)
    r = np.random.RandomState(43)
    data['points2d_gt_all'][..., 0] = 540 + 150 * r.standard_normal((V, Fr, 25))
    data['points2d_gt_all'][..., 1] = 960 + 150 * r.standard_normal((V, Fr, 25))
    orc = O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                      O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                      version=3).build_optimizers()
    vi, fi = torch.arange(V).repeat_interleave(Fr), torch.arange(Fr).repeat(V)
    ref_ld, _ = orc.step(vi, fi)
    m = nmm.NemoV3(args, SyntheticMultiViewSequence(data=data), DEV, init_params=params)
    ld, info = m.step(None, None, full_batch=True)
    assert info['view_idx'].numel() == V * Fr
    for k in ['kp_loss', 'gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'instance_loss', 'loss_3d', 'total_loss']:
        assert_rel(ld[k], ref_ld[k], TOL_LOSS_REL, k)
    grads = m.engine.named_grads()
    for k, p in orc.p.items():
        if p.grad is not None and k != 'learned_betas':
            assert_grad(npy(grads[k]), npy(p.grad), what=k)

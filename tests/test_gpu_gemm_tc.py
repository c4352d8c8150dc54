"""tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) against float64 numpy and the CUDA-core kernel: every operand layout
(K-major / MN-major A and B), ragged M/N/K, all fused epilogues, split-K, strided C."""
import numpy as np
import pytest
import torch

from .util import npy

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def G(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device=DEV)


@pytest.fixture(scope='module')
def lib():
    from nemo_cvpr2023_b200._lib import lib as L
    return L


def pad_ld(x, mult=4):
    """(rows, cols) -> storage with a row stride that is a multiple of `mult` floats; returns (tensor, ld)."""
    r, c = x.shape
    ld = (c + mult - 1) // mult * mult
    buf = np.zeros((r, ld), np.float32)
    buf[:, :c] = x
    return G(buf), ld


SHAPES = [(128, 128, 32), (4097, 1000, 1000), (300, 207, 6912), (513, 1000, 104), (147, 1000, 4097), (77, 147, 1000),
          (4096, 2 * 3456, 207), (33, 9, 40)]


@pytest.mark.parametrize('ta,tb', [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm_tc_layouts(lib, ta, tb):
    r = np.random.RandomState(11 + 2 * ta + tb)
    for (M, N, K) in SHAPES:
        A = r.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        B = r.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
        Ad, lda = pad_ld(A)
        Bd, ldb = pad_ld(B)
        C = torch.full((M, N), 7.0, device=DEV)
        lib.gemm_tc(Ad, Bd, None, C, M, N, K, lda, ldb, N, ta, tb, 0, None, 0, 0.0, 0, 1)
        torch.cuda.synchronize()
        err = np.abs(npy(C) - ref).max() / np.abs(ref).max()
        # 3xTF32: ~2^-21 relative per product, fp32 accumulation
        assert err < 4e-6, 'gemm_tc ta=%d tb=%d %s: rel-max err %.3e' % (ta, tb, (M, N, K), err)


def test_gemm_tc_epilogues_and_splitk(lib):
    r = np.random.RandomState(3)
    M, N, K = 1000, 207, 2080
    A = r.standard_normal((M, K)).astype(np.float32)
    B = r.standard_normal((N, K)).astype(np.float32)
    bias = r.standard_normal(N).astype(np.float32)
    gate = r.standard_normal((M, N)).astype(np.float32)
    C0 = r.standard_normal((M, 216)).astype(np.float32)
    acc = A.astype(np.float64) @ B.T.astype(np.float64)
    Ad, Bd = G(A), G(B)
    # bias + leaky relu
    C = torch.empty(M, N, device=DEV)
    lib.gemm_tc(Ad, Bd, G(bias), C, M, N, K, K, K, N, 0, 1, 2, None, 0, 0.0, 0, 0)
    ref = acc + bias
    ref = np.where(ref > 0, ref, 0.01 * ref)
    assert np.abs(npy(C) - ref).max() / np.abs(ref).max() < 4e-6
    # bias + relu
    lib.gemm_tc(Ad, Bd, G(bias), C, M, N, K, K, K, N, 0, 1, 1, None, 0, 0.0, 0, 0)
    assert np.abs(npy(C) - np.maximum(acc + bias, 0)).max() / np.abs(ref).max() < 4e-6
    # gate (activation backward) + accumulate into a strided C, with and without split-K
    for splits in (1, 5):
        Cs = G(C0)
        lib.gemm_tc(Ad, Bd, None, Cs, M, N, K, K, K, 216, 0, 1, 0, G(gate), N, 0.01, 1, splits)
        ref = C0.astype(np.float64).copy()
        ref[:, :N] += acc * np.where(gate > 0, 1.0, 0.01)
        assert np.abs(npy(Cs) - ref).max() / np.abs(ref).max() < 4e-6, 'splits=%d' % splits
    # split-K overwrite (the kernel zeroes the N used columns only) + bias
    Cs = G(C0)
    lib.gemm_tc(Ad, Bd, G(bias), Cs, M, N, K, K, K, 216, 0, 1, 0, None, 0, 0.0, 0, 4)
    ref = C0.astype(np.float64).copy()
    ref[:, :N] = acc + bias
    assert np.abs(npy(Cs) - ref).max() / np.abs(ref).max() < 4e-6
    # heuristic split (wgrad-like: small M x N, deep K)
    M2, N2, K2 = 147, 1000, 4097
    A2 = r.standard_normal((K2, 148)).astype(np.float32)
    B2 = r.standard_normal((K2, N2)).astype(np.float32)
    C = torch.empty(M2, N2, device=DEV)
    lib.gemm_tc(G(A2), G(B2), None, C, M2, N2, K2, 148, N2, N2, 1, 0, 0, None, 0, 0.0, 0, 0)
    ref = A2[:, :M2].T.astype(np.float64) @ B2.astype(np.float64)
    assert np.abs(npy(C) - ref).max() / np.abs(ref).max() < 4e-6


def test_gemm_dispatch_agrees_with_cuda_core_kernel(lib):
    """nemo_b200_gemm (tensor-core path for these shapes) vs nemo_b200_gemm_simt on the same inputs."""
    r = np.random.RandomState(9)
    M, N, K = 512, 20736, 208
    A, B = G(r.standard_normal((M, K))), G(0.001 * r.standard_normal((K, N)))
    C1, C2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    lib.gemm(A, B, None, C1, M, N, K, K, N, N, 0, 0, 0, None, 0, 0.0, 0)
    lib.gemm_simt(A, B, None, C2, M, N, K, K, N, N, 0, 0, 0, None, 0, 0.0, 0)
    assert (C1 - C2).abs().max().item() < 4e-6 * C2.abs().max().item()


def test_gemm_tc_tma_epilogue(lib):
    """16-byte addressable C and gate: the tile leaves through the swizzled shared-memory box and TMA store /
    reduce-add (ragged M and N tails are clipped by the tensor map); gate box fetched by TMA."""
    r = np.random.RandomState(21)
    M, N, K = 1030, 1000, 520
    A = r.standard_normal((M, K)).astype(np.float32)
    B = r.standard_normal((N, K)).astype(np.float32)
    bias = r.standard_normal(N).astype(np.float32)
    gate = r.standard_normal((M, N)).astype(np.float32)
    C0 = r.standard_normal((M, 1004)).astype(np.float32)
    acc = A.astype(np.float64) @ B.T.astype(np.float64)
    Ad, Bd, Gd = G(A), G(B), G(gate)
    scale = np.abs(acc).max()
    # store: bias + relu
    C = torch.full((M, N), 7.0, device=DEV)
    lib.gemm_tc(Ad, Bd, G(bias), C, M, N, K, K, K, N, 0, 1, 1, None, 0, 0.0, 0, 1)
    assert np.abs(npy(C) - np.maximum(acc + bias, 0)).max() / scale < 4e-6
    # store: gate (ReLU backward, slope 0) into a strided C; the columns beyond N stay untouched
    Cs = G(C0)
    lib.gemm_tc(Ad, Bd, None, Cs, M, N, K, K, K, 1004, 0, 1, 0, Gd, N, 0.0, 0, 1)
    ref = C0.astype(np.float64).copy()
    ref[:, :N] = acc * (gate > 0)
    assert np.abs(npy(Cs) - ref).max() / scale < 4e-6
    # reduce-add: accumulate, leaky gate, with and without split-K
    for splits in (1, 3):
        Cs = G(C0)
        lib.gemm_tc(Ad, Bd, None, Cs, M, N, K, K, K, 1004, 0, 1, 0, Gd, N, 0.01, 1, splits)
        ref = C0.astype(np.float64).copy()
        ref[:, :N] += acc * np.where(gate > 0, 1.0, 0.01)
        assert np.abs(npy(Cs) - ref).max() / scale < 4e-6, 'splits=%d' % splits
    # split-K overwrite + bias
    Cs = G(C0)
    lib.gemm_tc(Ad, Bd, G(bias), Cs, M, N, K, K, K, 1004, 0, 1, 0, None, 0, 0.0, 0, 4)
    ref = C0.astype(np.float64).copy()
    ref[:, :N] = acc + bias
    assert np.abs(npy(Cs) - ref).max() / scale < 4e-6
    # a persistent CTA walking many tiles (36 per SM) reuses its boxes: pose-blend shape, MN-major B
    M2, N2, K2 = 2048, 20736, 207
    A2 = r.standard_normal((M2, 208)).astype(np.float32)
    B2 = (0.01 * r.standard_normal((K2, N2))).astype(np.float32)
    C = torch.empty(M2, N2, device=DEV)
    lib.gemm_tc(G(A2), G(B2), None, C, M2, N2, K2, 208, N2, N2, 0, 0, 0, None, 0, 0.0, 0, 0)
    ref = A2[:, :K2].astype(np.float64) @ B2.astype(np.float64)
    assert np.abs(npy(C) - ref).max() / np.abs(ref).max() < 4e-6


@pytest.mark.parametrize('ta,tb', [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm_tc2_layouts(lib, ta, tb):
    """second-generation kernel (128 x 256 x 16 tiles, 64-byte swizzled K-major rows): every layout, ragged tails"""
    r = np.random.RandomState(31 + 2 * ta + tb)
    for (M, N, K) in [(128, 256, 16), (4097, 1000, 1000), (300, 207, 6912), (513, 1000, 104), (1000, 1000, 4097),
                      (2048, 2 * 3456, 207), (77, 300, 1000), (33, 9, 40)]:
        A = r.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        B = r.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
        Ad, lda = pad_ld(A)
        Bd, ldb = pad_ld(B)
        # N % 4 == 0: C may sit inside a wider matrix, its neighbours stay untouched; else only trailing row padding
        ldc = N + 8 if N % 4 == 0 else (N + 3) // 4 * 4
        C = torch.full((M + 1, ldc), 7.0, device=DEV)
        lib.gemm_tc2(Ad, Bd, None, C, M, N, K, lda, ldb, ldc, ta, tb, 0, None, 0, 0.0, 0, 1)
        torch.cuda.synchronize()
        err = np.abs(npy(C)[:M, :N] - ref).max() / np.abs(ref).max()
        assert err < 4e-6, 'gemm_tc2 ta=%d tb=%d %s: rel-max err %.3e' % (ta, tb, (M, N, K), err)
        assert float((C[M:] - 7.0).abs().max()) == 0.0
        if N % 4 == 0:
            assert float((C[:, N:] - 7.0).abs().max()) == 0.0


def test_gemm_tc2_epilogues_and_splitk(lib):
    r = np.random.RandomState(41)
    M, N, K = 1030, 1000, 520
    A = r.standard_normal((M, K)).astype(np.float32)
    B = r.standard_normal((N, K)).astype(np.float32)
    bias = r.standard_normal(N).astype(np.float32)
    gate = r.standard_normal((M, N)).astype(np.float32)
    C0 = r.standard_normal((M, 1004)).astype(np.float32)
    acc = A.astype(np.float64) @ B.T.astype(np.float64)
    Ad, Bd, Gd = G(A), G(B), G(gate)
    scale = np.abs(acc).max()
    C = torch.full((M, N), 7.0, device=DEV)
    lib.gemm_tc2(Ad, Bd, G(bias), C, M, N, K, K, K, N, 0, 1, 1, None, 0, 0.0, 0, 1)
    assert np.abs(npy(C) - np.maximum(acc + bias, 0)).max() / scale < 4e-6
    lib.gemm_tc2(Ad, Bd, G(bias), C, M, N, K, K, K, N, 0, 1, 2, None, 0, 0.0, 0, 1)
    ref = acc + bias
    assert np.abs(npy(C) - np.where(ref > 0, ref, 0.01 * ref)).max() / scale < 4e-6
    Cs = G(C0)
    lib.gemm_tc2(Ad, Bd, None, Cs, M, N, K, K, K, 1004, 0, 1, 0, Gd, N, 0.0, 0, 1)
    ref = C0.astype(np.float64).copy()
    ref[:, :N] = acc * (gate > 0)
    assert np.abs(npy(Cs) - ref).max() / scale < 4e-6
    for splits in (1, 3):
        Cs = G(C0)
        lib.gemm_tc2(Ad, Bd, None, Cs, M, N, K, K, K, 1004, 0, 1, 0, Gd, N, 0.01, 1, splits)
        ref = C0.astype(np.float64).copy()
        ref[:, :N] += acc * np.where(gate > 0, 1.0, 0.01)
        assert np.abs(npy(Cs) - ref).max() / scale < 4e-6, 'splits=%d' % splits
    Cs = G(C0)
    lib.gemm_tc2(Ad, Bd, G(bias), Cs, M, N, K, K, K, 1004, 0, 1, 0, None, 0, 0.0, 0, 4)
    ref = C0.astype(np.float64).copy()
    ref[:, :N] = acc + bias
    assert np.abs(npy(Cs) - ref).max() / scale < 4e-6
    # heuristic split-K (dgrad-like: one tile column, deep K) and many tiles per CTA (pose-blend forward)
    M2, N2, K2 = 2048, 207, 20736
    A2 = (0.05 * r.standard_normal((M2, K2))).astype(np.float32)
    B2 = r.standard_normal((N2, K2)).astype(np.float32)
    C = torch.empty(M2, 208, device=DEV)
    lib.gemm_tc2(G(A2), G(B2), None, C, M2, N2, K2, K2, K2, 208, 0, 1, 0, None, 0, 0.0, 0, 0)
    ref = A2.astype(np.float64) @ B2.T.astype(np.float64)
    assert np.abs(npy(C)[:, :N2] - ref).max() / np.abs(ref).max() < 4e-6


@pytest.mark.parametrize('M,N,K', [(4096, 20736, 207), (37, 20736, 207), (300, 1000, 64), (129, 260, 33)])
def test_gemm_h16_pose_blend(lib, M, N, K):
    """The fp16-pair pose-blend GEMM (csrc/gemm_h16.cu: tcgen05 kind::f16, constant operand split once) against
    float64: A = rotation-like pose features in [-2, 2] with a wide spread of magnitudes, B = a posedirs-like constant;
    ragged M / N / K."""
    r = np.random.RandomState(M + N + K)
    A = np.clip(r.standard_normal((M, K)) * 10.0 ** r.uniform(-4, 0, size=(M, K)), -2, 2).astype(np.float32)
    A[:, ::17] = 0.0
    B = (1e-3 * r.standard_normal((K, N)) * 10.0 ** r.uniform(-3, 0, size=(K, N))).astype(np.float32)
    Ad, lda = pad_ld(A)
    ldb = (K + 31) // 32 * 32
    pmax = float(np.abs(B).max())
    bs = 2.0 ** (13 - int(np.floor(np.log2(pmax))))
    x = torch.zeros(N, ldb, device=DEV)
    x[:, :K] = G(B).t() * bs
    Bh = x.half().contiguous()
    Bl = (x - Bh.float()).half().contiguous()
    ldc = (N + 3) // 4 * 4
    C = torch.full((M, ldc), float('nan'), device=DEV)
    lib.gemm_h16(Ad, Bh, Bl, None, C, M, N, K, lda, ldb, ldc, float(2 ** 13), 1.0 / (2 ** 13 * bs))
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = npy(C)[:, :N]
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= 3e-6, 'rel-max error %.2e' % err
    # row-wise: rows of tiny features keep their relative accuracy
    rows = np.abs(ref).max(1) > 0
    rerr = (np.abs(got - ref).max(1)[rows] / np.abs(ref).max(1)[rows]).max()
    assert rerr <= 2e-5, 'worst row-relative error %.2e' % rerr


def test_gemm_h16_split_k_dgrad(lib):
    """the same kernel on the pose-blend dgrad shape: K = 20 736 cut over the SMs (split-K, TMA reduce-add into the zeroed
    C), N = 207 in a 208-wide buffer, A = a bounded upstream gradient."""
    M, N, K = 1031, 207, 20736
    r = np.random.RandomState(5)
    bound = 3.7e-7
    A = (bound * np.clip(r.standard_normal((M, K)), -1, 1) * (r.uniform(size=(M, K)) > 0.3)).astype(np.float32)
    B = (1e-3 * r.standard_normal((N, K))).astype(np.float32)            # posedirs' own orientation: (207, 3VP), K-major
    bs = 2.0 ** (13 - int(np.floor(np.log2(float(np.abs(B).max())))))
    y = G(B) * bs
    Bh = y.half().contiguous()
    Bl = (y - Bh.float()).half().contiguous()
    a_scale = 2.0 ** (13 - int(np.ceil(np.log2(bound))))
    C = torch.full((M, 208), float('nan'), device=DEV)
    lib.gemm_h16(G(A), Bh, Bl, None, C, M, N, K, K, K, 208, float(a_scale), 1.0 / (a_scale * bs))
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    got = npy(C)[:, :N]
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= 3e-6, 'rel-max error %.2e' % err

"""Device time of the GEMM shapes of one fit step: tcgen05 3xTF32 kernel vs the CUDA-core kernel.
    python tools/bench_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemo_cvpr2023_b200._lib import lib  # noqa: E402

dev = 'cuda'
SHAPES = [  # name, M, N, K, transA, transB
    ('mlp fwd 1000x1000', 4097, 1000, 1000, 0, 1),
    ('mlp dgrad', 4097, 1000, 1000, 0, 0),
    ('mlp wgrad', 1000, 1000, 4097, 1, 0),
    ('head fwd', 4097, 147, 1000, 0, 1),
    ('head wgrad', 147, 1000, 4097, 1, 0),
    ('pose-blend fwd', 4096, 20736, 207, 0, 0),
    ('pose-blend dgrad', 4096, 207, 20736, 0, 1),
]


def timeit(fn, iters=10):
    if os.environ.get('NCU_ONCE'):   # under ncu: one launch per case
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, M, N, K, ta, tb in SHAPES:
    r4 = lambda x: (x + 3) // 4 * 4
    A = torch.randn((K, r4(M)) if ta else (M, r4(K)), device=dev)
    B = torch.randn((N, r4(K)) if tb else (K, r4(N)), device=dev)
    C = torch.empty(M, N, device=dev)
    lda, ldb = A.shape[1], B.shape[1]
    t_tc = timeit(lambda: lib.gemm_tc(A, B, None, C, M, N, K, lda, ldb, N, ta, tb, 0, None, 0, 0.0, 0, 0))
    t_t2 = timeit(lambda: lib.gemm_tc2(A, B, None, C, M, N, K, lda, ldb, N, ta, tb, 0, None, 0, 0.0, 0, 0)) if N % 4 == 0 else float('nan')
    t_si = timeit(lambda: lib.gemm_simt(A, B, None, C, M, N, K, lda, ldb, N, ta, tb, 0, None, 0, 0.0, 0), iters=3)
    fl = 2.0 * M * N * K
    lib.gemm_tc(A, B, None, C, M, N, K, lda, ldb, N, ta, tb, 0, None, 0, 0.0, 0, 0)
    Ad = (A[:, :M].T if ta else A[:, :K]).double()
    Bd = (B[:, :K].T if tb else B[:, :N]).double()
    ref = Ad @ Bd
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    lib.gemm_simt(A, B, None, C, M, N, K, lda, ldb, N, ta, tb, 0, None, 0, 0.0, 0)
    err_s = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    print('%-20s M=%5d N=%5d K=%5d  tc2 %.3f ms (%.1f)  tc %.3f ms (%.1f TFLOP/s fp32-equivalent)   simt %.3f ms (%.1f TFLOP/s)  rel-max err tc %.2e simt %.2e' % (
        name, M, N, K, t_t2, fl / t_t2 / 1e9, t_tc, fl / t_tc / 1e9, t_si, fl / t_si / 1e9, err, err_s))

# the fp16-pair pose-blend kernel (csrc/gemm_h16.cu), same shape as 'pose-blend fwd'
M, N, K = 4096, 20736, 207
A = torch.randn(M, 208, device=dev).clamp_(-2, 2)
B = 1e-3 * torch.randn(K, N, device=dev)
bs = 2.0 ** (13 - int(torch.floor(torch.log2(B.abs().max()))))
x = torch.zeros(N, 224, device=dev)
x[:, :K] = B.t() * bs
Bh = x.half().contiguous()
Bl = (x - Bh.float()).half().contiguous()
C = torch.empty(M, N, device=dev)
t = timeit(lambda: lib.gemm_h16(A, Bh, Bl, None, C, M, N, K, 208, 224, N, float(2 ** 13), 1.0 / (2 ** 13 * bs)))
ref = A[:, :K].double() @ B.double()
err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
print('%-20s M=%5d N=%5d K=%5d  fp16-pair kernel %.3f ms (%.1f TFLOP/s fp32-equivalent)  rel-max err %.2e' % (
    'pose-blend fwd h16', M, N, K, t, 2.0 * M * N * K / t / 1e9, err))

# ... and on the dgrad shape (K = 20 736, split-K)
M, N, K = 4096, 207, 20736
A = (1e-6 * torch.randn(M, K, device=dev)).clamp_(-4e-6, 4e-6)
B = 1e-3 * torch.randn(N, K, device=dev)
y = B * bs
Bh = y.half().contiguous()
Bl = (y - Bh.float()).half().contiguous()
C = torch.empty(M, 208, device=dev)
asc = 2.0 ** (13 + 18)
t = timeit(lambda: lib.gemm_h16(A, Bh, Bl, None, C, M, N, K, K, K, 208, asc, 1.0 / (asc * bs)))
ref = A.double() @ B.double().t()
err = ((C[:, :N].double() - ref).abs().max() / ref.abs().max()).item()
print('%-20s M=%5d N=%5d K=%5d  fp16-pair kernel %.3f ms (%.1f TFLOP/s fp32-equivalent)  rel-max err %.2e' % (
    'pose-blend dgrad h16', M, N, K, t, 2.0 * M * N * K / t / 1e9, err))

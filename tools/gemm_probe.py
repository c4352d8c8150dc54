"""What bounds k_gemm_tc's main loop?  Cycles per k-block (128 x 128 x 32, 12 tcgen05.mma) as a function of
  - how many CTAs are active (M = 1024 .. 8192 at N = K = 1024: 64, 128, 148 CTAs): a shared resource (L2 -> SM
    bandwidth) shows as a period that grows with the CTA count, a per-SM resource does not;
  - how often the TMEM partial is folded into registers (NEMO_B200_GEMM_FLUSH, set per process).
    python tools/gemm_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemo_cvpr2023_b200._lib import lib  # noqa: E402

dev = 'cuda'
CLK = 1.965e9


def t_ms(fn, iters=20):
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


print('flush env:', os.environ.get('NEMO_B200_GEMM_FLUSH', '(heuristic)'))
for (M, N, K, ta, tb) in [(1024, 1024, 1024, 0, 1), (2048, 1024, 1024, 0, 1), (4096, 1024, 1024, 0, 1), (8192, 1024, 1024, 0, 1),
                          (16384, 1024, 1024, 0, 1), (4096, 1024, 4096, 0, 1), (4096, 1024, 1024, 0, 0), (1024, 1024, 4096, 1, 0),
                          (4096, 20736, 207, 0, 0)]:
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((N, K) if tb else (K, N), device=dev)
    C = torch.empty(M, N, device=dev)
    lda, ldb = A.shape[1], B.shape[1]
    ms = t_ms(lambda: lib.gemm_tc(A, B, None, C, M, N, K, lda, ldb, N, ta, tb, 0, None, 0, 0.0, 0, 1))
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    ctas = min(tiles, 148)
    kb = (K + 31) // 32
    per_sm = -(-tiles // ctas) * kb
    print('M=%6d N=%6d K=%5d %s%s  %.4f ms  %6.1f TFLOP/s fp32-eq  tiles %5d ctas %3d  k-blocks on the busiest SM %5d -> %6.0f cycles per k-block, '
          'L2->SM %.2f TB/s' % (M, N, K, 'T' if ta else 'N', 'T' if tb else 'N', ms, 2.0 * M * N * K / ms / 1e9, tiles, ctas, per_sm,
                               ms * 1e-3 * CLK / per_sm, tiles * kb * 32768 / (ms * 1e-3) / 1e12))

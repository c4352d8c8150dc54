"""CTA 0's timeline (cycles since kernel entry) of k_gemm_tc for a few shapes:  NEMO_B200_GEMM_TIMING=1 python tools/gemm_timeline.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemo_cvpr2023_b200._lib import lib  # noqa: E402

for (M, N, K, ta, tb) in [(4096, 512, 72, 0, 1), (4096, 512, 512, 0, 1), (4097, 1000, 1000, 0, 1), (4097, 1000, 1000, 0, 0),
                          (1000, 1000, 4097, 1, 0), (4096, 20736, 207, 0, 0), (4096, 207, 20736, 0, 1)]:
    r4 = lambda x: (x + 3) // 4 * 4
    A = torch.randn((K, r4(M)) if ta else (M, r4(K)), device='cuda')
    B = torch.randn((N, r4(K)) if tb else (K, r4(N)), device='cuda')
    C = torch.empty(M, r4(N), device='cuda')
    for _ in range(2):
        lib.gemm(A, B, None, C, M, N, K, A.shape[1], B.shape[1], r4(N), ta, tb, 0, None, 0, 0.0, 0)
    torch.cuda.synchronize()

import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemo_cvpr2023_b200._lib import lib
r = np.random.RandomState(1)
for (M, N, K, ta, tb) in [(300, 207, 6912, 0, 1), (33, 9, 40, 0, 1), (300, 207, 64, 0, 1), (128, 100, 64, 0, 1), (300, 207, 64, 0, 0)]:
    A = torch.tensor(r.standard_normal((K, (M+3)//4*4) if ta else (M, (K+3)//4*4)).astype(np.float32), device='cuda')
    B = torch.tensor(r.standard_normal((N, (K+3)//4*4) if tb else (K, (N+3)//4*4)).astype(np.float32), device='cuda')
    ldc = (N + 3) // 4 * 4 + 4
    for name, fn in (('tc2', lib.gemm_tc2), ('tc1', lib.gemm_tc)):
        C = torch.full((M + 2, ldc), 7.0, device='cuda')
        fn(A, B, None, C, M, N, K, A.shape[1], B.shape[1], ldc, ta, tb, 0, None, 0, 0.0, 0, 1)
        torch.cuda.synchronize()
        bad = (C[:M, N:] != 7.0).nonzero()
        badr = (C[M:] != 7.0).nonzero()
        print(name, (M, N, K, ta, tb), 'cols beyond N touched:', bad.shape[0], bad[:5].tolist(), C[:M, N:][C[:M, N:] != 7.0][:5].tolist(),
              'rows beyond M touched:', badr.shape[0])

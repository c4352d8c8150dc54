"""K7 alone: device time and algorithmic HBM GB/s of the skinning kernels at the bench batch (N = 4096).
    python tools/bench_skin.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemo_cvpr2023_b200 import assets, functional as Fn  # noqa: E402
from nemo_cvpr2023_b200._lib import lib  # noqa: E402

dev = 'cuda'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
STREAM = 6890 * 3 * 4
C = Fn.SMPLConstants(assets.make_smpl_assets(0), dev)
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(N, 24, 12, device=dev, generator=g)
A2 = torch.randn(N, 24, 12, device=dev, generator=g)
PO = 0.01 * torch.randn(N, 3 * Fn.VP, device=dev, generator=g)
PO2 = 0.01 * torch.randn(N, 3 * Fn.VP, device=dev, generator=g)
dV = torch.randn(N, Fn.NV, 3, device=dev, generator=g)
ro = C.readout
nxr = ro['np'] + ro['nx']
dJX = torch.randn(N, nxr, 3, device=dev, generator=g)
V = torch.empty(N, Fn.NV, 3, device=dev)
JX = torch.empty(N, nxr, 3, device=dev)
dPO, dA = torch.empty_like(PO), torch.empty(N, 24, 12, device=dev)
signs = torch.zeros(N, 3, 2, Fn.VP // 32, dtype=torch.int32, device=dev)
l1 = torch.zeros(1, device=dev)


def timeit(fn, iters=10):
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


cases = [
    ('fwd  V + all joints (np=%d nx=%d)' % (ro['np'], ro['nx']), 2,
     lambda: lib.skin_fwd(A, PO, C.vt, C.Wt, ro['Jx'], ro['nx'], ro['slot'], ro['np'], V, JX, N)),
    ('fwd  V only', 2, lambda: lib.skin_fwd(A, PO, C.vt, C.Wt, None, 0, None, 0, V, None, N)),
    ('fwd  joints only (no V, nx=1)', 1,
     lambda: lib.skin_fwd(A, PO, C.vt, C.Wt, ro['Jx'], 1, ro['slot'], ro['np'], None, JX, N)),
    ('l1   two meshes + sign bits', 2, lambda: lib.skin_l1_fwd(A, PO, A2, PO2, C.vt, C.Wt, signs, l1, N)),
    ('bwd  dV + dJX -> dPO, dA', 3,
     lambda: lib.skin_bwd(A, PO, C.vt, C.Wt, C.W2, ro['Jx'], ro['nx'], ro['slot'], ro['np'], dV, dJX, None, 0.0, dPO, dA, N)),
    ('bwd  dJX only (nx=1) -> dPO, dA', 2,
     lambda: lib.skin_bwd(A, PO, C.vt, C.Wt, C.W2, ro['Jx'], 1, ro['slot'], ro['np'], None, dJX, None, 0.0, dPO, dA, N)),
    ('bwd  signs -> dPO, dA', 2,
     lambda: lib.skin_bwd(A, PO, C.vt, C.Wt, C.W2, None, 0, None, 0, None, None, signs, 1e-3, dPO, dA, N)),
]
for name, streams, fn in cases:
    t = timeit(fn)
    print('%-40s %.3f ms   %7.1f GB/s algorithmic (%d streams/frame)' % (name, t, streams * STREAM * N / t / 1e6, streams))

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
V = torch.empty(N, Fn.NV, 3, device=dev)
dPO, dAs = torch.empty_like(PO), torch.empty(N, 24, 12, device=dev)
signs = torch.zeros(N, 3, 2, Fn.VP // 32, dtype=torch.int32, device=dev)
l1 = torch.zeros(1, device=dev)
ro = C.readout
nq = ro['np'] + ro['nx']
Fp = torch.randn(N, 208, device=dev, generator=g)
Z = torch.empty(N, ro['ldz'], device=dev)
dZ = torch.empty_like(Z)
JX = torch.empty(N, nq, 3, device=dev)
dJX = torch.randn(N, nq, 3, device=dev, generator=g)
dA = torch.empty(N, 24, 12, device=dev)


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


def readout_fwd():
    Fn.gemm(Fp, ro['MZ'], Z, N, ro['ldz'], 207, 208, ro['ldz'], ro['ldz'], bias=ro['Z0'])
    lib.readout_fwd(A, Z, ro['ldz'], ro['Wp'], ro['np'], ro['Sj'], ro['nx'], JX, N)


cases = [
    ('fwd  V', 2, lambda: lib.skin_fwd(A, PO, C.vt, C.Wt, V, N)),
    ('l1   two meshes + sign bits', 2, lambda: lib.skin_l1_fwd(A, PO, A2, PO2, C.vt, C.Wt, signs, l1, N)),
    ('bwd  dV -> dPO, dA', 3, lambda: lib.skin_bwd(A, PO, C.vt, C.Wt, dV, None, 0.0, dPO, dAs, N)),
    ('bwd  signs -> dPO, dA', 2, lambda: lib.skin_bwd(A, PO, C.vt, C.Wt, None, signs, 1e-3, dPO, dAs, N)),
    ('read-outs fwd (np=%d nx=%d): Z GEMM + kernel' % (ro['np'], ro['nx']), 0, readout_fwd),
    ('read-outs bwd: dA, dZ', 0,
     lambda: lib.readout_bwd(A, Z, ro['ldz'], ro['Wp'], ro['np'], ro['Sj'], ro['nx'], dJX, dA, 0, dZ, N)),
]
ONLY = sys.argv[2] if len(sys.argv) > 2 else ''
for name, streams, fn in cases:
    if ONLY and ONLY not in name:
        continue
    t = timeit(fn)
    print('%-48s %.3f ms   %7.1f GB/s algorithmic (%d streams/frame)' % (name, t, streams * STREAM * N / t / 1e6, streams))

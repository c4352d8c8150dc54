"""Per-entry-point device time of one fit step (CUDA events around every C-ABI call).
    python tools/profile_step.py [N_views] [N_frames] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bench_args  # noqa: E402
from nemo_cvpr2023_b200 import assets  # noqa: E402
from nemo_cvpr2023_b200._lib import lib  # noqa: E402
from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm  # noqa: E402
from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
args = bench_args(batch_size=V * F)
m = nmm.NemoV2(args, SyntheticMultiViewSequence(V, F), 'cuda', init_params=assets.make_fit_params(args, V, seed=4))
vi = torch.arange(V).repeat_interleave(F).cuda()
fi = torch.arange(F).repeat(V).cuda()
for _ in range(3):
    m.step(vi, fi)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    m.engine.forward_backward(vi, fi, update=True)
    m.engine.optimizer_step()
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / steps
lib.profile = {}
# tag every GEMM with its shape so the breakdown separates the tensor-core calls from the CUDA-core leftovers
from nemo_cvpr2023_b200 import fit_engine as _fe  # noqa: E402
_gemm = _fe.gemm


def _tagged(A, B, C, M, N, K, lda, ldb, ldc, transA=False, transB=False, **kw):
    tc = A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0 and lda % 4 == 0 and ldb % 4 == 0 and M >= 16 and N >= 8 \
        and K >= 8 and float(M) * N * K >= 2e6
    lib.tag = '%s %dx%dx%d %s%s' % ('tc' if tc else 'simt', M, N, K, 'T' if transA else 'N', 'T' if transB else 'N')
    _gemm(A, B, C, M, N, K, lda, ldb, ldc, transA=transA, transB=transB, **kw)
    lib.tag = ''


_fe.gemm = _tagged
for _ in range(steps):
    m.engine.forward_backward(vi, fi, update=True)
    m.engine.optimizer_step()
prof = lib.collect_profile()
rows = sorted(((t / steps, n // steps, k) for k, (t, n) in prof.items()), reverse=True)
print('step %.3f ms  (%d samples, %.0f frames/s); sum of entry points %.3f ms' % (
    total, V * F, V * F / total * 1e3, sum(r[0] for r in rows)))
for t, n, k in rows:
    print('%9.3f ms  %5.1f%%  x%-3d %s' % (t, 100 * t / total, n, k))
json.dump({'step_ms': total, 'entries': {k: {'ms': t, 'calls': n} for t, n, k in rows}},
          open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'profile_step.json'), 'w'), indent=1)

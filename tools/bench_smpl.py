"""BASELINE configs[1] proper: `SMPL.forward` + backward through the public autograd API, batch 4096 poses, 6890 vertices,
24 joints, betas (1,10), upstream dV ~ N(0,1), dJ ~ N(0,1), fp32, one B200.  Device time per forward + backward.
    python tools/bench_smpl.py [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemo_cvpr2023_b200.hmr.smpl import SMPL  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(5)
q = torch.randn(N * 24, 4, device=dev, generator=g)
q = q / q.norm(dim=1, keepdim=True)
w, x, y, z = q.unbind(1)
rot = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                   2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                   2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).view(N, 24, 3, 3).contiguous()
betas = torch.randn(1, 10, device=dev, generator=g)
dV = torch.randn(N, 6890, 3, device=dev, generator=g)
dJ = torch.randn(N, 49, 3, device=dev, generator=g)
smpl = SMPL('synthetic')


def once():
    b, r = betas.clone().requires_grad_(True), rot.clone().requires_grad_(True)
    out = smpl(betas=b, body_pose=r[:, 1:], global_orient=r[:, :1], pose2rot=False)
    torch.autograd.backward([out.vertices, out.joints], [dV, dJ])
    return r.grad


for _ in range(3):
    once()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    once()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print('SMPL.forward + backward, N = %d, betas (1,10): %.3f ms  (%.2f M poses/s; working set %.0f MB per call > L2)'
      % (N, t, N / t / 1e3, 5 * N * 82680 / 1e6))

import torch, time, numpy as np, sys
sys.path.insert(0, '.')
from nemo_cvpr2023_b200 import assets
from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
V, Fr = 64, 512
args = assets.default_args(batch_size=V * Fr, h_dim=1024)
m = nmm.NemoV2(args, SyntheticMultiViewSequence(V, Fr, seed=1), 'cuda', init_params=assets.make_fit_params(args, V, seed=4))
vi, fi = torch.arange(V).repeat_interleave(Fr), torch.arange(Fr).repeat(V)
for i in range(3):
    torch.cuda.synchronize(); t=time.time()
    ld, info = m.step(vi, fi)
    torch.cuda.synchronize(); print(i, round((time.time()-t)*1e3,2), 'ms', {k: float(v) for k, v in ld.items()})
print('frames/s', V*Fr/((time.time()-t)))
print(torch.cuda.max_memory_allocated()/2**30, 'GiB')
m3 = nmm.NemoV3(assets.default_args(batch_size=4096, weight_3d_loss=1000.0, instance_code_size=5), SyntheticMultiViewSequence(16, 256, seed=2), 'cuda')
ld, _ = m3.step(None, None, full_batch=True)
print('V3 custom-video full batch', {k: float(v) for k, v in ld.items()})

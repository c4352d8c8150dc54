"""The product path has no CPU / PyTorch fallback: without a CUDA device, with a CPU tensor, or
without the built library every entry fails loudly instead of computing something else.  Runs on the CPU box."""
import os

import pytest
import torch

from nemo_cvpr2023_b200 import _lib, assets


@pytest.mark.skipif(torch.cuda.is_available(), reason='this is the no-GPU behaviour')
def test_ops_and_models_refuse_to_run_without_cuda():
    import nemo_cvpr2023_b200.functional as Fn
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Fn.rot6d_to_rotmat(torch.zeros(2, 6))
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    args = assets.default_args(h_dim=16, phase_rbf_dim=4, instance_code_size=2, monotonic_network_n_nodes=4, batch_size=8)
    with pytest.raises((RuntimeError, AssertionError)):
        nmm.NemoV2(args, SyntheticMultiViewSequence(2, 4), 'cuda')


def test_binding_rejects_cpu_and_strided_tensors():
    with pytest.raises(RuntimeError, match='CUDA tensors'):
        _lib._arg(torch.zeros(4))
    assert _lib._arg(None) is None and _lib._arg(3) == 3


def test_missing_library_is_an_error_not_a_fallback(monkeypatch, tmp_path):
    fresh = _lib._Lib()
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'libnemo_b200.so'))
    with pytest.raises(RuntimeError, match='not built'):
        fresh.load()
    assert os.path.exists(os.path.join(os.path.dirname(_lib.__file__), 'libnemo_b200.so')), \
        'the in-tree library must be built for the other tests (python -c "import __graft_entry__ as g; g.build()")'

"""The custom-video adaptor (`DemoMultiViewSequence`, nemo/multi_view_sequence.py:92-247) against the output of the
reference's own loader on the same seeded directory tree (tests/golden/demo_sequence.npz, written by
oracle/make_golden_sequence.py)."""
import json
import os

import numpy as np
import pytest

from tests.util import gold, make_demo_tree


def test_demo_sequence_matches_reference_loader(tmp_path):
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import DemoMultiViewSequence
    g = gold('demo_sequence')
    seq = DemoMultiViewSequence(make_demo_tree(str(tmp_path)), 0, 9)
    assert seq.num_frames == int(g['num_frames']) and seq.num_views == int(g['num_views'])
    assert [seq.IMG_D0, seq.IMG_D1] == list(g['IMG'])
    assert np.array_equal(np.array(seq.framerate_multiplier), g['framerate_multiplier'])
    assert len(seq.gt_camera_dicts) == 1
    for v, s in enumerate(seq.sequences):
        keys = sorted(k[3:] for k in g if k.startswith('v%d_' % v))
        assert sorted(s.keys()) == keys
        for k in keys:
            assert np.array_equal(np.array(s[k]), g['v%d_%s' % (v, k)]), (v, k)
    assert np.array_equal(seq.get_image(1, 2), g['image_0_2'])
    # the frame without a detection is all zeros, and the picked VIBE track is the one near the labels
    assert any((np.array(s['pose_2d_op']) == 0).all(axis=(1, 2)).any() for s in seq.sequences)
    assert np.array(seq.sequences[0]['vibe_mask']).min() == 1.0


def test_demo_sequence_rejects_multiple_people(tmp_path):
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import DemoMultiViewSequence
    cfg = make_demo_tree(str(tmp_path))
    f = os.path.join(str(tmp_path), 'swing.1.op', '000002_keypoints.json')
    d = json.load(open(f))
    d['people'] = d['people'] * 2
    json.dump(d, open(f, 'w'))
    with pytest.raises(ValueError):
        DemoMultiViewSequence(cfg, 0, 9)


@pytest.mark.gpu
def test_fit_runs_on_a_demo_sequence(tmp_path):
    """the custom-video path end to end: adaptor -> NemoV3 (3-D loss on the VIBE poses) -> warmup + a few steps"""
    import torch

    from nemo_cvpr2023_b200 import assets
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import DemoMultiViewSequence
    seq = DemoMultiViewSequence(make_demo_tree(str(tmp_path)), 0, 9)
    args = assets.default_args(h_dim=32, phase_rbf_dim=8, instance_code_size=3, monotonic_network_n_nodes=10,
                               batch_size=12, weight_3d_loss=1000.0, lr_human=1e-3)
    m = nmm.NemoV3(args, seq, 'cuda')
    assert m.points2d_gt_all.shape == (2, 9, 25, 3) and m.hmr_theta.shape == (2, 9, 69)
    torch.manual_seed(0)
    w = m.warmup(5)
    first, _ = m.step(None, None, full_batch=True)
    for _ in range(5):
        last, _ = m.step(None, None, full_batch=True)
    assert all(np.isfinite(list(w))) and np.isfinite(last['total_loss']) and last['loss_3d'] <= first['loss_3d']

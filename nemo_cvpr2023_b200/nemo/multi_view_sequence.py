"""Data adaptor with the layout of `nemo/multi_view_sequence.py` (`sequences[v][key][f]`,
:415-443): OpenPose BODY_25 keypoints `pose_2d_op` (25,3 = x,y,conf) and VIBE `pose` (72 theta + 1 mask).
File IO of the reference (licensed NeMo-MoCap data, OpenPose JSON, VIBE pkl) is out of scope; the
synthetic sequence below feeds the fit models the same structure from seeded tensors."""
import numpy as np

from .. import assets


class SyntheticMultiViewSequence:
    def __init__(self, num_views=None, num_frames=None, seed=1, data=None):
        d = data if data is not None else assets.make_fit_data(num_views, num_frames, seed)
        num_views, num_frames = d['points2d_gt_all'].shape[:2]
        self.num_views, self.num_frames = num_views, num_frames
        self.IMG_D0, self.IMG_D1 = d['IMG_D0'], d['IMG_D1']
        self.sequences = []
        r = np.random.RandomState(1000 + seed)
        for v in range(num_views):
            pose = np.concatenate([np.zeros((num_frames, 3), np.float32), d['hmr_theta'][v], d['hmr_mask'][v]], 1)
            # evaluation labels (the NeMo-MoCap `pose_3d_gt` / `trans_3d_gt` tracks and the VIBE 2-D joints): seeded
            # perturbations of the fit targets, enough to drive eval_3d / eval_2d
            gt3d = np.concatenate([np.zeros((num_frames, 3)), d['hmr_theta'][v] + 0.05 * r.standard_normal((num_frames, 69))],
                                  1).astype(np.float32)
            vibe2d = d['points2d_gt_all'][v].copy()
            vibe2d[..., :2] += (10.0 * r.standard_normal(vibe2d[..., :2].shape)).astype(np.float32)
            self.sequences.append({'pose_2d_op': d['points2d_gt_all'][v], 'pose_2d_gt': d['points2d_gt_all'][v],
                                   'pose': pose, 'pose_3d_gt': gt3d,
                                   'trans_3d_gt': np.zeros((num_frames, 3), np.float32), 'vibe_joints2d': vibe2d})
        self.gt_bbox_size = d['gt_bbox_size']
        self.framerate_multiplier = [1.0] * num_views


MultiViewSequence = SyntheticMultiViewSequence

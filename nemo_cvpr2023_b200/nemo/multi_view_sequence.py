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
        for v in range(num_views):
            pose = np.concatenate([np.zeros((num_frames, 3), np.float32), d['hmr_theta'][v], d['hmr_mask'][v]], 1)
            self.sequences.append({'pose_2d_op': d['points2d_gt_all'][v], 'pose_2d_gt': d['points2d_gt_all'][v],
                                   'pose': pose})
        self.gt_bbox_size = d['gt_bbox_size']


MultiViewSequence = SyntheticMultiViewSequence

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


# ------------------------------------------------------------------------------------------------------------------
# Custom-video adaptor: `DemoMultiViewSequence` of nemo/multi_view_sequence.py:92-247, the
# loader behind custom_video/nemo-run.sh.  Same directory convention (`<exp_dir>/<name>.frames/*.png`,
# `<name>.op/%06d_keypoints.json` = OpenPose BODY_25, `<name>.vibe/vibe_output.pkl`, optional `<name>.cam.pickle`),
# same frame sampling (`tidx = floor(phase * n_seq_frames)`), same person selection, same output layout
# (`sequences[v][key][f]`).  Host-side file IO only; nothing here touches the GPU.
_PASS_THROUGH = ('betas', 'frame_ids')   # per-track entries the reference keeps at track length (:33-35)


def _densify(track, n_frames):
    """One VIBE track -> arrays indexed by video frame.  Every per-frame array of the track is scattered to its
    `frame_ids` rows of a zero float32 array of `n_frames` rows, and `mask` marks the frames the person was tracked on
    (semantics of prepare_person_dict, nemo/multi_view_sequence.py:30-48; `None` entries are dropped)."""
    ids = np.asarray(track['frame_ids'], dtype=np.int64)
    present = np.zeros(n_frames, dtype=np.float32)
    present[ids] = 1.0
    dense = {k: track[k] for k in _PASS_THROUGH if k in track}
    for name, values in track.items():
        if name in _PASS_THROUGH or values is None:
            continue
        values = np.asarray(values)
        out = np.zeros((n_frames,) + values.shape[1:], dtype=np.float32)
        out[ids] = values
        dense[name] = out
    dense['mask'] = present
    return dense


def prepare_person_dict(person_output, max_frames):
    return _densify(person_output, max_frames)


def prepare_vibe_dict(vibe_output, max_frames):
    """:51-55: every track of a `vibe_output.pkl` densified."""
    return {pid: _densify(track, max_frames) for pid, track in vibe_output.items()}


def select_person_at_center(vibe_output, img_size, max_frames, all_gt_2d):
    """:58-89: the track that follows the 2-D labels most closely.  Score of a track = distance between the mean of
    its first 15 projected joints and the mean of the labels, averaged over the frames the track is present on; the
    lowest score wins, ties go to the first track (`img_size` / `max_frames` are unused there as well)."""
    if not vibe_output:
        return None
    keys = list(vibe_output)
    target = np.asarray(all_gt_2d).mean(1)                                     # (T, 2)
    scores = np.empty(len(keys))
    for i, k in enumerate(keys):
        tr = vibe_output[k]
        j2d = np.asarray(tr['joints2d_img_coord'] if 'joints2d_img_coord' in tr else tr['smpl_joints2d'])
        off = np.linalg.norm(j2d[:, :15].mean(1) - target, axis=-1)            # (T,)
        scores[i] = (off * tr['mask']).sum() / tr['mask'].sum()
    # the reference starts from `best = inf` with a strict `<`: a NaN score never wins
    scores = np.where(np.isnan(scores), np.inf, scores)
    best = int(np.argmin(scores))
    return vibe_output[keys[best]] if np.isfinite(scores[best]) else vibe_output.get(0, vibe_output[keys[0]])


def _read_openpose(op_file, scale2=False):
    """one BODY_25 person -> (25,3) [x, y, conf]; nobody -> zeros; several people -> ValueError (:147-158, :185-197)"""
    import json
    people = json.load(open(op_file, 'r'))['people']
    if len(people) == 1:
        pose = np.array(people[0]['pose_keypoints_2d']).reshape(25, 3)
        if scale2:
            pose[:, :2] *= 2
        return pose
    if len(people) == 0:
        return np.zeros((25, 3))
    raise ValueError('multiple people in %s' % op_file)


class DemoMultiViewSequence:
    def __init__(self, nemo_data_cfg, start_phase, num_frames):
        import os
        import os.path as osp
        from collections import defaultdict

        import joblib
        self.nemo_data_cfg, self.start_phase = nemo_data_cfg, start_phase
        self.sequences = []
        exp_dir, names = nemo_data_cfg['exp_dir'], nemo_data_cfg['videos']['names']

        def n_png(name):
            return len([s for s in os.listdir(osp.join(exp_dir, name + '.frames')) if s.endswith('.png')])

        min_num_frames = min(n_png(name) for name in names)
        start_min_frame = np.round(min_num_frames * start_phase)
        self.num_frames = int(min([num_frames, min_num_frames - start_min_frame - 1]))
        max_sizes = [0, 0]
        self.video_img_dir_list, self.n_seq_frames_list, self.gt_camera_dicts, self.framerate_multiplier = [], [], [], []
        for view_idx, name in enumerate(names):
            cur_seq = defaultdict(list)
            img_dir, op_out_dir = osp.join(exp_dir, name + '.frames'), osp.join(exp_dir, name + '.op')
            cam_path = osp.join(exp_dir, name + '.cam.pickle')
            if osp.exists(cam_path):
                self.gt_camera_dicts.append(joblib.load(cam_path))
            n_seq_frames = n_png(name)
            assert start_phase == 0
            self.framerate_multiplier.append(n_seq_frames / self.num_frames)
            self.video_img_dir_list.append(img_dir)
            self.n_seq_frames_list.append(n_seq_frames)
            dummy_img = self.get_raw_image(view_idx, 0)
            # the reference doubles the coordinates here unconditionally (:151) -- only used to pick the VIBE track
            all_op_2d = np.array([_read_openpose(osp.join(op_out_dir, f"{t + 1:06d}_keypoints.json"), True)[:15, :2]
                                  for t in range(n_seq_frames)])
            hmr_output = prepare_vibe_dict(joblib.load(osp.join(exp_dir, name + '.vibe', 'vibe_output.pkl')), n_seq_frames)
            person = select_person_at_center(hmr_output, dummy_img.shape[:2], n_seq_frames, all_op_2d)
            hmr_pose = np.concatenate([person['pose'], np.ones((n_seq_frames, 1))], 1)
            for frame_idx in range(self.num_frames):
                phase = start_phase + (1 - start_phase) * float(frame_idx / self.num_frames)
                tidx = int(np.floor(phase * n_seq_frames))
                cur_seq['pose_2d_op'].append(_read_openpose(osp.join(op_out_dir, f"{tidx + 1:06d}_keypoints.json"),
                                                            'opencap' in exp_dir))
                cur_seq['pose'].append(hmr_pose[tidx])
                cur_seq['vibe_cam'].append(person['orig_cam'][tidx])
                cur_seq['vibe_verts'].append(person['verts'][tidx])
                cur_seq['vibe_mask'].append(person['mask'][tidx])
                cur_seq['vibe_joints2d'].append(person['joints2d_img_coord'][tidx])
            self.sequences.append(cur_seq)
            for i in range(2):
                max_sizes[i] = max(max_sizes[i], dummy_img.shape[i])
        self.IMG_D0, self.IMG_D1 = max_sizes
        self.num_views = len(self.sequences)

    def get_raw_image(self, view_idx, frame_idx):
        """:224-240"""
        import os

        import cv2
        phase = self.start_phase + (1 - self.start_phase) * float(frame_idx / self.num_frames)
        tidx = int(np.floor(phase * self.n_seq_frames_list[view_idx]))
        path = os.path.join(self.video_img_dir_list[view_idx], f"{tidx + 1:06d}.png")
        if not os.path.exists(path):
            raise ValueError(path, ' doesnt exist.')
        return cv2.imread(path)[..., ::-1]

    def get_image(self, view_idx, frame_idx):
        """:242-247"""
        raw = self.get_raw_image(view_idx, frame_idx)
        im = np.zeros((self.IMG_D0, self.IMG_D1, 3))
        im[:raw.shape[0], :raw.shape[1]] = raw
        return im.astype('uint8')

"""Constants of the reference's `hmr/hmr_constants.py` that the fit path reads."""
from ..assets import FOCAL_LENGTH, JOINT_MAP_49  # noqa: F401

IMG_RES = 224
JOINT_NAMES = [
    'OP Nose', 'OP Neck', 'OP RShoulder', 'OP RElbow', 'OP RWrist', 'OP LShoulder', 'OP LElbow', 'OP LWrist',
    'OP MidHip', 'OP RHip', 'OP RKnee', 'OP RAnkle', 'OP LHip', 'OP LKnee', 'OP LAnkle', 'OP REye', 'OP LEye',
    'OP REar', 'OP LEar', 'OP LBigToe', 'OP LSmallToe', 'OP LHeel', 'OP RBigToe', 'OP RSmallToe', 'OP RHeel',
    'Right Ankle', 'Right Knee', 'Right Hip', 'Left Hip', 'Left Knee', 'Left Ankle', 'Right Wrist', 'Right Elbow',
    'Right Shoulder', 'Left Shoulder', 'Left Elbow', 'Left Wrist', 'Neck (LSP)', 'Top of Head (LSP)',
    'Pelvis (MPII)', 'Thorax (MPII)', 'Spine (H36M)', 'Jaw (H36M)', 'Head (H36M)', 'Nose', 'Left Eye', 'Right Eye',
    'Left Ear', 'Right Ear']
JOINT_IDS = {n: i for i, n in enumerate(JOINT_NAMES)}
JOINT_MAP = {n: JOINT_MAP_49[i] for i, n in enumerate(JOINT_NAMES)}

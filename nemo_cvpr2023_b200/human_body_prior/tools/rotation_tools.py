"""Mirror of the two rotation helpers of `human_body_prior/tools/rotation_tools.py` the fit path uses."""
from ... import functional as Fn


def matrot2aa(pose_matrot):
    """rotation_tools.py:73-81 -> tgm_conversion.py:219-239 (no NaN patch).  (N,3,3) -> (N,3)."""
    return Fn.rotmat_to_aa(pose_matrot, nan_to_zero=False)


def aa2matrot(pose):
    """rotation_tools.py:83-92 (tgm.angle_axis_to_rotation_matrix == skew Rodrigues up to its small-angle branch)."""
    return Fn.batch_rodrigues(pose.reshape(-1, 3), kind=1)

"""Pose errors used as the stage-2 loss and by choose-closest (contract of
pose_optimization/two_view/compute_pose_error.py:3-21): geodesic rotation angle and the angle between translation
directions, on [B,4,4] transforms, differentiable.  The batched version used inside the w8pt kernel is in
csrc/pose_w8pt.cu."""
import torch


def _angle(cos):
    return torch.acos(cos.clamp(-1., 1.)).abs()


def compute_rotation_error(T0, T1, reduce=True):
    """acos((trace(R0^T R1) - 1) / 2); mean over the batch unless reduce=False."""
    trace = (T0[..., :3, :3] * T1[..., :3, :3]).sum((-2, -1))      # trace(R0^T R1) = <R0, R1>_F
    err = _angle((trace - 1.) / 2.)
    return err.mean() if reduce else err


def compute_translation_error_as_angle(T0, T1, reduce=True):
    """Angle between the translation vectors; items whose |t0||t1| <= 1e-6 are left out (as the reference does)."""
    t0, t1 = T0[..., :3, 3], T1[..., :3, 3]
    scale = t0.norm(dim=-1) * t1.norm(dim=-1)
    keep = scale > 1e-6
    err = _angle((t0[keep] * t1[keep]).sum(-1) / scale[keep])
    return err.mean() if reduce else err

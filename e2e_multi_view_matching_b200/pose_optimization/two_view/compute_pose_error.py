"""Pose errors used as the stage-2 loss and by choose-closest
(pose_optimization/two_view/compute_pose_error.py:3-21).  A handful of elementwise torch ops on
[B,4,4] tensors (kept differentiable); the batched version used inside the w8pt kernel is in
csrc/pose_w8pt.cu."""
import torch


def compute_rotation_error(T0, T1, reduce=True):
    R = T0[..., :3, :3].transpose(-1, -2) @ T1[..., :3, :3]
    cos_a = (R.diagonal(offset=0, dim1=-1, dim2=-2).sum(-1) - 1.) / 2.
    err = torch.abs(torch.arccos(torch.clamp(cos_a, -1., 1.)))
    return err.mean() if reduce else err


def compute_translation_error_as_angle(T0, T1, reduce=True):
    n = torch.linalg.norm(T0[..., :3, 3], dim=-1) * torch.linalg.norm(T1[..., :3, 3], dim=-1)
    valid_n = n > 1e-6
    dot = (T0[..., :3, 3][valid_n] * T1[..., :3, 3][valid_n]).sum(-1)
    err = torch.abs(torch.arccos((dot / n[valid_n]).clamp(-1., 1.)))
    return err.mean() if reduce else err

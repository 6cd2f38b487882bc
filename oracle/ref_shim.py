"""Run the reference's OWN two-view pose files, unmodified, on CPU.  TEST INFRASTRUCTURE ONLY.

``pose_optimization/two_view/estimate_relative_pose.py`` and ``bundle_adjust_gauss_newton_2_view.py``
are device-agnostic torch code; they fail to import here only because three pip-pinned third-party
packages are absent from this image (kornia==0.7.0, pytorch3d==0.7.5, coloredlogs; requirements.txt:20,36).
Of those packages the two files use nine LEAF functions and nothing else:

    kornia.geometry.epipolar.{normalize_points, normalize_transformation, motion_from_essential,
        motion_from_essential_choose_solution, triangulate_points, symmetrical_epipolar_distance}
        (estimate_relative_pose.py:2-3, bundle_adjust_gauss_newton_2_view.py:122)
    kornia.geometry.epipolar.projection.depth_from_point          (estimate_relative_pose.py:4)
    pytorch3d.transforms.so3.hat, pytorch3d.transforms.se3_exp_map (bundle_adjust_gauss_newton_2_view.py:63,194)
    coloredlogs.install                                            (bundle_adjust_gauss_newton_2_view.py:2-3)

This module registers stub modules that provide exactly those leaves (torch restatements of the published
kornia 0.7.0 / pytorch3d 0.7.5 functions, statement for statement, including kornia's batch-0 indexing in
``motion_from_essential_choose_solution``) and then imports the two reference files by their real package
path from /root/reference.  Everything ABOVE the leaves -- the weighting, the design matrix, the rank-2
projection, the candidate selection, the inlier test, the observation bookkeeping, the Jacobians, the LM
schedule, the dense LU step, the best-iterate bookkeeping -- is therefore the reference's own code, executed
as written.  ``oracle/make_pose_golden.py`` uses it to pin ``oracle/pose.py`` and to write
``tests/golden/pose_*.npz``.

Needs /root/reference, so it only runs in the authoring container (never on the GPU box).
"""
import importlib
import sys
import types

import torch

REF = '/root/reference'


# ------------------------------------------------------------------------------------------------
# kornia 0.7.0 leaves (kornia/geometry/{conversions,linalg,epipolar/*}.py)
# ------------------------------------------------------------------------------------------------
def convert_points_from_homogeneous(points, eps=1e-8):
    z_vec = points[..., -1:]
    mask = torch.abs(z_vec) > eps
    scale = torch.where(mask, 1.0 / (z_vec + eps), torch.ones_like(z_vec))
    return scale * points[..., :-1]


def convert_points_to_homogeneous(points):
    return torch.nn.functional.pad(points, [0, 1], 'constant', 1.0)


def transform_points(trans_01, points_1):
    shape_inp = list(points_1.shape)
    points_1 = points_1.reshape(-1, points_1.shape[-2], points_1.shape[-1])
    trans_01 = trans_01.reshape(-1, trans_01.shape[-2], trans_01.shape[-1])
    trans_01 = torch.repeat_interleave(trans_01, repeats=points_1.shape[0] // trans_01.shape[0], dim=0)
    points_1_h = convert_points_to_homogeneous(points_1)
    points_0_h = torch.bmm(points_1_h, trans_01.permute(0, 2, 1))
    points_0_h = torch.squeeze(points_0_h, dim=-1)
    points_0 = convert_points_from_homogeneous(points_0_h)
    shape_inp[-2] = points_0.shape[-2]
    shape_inp[-1] = points_0.shape[-1]
    return points_0.reshape(shape_inp)


def normalize_points(points, eps=1e-8):
    x_mean = torch.mean(points, dim=1, keepdim=True)
    scale = (points - x_mean).norm(dim=-1, p=2).mean(dim=-1)
    scale = torch.sqrt(torch.tensor(2.0)) / (scale + eps)
    ones, zeros = torch.ones_like(scale), torch.zeros_like(scale)
    transform = torch.stack([scale, zeros, -scale * x_mean[..., 0, 0], zeros, scale, -scale * x_mean[..., 0, 1],
                             zeros, zeros, ones], dim=-1)
    transform = transform.view(-1, 3, 3)
    points_norm = transform_points(transform, points)
    return points_norm, transform


def normalize_transformation(M, eps=1e-8):
    norm_val = M[..., -1:, -1:]
    return torch.where(norm_val.abs() > eps, M / (norm_val + eps), M)


def _torch_svd_cast(x):
    # kornia.utils.helpers._torch_svd_cast: torch.svd in >= fp32, cast back
    dtype = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32
    out1, out2, out3 = torch.svd(x.to(dtype))
    return out1.to(x.dtype), out2.to(x.dtype), out3.to(x.dtype)


def cross_product_matrix(x):
    x0, x1, x2 = x[..., 0], x[..., 1], x[..., 2]
    zeros = torch.zeros_like(x0)
    cross = torch.stack([zeros, -x2, x1, x2, zeros, -x0, -x1, x0, zeros], dim=-1)
    return cross.view(*x.shape[:-1], 3, 3)


def decompose_essential_matrix(E_mat):
    U, _, V = _torch_svd_cast(E_mat)
    Vt = V.transpose(-2, -1)
    mask = torch.ones_like(E_mat)
    mask[..., -1:] *= -1.0
    maskt = mask.transpose(-2, -1)
    U = torch.where((torch.det(U) < 0.0)[..., None, None], U * mask, U)
    Vt = torch.where((torch.det(Vt) < 0.0)[..., None, None], Vt * maskt, Vt)
    W = cross_product_matrix(torch.tensor([[0.0, 0.0, 1.0]]).type_as(E_mat))
    W[..., 2, 2] += 1.0
    U_W_Vt = U @ W @ Vt
    U_Wt_Vt = U @ W.transpose(-2, -1) @ Vt
    return U_W_Vt, U_Wt_Vt, U[..., -1:]


def motion_from_essential(E_mat):
    R1, R2, t = decompose_essential_matrix(E_mat)
    Rs = torch.stack([R1, R1, R2, R2], dim=-3)
    Ts = torch.stack([t, -t, t, -t], dim=-3)
    return Rs, Ts


def projection_from_KRt(K, R, t):
    Rt = torch.cat([R, t], dim=-1)
    Rt_h = torch.nn.functional.pad(Rt, [0, 0, 0, 1], 'constant', 0.0)
    Rt_h[..., -1, -1] += 1.0
    K_h = torch.nn.functional.pad(K, [0, 1, 0, 1], 'constant', 0.0)
    K_h[..., -1, -1] += 1.0
    return K @ Rt


def triangulate_points(P1, P2, points1, points2):
    points_shape = max(points1.shape, points2.shape)
    X = torch.zeros(points_shape[:-1] + (4, 4)).type_as(points1)
    for i in range(4):
        X[..., 0, i] = points1[..., 0] * P1[..., 2:3, i] - P1[..., 0:1, i]
        X[..., 1, i] = points1[..., 1] * P1[..., 2:3, i] - P1[..., 1:2, i]
        X[..., 2, i] = points2[..., 0] * P2[..., 2:3, i] - P2[..., 0:1, i]
        X[..., 3, i] = points2[..., 1] * P2[..., 2:3, i] - P2[..., 1:2, i]
    _, _, V = _torch_svd_cast(X)
    points3d_h = V[..., -1]
    return convert_points_from_homogeneous(points3d_h)


def depth_from_point(R, t, X):
    X_tmp = R @ X.transpose(-2, -1)
    return X_tmp[..., 2, :] + t[..., 2, :]


def motion_from_essential_choose_solution(E_mat, K1, K2, x1, x2, mask=None):
    unbatched = len(E_mat.shape) == 2
    if unbatched:
        E_mat, K1, K2, x1, x2 = E_mat[None], K1[None], K2[None], x1[None], x2[None]
        if mask is not None:
            mask = mask[None]
    Rs, ts = motion_from_essential(E_mat)
    R1 = torch.eye(3, device=E_mat.device, dtype=E_mat.dtype)[None].repeat(E_mat.shape[0], 1, 1)
    t1 = torch.zeros(E_mat.shape[0], 3, 1, device=E_mat.device, dtype=E_mat.dtype)
    R1 = R1[:, None].expand(-1, 4, -1, -1)
    t1 = t1[:, None].expand(-1, 4, -1, -1)
    K1 = K1[:, None].expand(-1, 4, -1, -1)
    P1 = projection_from_KRt(K1, R1, t1)
    R2 = Rs
    t2 = ts
    K2 = K2[:, None].expand(-1, 4, -1, -1)
    P2 = projection_from_KRt(K2, R2, t2)
    x1 = x1[:, None].expand(-1, 4, -1, -1)
    x2 = x2[:, None].expand(-1, 4, -1, -1)
    X = triangulate_points(P1, P2, x1, x2)
    d1 = depth_from_point(R1, t1, X)
    d2 = depth_from_point(R2, t2, X)
    depth_mask = (d1 > 0.0) & (d2 > 0.0)
    if mask is not None:
        depth_mask &= mask.unsqueeze(1)
    mask_indices = torch.max(depth_mask.sum(-1), dim=-1, keepdim=True)[1]
    # kornia 0.7.0 indexes with batch element 0's choice; kept as published
    R_out = Rs[:, mask_indices][:, 0, 0]
    t_out = ts[:, mask_indices][:, 0, 0]
    points3d_out = X[:, mask_indices][:, 0, 0]
    if unbatched:
        return R_out[0], t_out[0], points3d_out[0]
    return R_out, t_out, points3d_out


def symmetrical_epipolar_distance(pts1, pts2, Fm, squared=True, eps=1e-8):
    if pts1.shape[-1] == 2:
        pts1 = convert_points_to_homogeneous(pts1)
    if pts2.shape[-1] == 2:
        pts2 = convert_points_to_homogeneous(pts2)
    F_t = Fm.transpose(dim0=len(Fm.shape) - 2, dim1=len(Fm.shape) - 1)
    line1_in_2 = pts1 @ F_t
    line2_in_1 = pts2 @ Fm
    numerator = (pts2 * line1_in_2).sum(dim=-1).pow(2)
    denominator_inv = 1.0 / (line1_in_2[..., :2].norm(2, dim=-1).pow(2)) + 1.0 / (line2_in_1[..., :2].norm(2, dim=-1).pow(2))
    out = numerator * denominator_inv
    if squared:
        return out
    return (out + eps).sqrt()


# ------------------------------------------------------------------------------------------------
# pytorch3d 0.7.5 leaves (pytorch3d/transforms/{so3,se3}.py)
# ------------------------------------------------------------------------------------------------
def hat(v):
    N, dim = v.shape
    if dim != 3:
        raise ValueError('Input vectors have to be 3-dimensional.')
    h = torch.zeros((N, 3, 3), dtype=v.dtype, device=v.device)
    x, y, z = v.unbind(1)
    h[:, 0, 1] = -z
    h[:, 0, 2] = y
    h[:, 1, 0] = z
    h[:, 1, 2] = -x
    h[:, 2, 0] = -y
    h[:, 2, 1] = x
    return h


def _so3_exp_map(log_rot, eps=0.0001):
    nrms = (log_rot * log_rot).sum(1)
    rot_angles = torch.clamp(nrms, eps).sqrt()
    rot_angles_inv = 1.0 / rot_angles
    fac1 = rot_angles_inv * rot_angles.sin()
    fac2 = rot_angles_inv * rot_angles_inv * (1.0 - rot_angles.cos())
    skews = hat(log_rot)
    skews_square = torch.bmm(skews, skews)
    R = fac1[:, None, None] * skews + fac2[:, None, None] * skews_square + \
        torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None]
    return R, rot_angles, skews, skews_square


def _se3_V_matrix(log_rotation, log_rotation_hat, log_rotation_hat_square, rotation_angles, eps=1e-4):
    V = (torch.eye(3, dtype=log_rotation.dtype, device=log_rotation.device)[None]
         + log_rotation_hat * ((1 - torch.cos(rotation_angles)) / (rotation_angles ** 2))[:, None, None]
         + log_rotation_hat_square * ((rotation_angles - torch.sin(rotation_angles)) / (rotation_angles ** 3))[:, None, None])
    return V


def se3_exp_map(log_transform, eps=1e-4):
    if log_transform.ndim != 2 or log_transform.shape[1] != 6:
        raise ValueError('Expected input to be of shape (N, 6).')
    N, _ = log_transform.shape
    log_translation = log_transform[..., :3]
    log_rotation = log_transform[..., 3:]
    R, rotation_angles, log_rotation_hat, log_rotation_hat_square = _so3_exp_map(log_rotation, eps=eps)
    V = _se3_V_matrix(log_rotation, log_rotation_hat, log_rotation_hat_square, rotation_angles, eps=eps)
    T = torch.bmm(V, log_translation[:, :, None])[:, :, 0]
    transform = torch.zeros(N, 4, 4, dtype=log_transform.dtype, device=log_transform.device)
    transform[:, :3, :3] = R
    transform[:, :3, 3] = T
    transform[:, 3, 3] = 1.0
    return transform.permute(0, 2, 1)


# ------------------------------------------------------------------------------------------------
# stub modules + import of the unmodified reference files
# ------------------------------------------------------------------------------------------------
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__oracle_stub__ = True
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """-> (estimate_relative_pose module, bundle_adjust_gauss_newton_2_view module) of the REFERENCE."""
    global _loaded
    if _loaded is not None:
        return _loaded
    for real in ('kornia', 'pytorch3d', 'coloredlogs'):
        if real in sys.modules and not getattr(sys.modules[real], '__oracle_stub__', False):
            raise RuntimeError('%s is really installed: import the reference directly instead of the shim' % real)
    proj = _module('kornia.geometry.epipolar.projection', depth_from_point=depth_from_point,
                   projection_from_KRt=projection_from_KRt)
    epi = _module('kornia.geometry.epipolar', normalize_points=normalize_points,
                  normalize_transformation=normalize_transformation, motion_from_essential=motion_from_essential,
                  motion_from_essential_choose_solution=motion_from_essential_choose_solution,
                  triangulate_points=triangulate_points, symmetrical_epipolar_distance=symmetrical_epipolar_distance,
                  projection=proj)
    geo = _module('kornia.geometry', epipolar=epi)
    _module('kornia', geometry=geo)
    so3 = _module('pytorch3d.transforms.so3', hat=hat)
    tr = _module('pytorch3d.transforms', so3=so3, se3_exp_map=se3_exp_map)
    _module('pytorch3d', transforms=tr)
    _module('coloredlogs', install=lambda *a, **k: None)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    erp = importlib.import_module('pose_optimization.two_view.estimate_relative_pose')
    ba = importlib.import_module('pose_optimization.two_view.bundle_adjust_gauss_newton_2_view')
    assert erp.__file__.startswith(REF) and ba.__file__.startswith(REF), (erp.__file__, ba.__file__)
    _loaded = (erp, ba)
    return _loaded

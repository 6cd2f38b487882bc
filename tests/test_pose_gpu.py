"""GPU parity of the two-view pose kernels against the oracle restatement (oracle/pose.py).
The reference computes in fp32 (cuSOLVER SVD / dense LU); the kernels compute in fp64.  They are
compared with the fp64 run of the oracle (tight) and with its fp32 run (at the fp32 path's own
noise level, measured in DESIGN.md)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene_batch(seeds, n, outlier_frac):
    from oracle import pose as P
    sc = [P.make_two_view_scene(s, n, outlier_frac=outlier_frac) for s in seeds]
    return {k: np.concatenate([s[k] for s in sc], 0) for k in sc[0]}


def _ours_w8pt(s, **kw):
    from e2e_multi_view_matching_b200.pose_optimization.two_view.estimate_relative_pose import estimate_relative_pose_w8pt
    t = {k: torch.from_numpy(v).cuda() for k, v in s.items() if k != 'outlier'}
    return estimate_relative_pose_w8pt(t['kpts0'], t['kpts1'], t['intr'], t['intr'], t['conf'], **kw)


@pytest.mark.parametrize('outlier_frac', [0.0, 0.3])
def test_w8pt_vs_oracle(outlier_frac):
    from oracle import pose as P
    s = _scene_batch(range(6), 200, outlier_frac)
    T, info = _ours_w8pt(s, determine_inliers=True)
    T = T.cpu().numpy()
    for b in range(6):
        one = {k: v[b:b + 1] for k, v in s.items()}
        for dt, tol in ((np.float64, 2e-6), (np.float32, 1e-4)):
            Tr, ir = P.estimate_relative_pose_w8pt(one['kpts0'].astype(dt), one['kpts1'].astype(dt), one['intr'].astype(dt),
                                                   one['intr'].astype(dt), one['conf'].astype(dt), determine_inliers=True)
            np.testing.assert_allclose(T[b], Tr[0], atol=tol, rtol=tol)
        np.testing.assert_allclose(info['kpts0_norm'][b].cpu().numpy(), ir['kpts0_norm'][0], atol=1e-6)
        np.testing.assert_allclose(info['confidence'][b].cpu().numpy(), ir['confidence'][0], rtol=1e-5)
        assert (info['pos_depth_mask'][b].cpu().numpy() == ir['pos_depth_mask'][0]).mean() > 0.99
        assert (info['inliers'][b].cpu().numpy() == ir['inliers'][0]).mean() > 0.99
        Fo = info['F'][b].cpu().numpy()
        np.testing.assert_allclose(Fo, ir['F'][0], atol=5e-4 * np.abs(ir['F'][0]).max())


def test_w8pt_choose_closest_and_short_input():
    from oracle import pose as P
    s = _scene_batch(range(4), 100, 0.1)
    T, info = _ours_w8pt(s, choose_closest=True, T_021=torch.from_numpy(s['T_021']).cuda())
    Tr, _ = P.estimate_relative_pose_w8pt(s['kpts0'].astype(np.float64), s['kpts1'].astype(np.float64), s['intr'].astype(np.float64),
                                          s['intr'].astype(np.float64), s['conf'].astype(np.float64), choose_closest=True,
                                          T_021=s['T_021'].astype(np.float64))
    np.testing.assert_allclose(T.cpu().numpy(), Tr, atol=2e-6)
    short = {k: (v[:, :5] if v.ndim == 3 and v.shape[1] == 100 else v) for k, v in s.items()}
    assert _ours_w8pt(short) == (None, None)


@pytest.mark.parametrize('outlier_frac', [0.0, 0.3])
def test_ba2view_vs_oracle(outlier_frac):
    from oracle import pose as P
    from e2e_multi_view_matching_b200.pose_optimization.two_view.estimate_relative_pose import run_bundle_adjust_2_view
    s = _scene_batch(range(5), 120, outlier_frac)
    T, info = _ours_w8pt(s, determine_inliers=True)
    conf = info['confidence'].clone()
    conf[torch.logical_not(info['pos_depth_mask'])] = 0.
    conf[4] = 0.                       # item with no valid match -> excluded
    ext, valid = run_bundle_adjust_2_view(info['kpts0_norm'], info['kpts1_norm'], conf, T, n_iterations=10)
    assert valid.cpu().tolist() == [True, True, True, True, False]
    ext = ext.cpu().numpy()
    k0, k1 = info['kpts0_norm'].cpu().numpy(), info['kpts1_norm'].cpu().numpy()
    cn, Tn = conf.cpu().numpy(), T.cpu().numpy()
    e64, v64, tr64 = P.run_bundle_adjust_2_view(k0.astype(np.float64), k1.astype(np.float64), cn.astype(np.float64),
                                                Tn.astype(np.float64), 10, return_trace=True)
    assert v64.tolist() == valid.cpu().tolist()
    np.testing.assert_allclose(ext, e64, atol=5e-6, rtol=5e-6)
    e32, _ = P.run_bundle_adjust_2_view(k0, k1, cn, Tn, 10)
    np.testing.assert_allclose(ext, e32, atol=1e-3)      # fp32 dense-LU noise of the reference path


def test_ba2view_improves_pose():
    """End-to-end property at full size (n = 1024): BA lowers the residual and the pose error."""
    from oracle import pose as P
    from e2e_multi_view_matching_b200.pose_optimization.two_view.bundle_adjust_gauss_newton_2_view import BundleAdjustGaussNewton2View
    s = _scene_batch([11], 1024, 0.0)
    T, info = _ours_w8pt(s, determine_inliers=True)
    ba = BundleAdjustGaussNewton2View(1, 10)
    ext, valid = ba.run(info['kpts0_norm'], info['kpts1_norm'], info['confidence'].squeeze(-1), T, return_trace=True)
    tr = ba.last_trace.cpu().numpy()[0]
    assert tr.min() < tr[0]
    g = s['T_021'][0].astype(np.float64)
    e0 = P.compute_pose_error(g, T[0, :3, :3].cpu().numpy().astype(np.float64), T[0, :3, 3].cpu().numpy().astype(np.float64))
    e1 = P.compute_pose_error(g, ext[0, 1, :3, :3].cpu().numpy().astype(np.float64), ext[0, 1, :3, 3].cpu().numpy().astype(np.float64))
    assert max(e1) <= max(e0) + 0.05

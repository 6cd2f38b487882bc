"""CPU: cross-checks of the optimisation parts of the pose oracle against SciPy (independent solvers that are in
the image): the SE(3) exponential of the two-view BA step vs scipy.linalg.expm, and the minimum the Ceres-style
LM of oracle/mvba.py reaches on the reference's gtest scene vs scipy.optimize.least_squares on the same residuals."""
import numpy as np
import pytest

scipy_linalg = pytest.importorskip('scipy.linalg')
from scipy.optimize import least_squares

from oracle import pose as P, mvba as M


def test_se3_exp_map_matches_expm():
    rng = np.random.default_rng(0)
    for _ in range(30):
        v = rng.standard_normal(3)
        w = rng.standard_normal(3) * rng.uniform(0.05, 2.5)          # above pytorch3d's eps clamp
        X = np.zeros((4, 4))
        X[:3, :3] = P.hat(w)
        X[:3, 3] = v
        ref = scipy_linalg.expm(X)
        got = P.se3_exp_map_T(np.concatenate([v, w])[None].astype(np.float64))[0]
        np.testing.assert_allclose(got, ref, atol=1e-10)


def test_two_view_ba_cost_decreases_and_jacobian_is_exact():
    sc = P.make_two_view_scene(3, 40, outlier_frac=0.0, noise_px=0.5, dtype=np.float64)
    conf = np.ones_like(sc['conf'])
    T0, info = P.estimate_relative_pose_w8pt(sc['kpts0'], sc['kpts1'], sc['intr'], sc['intr'], conf)
    ext, valid, trace = P.run_bundle_adjust_2_view(info['kpts0_norm'], info['kpts1_norm'], info['confidence'], T0,
                                                   n_iterations=10, return_trace=True)
    assert valid.all()
    tr = np.asarray(trace).reshape(-1)
    assert (np.diff(tr) <= 1e-12 * tr[0]).all() and tr[-1] < 0.95 * tr[0]      # the reference's damped LM: slow, monotone
    # analytic Jacobian of the step parameterisation T1 <- exp([v|w]) T1, p <- p + dp  vs central differences
    rng = np.random.default_rng(0)
    n = 6
    x0, x1 = info['kpts0_norm'][0, :n], info['kpts1_norm'][0, :n]
    w = np.full(n, 1.0 / n)
    T1 = T0[0].astype(np.float64)
    pts = P.triangulate_points(np.eye(4)[None, :3], T1[None, :3], x0[None], x1[None])[0]
    J, r = P._ba_residual_jacobian(T1, pts, x0, x1, w)
    eps = 1e-6
    for k in range(6 + 3 * n):
        d = np.zeros(6 + 3 * n)
        d[k] = eps
        rp = P._ba_residual_jacobian(P.se3_exp_map_T(d[None, :6], eps=1e-30)[0] @ T1, pts + d[6:].reshape(n, 3), x0, x1, w)[1]
        rm = P._ba_residual_jacobian(P.se3_exp_map_T(-d[None, :6], eps=1e-30)[0] @ T1, pts - d[6:].reshape(n, 3), x0, x1, w)[1]
        np.testing.assert_allclose(J[:, k], (rp - rm) / (2 * eps), atol=2e-7)


@pytest.mark.parametrize('args', [(0., 0., 0.), (0.1, 10., 0.2), (0.2, 0., 0.3)])
def test_lm_reaches_the_least_squares_minimum_of_the_gtest_scene(args):
    """Same residuals (ba_problem.h:60-151, camera 0 fixed), two solvers: oracle/mvba.solve (Ceres-style LM with a
    Schur complement) and scipy's trust-region least squares reach the same cost."""
    pb = M.gtest_problem([0.3, -0.2, 0.5, 0.3, -0.4, 0.5], *args)
    cams, pts, info = M.solve(pb)
    ncam = pb.cams.shape[0]

    def residuals(x):
        c = pb.cams.copy()
        c[1:] = x[:6 * (ncam - 1)].reshape(-1, 6)
        p = x[6 * (ncam - 1):].reshape(-1, 3)
        r, _, _, _ = M._residuals_and_jacobian(pb, c, p, want_J=False)
        return r

    x0 = np.concatenate([pb.cams[1:].reshape(-1), pb.points.reshape(-1)])
    assert abs(0.5 * (residuals(x0) ** 2).sum() - info['initial_cost']) <= 1e-12 * max(1.0, info['initial_cost'])
    ref = least_squares(residuals, x0, method='trf', xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=2000)
    # the oracle stops at Ceres' default tolerances (function tolerance 1e-6): never below the true minimum,
    # and within that tolerance of it
    assert info['final_cost'] >= ref.cost - 1e-12
    assert info['final_cost'] - ref.cost <= 1e-4 * max(info['initial_cost'], 1e-12) + 1e-12

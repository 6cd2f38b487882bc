"""CPU: cross-checks of the two-view / multi-view pose oracle against OpenCV 4.13 (an independent implementation
that IS in this image; cv2.triangulatePoints is also what the reference itself calls, bundle_adjust_io.py:222).
kornia / pytorch3d, which the reference uses for the rest, are absent -- their restatements in oracle/pose.py
stay "parity unpinned" (DESIGN.md §5); these tests bound how far off they can be."""
import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')

from oracle import pose as P, mvba as M


def _cams(rng):
    R = P.rodrigues(rng.standard_normal(3) * 0.2)
    t = rng.standard_normal(3) * 0.3
    P0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P1 = np.hstack([R, t[:, None]])
    return R, t, P0, P1


def _project(Pm, X):
    q = X @ Pm[:, :3].T + Pm[:, 3]
    return q[:, :2] / q[:, 2:3]


@pytest.mark.parametrize('noise', [0.0, 2e-3])
def test_triangulation_matches_cv2(noise):
    rng = np.random.default_rng(0)
    R, t, P0, P1 = _cams(rng)
    X = np.column_stack([rng.uniform(-1, 1, 200), rng.uniform(-1, 1, 200), rng.uniform(2, 6, 200)])
    x0 = _project(P0, X) + noise * rng.standard_normal((200, 2))
    x1 = _project(P1, X) + noise * rng.standard_normal((200, 2))
    ref = cv2.triangulatePoints(P0, P1, x0.T.copy(), x1.T.copy())
    ref = (ref[:3] / ref[3]).T
    np.testing.assert_allclose(M.triangulate_dlt(P0, P1, x0, x1), ref, rtol=1e-7, atol=1e-9)
    got = P.triangulate_points(P0[None], P1[None], x0[None], x1[None])[0]      # the kornia restatement
    np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9)
    if noise == 0.0:
        np.testing.assert_allclose(got, X, atol=1e-9)


def test_essential_decomposition_matches_cv2():
    rng = np.random.default_rng(1)
    for _ in range(20):
        R, t, _, _ = _cams(rng)
        E = P.hat(t / np.linalg.norm(t)) @ R
        R1, R2, tt = P.decompose_essential_matrix(E[None])
        c1, c2, ct = cv2.decomposeEssentialMat(E)
        ours = sorted([R1[0], R2[0]], key=lambda A: A[0, 0])
        theirs = sorted([c1, c2], key=lambda A: A[0, 0])
        for a, b in zip(ours, theirs):
            np.testing.assert_allclose(a, b, atol=1e-9)
            np.testing.assert_allclose(np.linalg.det(a), 1.0, atol=1e-9)
        assert min(np.abs(tt[0, :, 0] - ct[:, 0]).max(), np.abs(tt[0, :, 0] + ct[:, 0]).max()) < 1e-9
        # the true motion is among the four candidates
        Rs, ts = P.motion_from_essential(E[None])
        err = [np.abs(Rs[0, k] - R).max() + np.abs(ts[0, k, :, 0] - t / np.linalg.norm(t)).max() for k in range(4)]
        assert min(err) < 1e-9


def test_w8pt_agrees_with_cv2_recover_pose():
    """Noise-free correspondences, uniform confidences: the weighted eight-point restatement and OpenCV's
    findEssentialMat + recoverPose (cheirality test) return the same rotation and translation direction."""
    for seed in range(5):
        sc = P.make_two_view_scene(seed, 120, outlier_frac=0.0, noise_px=0.0, dtype=np.float64)
        conf = np.ones_like(sc['conf'])
        T, info = P.estimate_relative_pose_w8pt(sc['kpts0'], sc['kpts1'], sc['intr'], sc['intr'], conf)
        K = sc['intr'][0]
        E, _ = cv2.findEssentialMat(sc['kpts0'][0], sc['kpts1'][0], K, method=cv2.LMEDS)
        _, Rc, tc, _ = cv2.recoverPose(E[:3], sc['kpts0'][0], sc['kpts1'][0], K)
        np.testing.assert_allclose(T[0, :3, :3], Rc, atol=1e-6)
        t = T[0, :3, 3] / np.linalg.norm(T[0, :3, 3])
        np.testing.assert_allclose(t, tc[:, 0] / np.linalg.norm(tc), atol=1e-5)
        et, er = P.compute_pose_error(sc['T_021'][0], T[0, :3, :3], T[0, :3, 3])
        assert er < 1e-4 and et < 1e-3
        assert info['pos_depth_mask'].all()


def test_rodrigues_and_angle_axis_match_cv2():
    rng = np.random.default_rng(2)
    for _ in range(50):
        w = rng.standard_normal(3) * rng.uniform(1e-6, 3.0)
        Rc, _ = cv2.Rodrigues(w)
        np.testing.assert_allclose(P.rodrigues(w), Rc, atol=1e-12)
        np.testing.assert_allclose(M.angle_axis_to_R(w), Rc, atol=1e-12)
        wc, _ = cv2.Rodrigues(Rc)
        np.testing.assert_allclose(M.R_to_angle_axis(Rc), wc[:, 0], atol=1e-8)
